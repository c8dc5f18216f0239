/*
 * fp_hostpack.h -- host-side 2-bit packing of the bases of one read row (the host half of fp_packed_batch, include/fastp_b200.h).
 * Plain host C++ (built by g++, no CUDA): an AVX-512 (BW + VL) or AVX2 path picked at run time (FP_HOSTPACK_ISA=avx2 | swar forces a lower one), a 64-bit SWAR path otherwise.
 */
#pragma once
#include <stdint.h>
#include <vector>
#include "fastp_b200.h"

/* Packs bases s[0, L) into d[0, ceil(L/4)): code (base >> 1) & 3 (A0 C1 T2 G3), base i of a byte in bits 2i..2i+1; an 'N' packs as 0 and
 * is appended to `nl` as (unit, pos, which).  Returns 0, or 1 when a byte is outside {A,C,G,T,N} (not representable). */
int fp_pack_bases_row(const uint8_t* s, int L, uint8_t* d, uint32_t unit, int which, std::vector<fp_npos>& nl);
