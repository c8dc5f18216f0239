/*
 * synth_host.c -- host instantiation of the shared synthetic generator (fastp_b200/csrc/synth.h) so the
 * CPU oracle can regenerate, bit for bit, any batch that fp_synth_fill produced in HBM.
 * TEST INFRASTRUCTURE (part of libfastp_oracle.so).
 */
#include "fastp_b200.h"
#include "../fastp_b200/csrc/synth.h"

int fp_synth_fill_host(const fp_batch* b, int64_t first_index, uint64_t seed, int32_t profile, int32_t read_len) {
    if (!b || read_len > b->stride || read_len < 16) return FP_E_INVAL;
    for (int64_t i = 0; i < b->n; i++) {
        int64_t o = i * b->stride;
        fp_synth_pair(seed, (uint64_t)(first_index + i), profile, read_len, b->stride,
                      b->seq1 + o, b->qual1 + o, b->len1 + i,
                      b->seq2 ? b->seq2 + o : 0, b->seq2 ? b->qual2 + o : 0, b->seq2 ? b->len2 + i : 0);
    }
    return FP_OK;
}
