/*
 * fp_dup.cuh -- kernels of the duplication bloom filter: thin wrappers around the per-thread bodies of fp_dup.h
 * (the bodies are what tests/host/dup_emulation.cpp runs on the host against the oracle).
 */
#pragma once
#include "fp_device.cuh"
#include "fp_dup.h"

__global__ void __launch_bounds__(256) fp_dup_hash_kernel(fp_dup_state S, long long n, const uint8_t* seq1, const uint16_t* len1,
                                                          const uint8_t* seq2, const uint16_t* len2, int stride, int paired) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n) fp_dup_hash_unit(S, u, seq1, len1, seq2, len2, stride, paired);
}
/* pass H, one WARP per unit: lanes take the positions p = lane, lane+32, ... (coalesced row reads), partial sums are combined by
 * shuffles -- the hash is a sum modulo 2^64, so the order of the terms is free (same bits as fp_dup_hash_unit) */
__global__ void __launch_bounds__(256) fp_dup_hash_warp_kernel(fp_dup_state S, long long n, const uint8_t* __restrict__ seq1, const uint16_t* __restrict__ len1,
                                                               const uint8_t* __restrict__ seq2, const uint16_t* __restrict__ len2, int stride, int paired) {
    const long long u = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (u >= n) return;
    const int lane = threadIdx.x & 31;
    unsigned long long acc[FP_DUP_MAX_ARRAYS];
    #pragma unroll
    for (int i = 0; i < FP_DUP_MAX_ARRAYS; i++) acc[i] = 0;
    const int l1 = len1[u], l2 = paired ? len2[u] : 0;
    for (int side = 0; side < (paired ? 2 : 1); side++) {
        const uint8_t* r = (side ? seq2 : seq1) + u * stride;
        const int L = side ? l2 : l1, off = side ? l1 : 0;
        for (int p = lane; p < L; p += 32) {
            const int q = p + off;
            const unsigned long long base = fp_dup_hash_val(r[p]) + (unsigned long long)q;
            #pragma unroll
            for (int i = 0; i < FP_DUP_MAX_ARRAYS; i++)
                if (i < S.buf_num) acc[i] += S.primes[(uint64_t)(q * S.buf_num + i) & S.offset_mask] * base;
        }
    }
    #pragma unroll
    for (int i = 0; i < FP_DUP_MAX_ARRAYS; i++) {
        if (i >= S.buf_num) break;
        unsigned long long v = acc[i];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
        if (lane == 0) S.pos[u * S.buf_num + i] = v % S.buf_bits;
    }
}
__global__ void __launch_bounds__(256) fp_dup_first_kernel(fp_dup_state S, long long total) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total) fp_dup_first(S, t);
}
__global__ void __launch_bounds__(256) fp_dup_decide_kernel(fp_dup_state S, long long n, uint8_t* flags, unsigned long long* dups) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int dup = 0;
    if (u < n) { dup = fp_dup_decide(S, u); if (flags) flags[u] = (uint8_t)dup; }
    const unsigned m = __ballot_sync(FULL_MASK, dup != 0);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(dups, (unsigned long long)__popc(m));
}
__global__ void __launch_bounds__(256) fp_dup_commit_kernel(fp_dup_state S, long long total) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total) fp_dup_commit(S, t);
}
