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
