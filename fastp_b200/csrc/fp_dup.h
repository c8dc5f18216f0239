/*
 * fp_dup.h -- duplication bloom filter (reference src/duplicate.cpp), per-thread bodies.
 *
 * The reference feeds reads / pairs one after the other through Duplicate::checkRead / checkPair (:126-154): a unit hashes to one
 * bit in each of bufNum bit arrays (seq2intvector :114-124) and is a duplicate iff all its bits were already set
 * (applyBloomFilter :156-169).  With plain atomicOr two equal units of one batch can each win one bit and both look new, so the
 * device form answers the sequential question directly: "in every array, was my bit set by an EARLIER unit?"
 *   pass H  positions of every unit (the hash)
 *   pass 1  first toucher: (array, bit) -> smallest unit index of this batch, an open-addressing table with atomicMin
 *   pass 2  unit i is a duplicate iff for every array its bit is set in the persistent arrays (earlier batches) or first < i
 *   pass 3  OR the batch's bits into the persistent arrays
 * Deterministic and equal to the reference fed in index order (tests/test_duplicate_oracle.py proves the formulation on CPU).
 *
 * Every pass is a function of the global thread index, compiled for the device (fp_dup.cuh wraps them in kernels) AND for the host:
 * tests/host/dup_emulation.cpp runs the very same bodies over the thread indices in shuffled order and compares with the oracle.
 */
#ifndef FP_DUP_H
#define FP_DUP_H
#include <stdint.h>

#ifdef __CUDACC__
#define FP_DUP_HD __host__ __device__ __forceinline__
#else
#define FP_DUP_HD static inline
#endif

#define FP_DUP_PRIME_LEN 512                       /* PRIME_ARRAY_LEN, duplicate.cpp:7 */
#define FP_DUP_MAX_ARRAYS 8
#define FP_DUP_EMPTY 0xFFFFFFFFFFFFFFFFull

typedef struct fp_dup_state {
    int buf_num;                                   /* mBufNum */
    uint64_t buf_bits;                             /* mBufLenInBits */
    uint64_t offset_mask;                          /* mOffsetMask */
    uint32_t* bits;                                /* [buf_num][buf_bits / 32] persistent arrays */
    const uint64_t* primes;                        /* [buf_num * 512] mPrimeArrays */
    /* per-batch scratch */
    uint64_t* pos;                                 /* [n][buf_num] */
    uint64_t* keys; uint32_t* vals; uint64_t table_mask;   /* first-toucher table, capacity table_mask + 1 (power of two) */
} fp_dup_state;

/* device: atomics; host emulation: the caller runs "threads" one at a time, so plain operations are atomic */
#if defined(__CUDA_ARCH__)
#define FP_DUP_CAS64(p, cmp, val) atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)(cmp), (unsigned long long)(val))
#define FP_DUP_MIN32(p, val) atomicMin((p), (val))
#define FP_DUP_OR32(p, val) atomicOr((p), (val))
#else
static inline uint64_t fp_dup_cas64_host(uint64_t* p, uint64_t cmp, uint64_t val) { const uint64_t old = *p; if (old == cmp) *p = val; return old; }
#define FP_DUP_CAS64(p, cmp, val) fp_dup_cas64_host((p), (cmp), (val))
#define FP_DUP_MIN32(p, val) do { if ((val) < *(p)) *(p) = (val); } while (0)
#define FP_DUP_OR32(p, val) do { *(p) |= (val); } while (0)
#endif

FP_DUP_HD uint64_t fp_dup_hash_val(uint8_t c) {    /* SEQ_HASH_VAL, duplicate.cpp:94-112 */
    return c == 'A' ? 7ull : c == 'T' ? 222ull : c == 'C' ? 74ull : c == 'G' ? 31ull : 13ull;
}
FP_DUP_HD uint64_t fp_dup_key(int array, uint64_t pos) { return ((uint64_t)array << 48) | pos; }   /* pos < 2^36 */
FP_DUP_HD uint64_t fp_dup_slot(uint64_t key, uint64_t mask) {
    uint64_t z = key * 0x9E3779B97F4A7C15ull; z ^= z >> 29;
    return z & mask;
}

/* pass H, one thread per unit: seq2intvector over r1 and (paired) r2 with posOffset = len1, then % buf_bits */
FP_DUP_HD void fp_dup_hash_unit(const fp_dup_state& S, long long u, const uint8_t* seq1, const uint16_t* len1, const uint8_t* seq2, const uint16_t* len2,
                                int stride, int paired) {
    uint64_t acc[FP_DUP_MAX_ARRAYS];
    for (int i = 0; i < S.buf_num; i++) acc[i] = 0;
    const int l1 = len1[u];
    const uint8_t* r = seq1 + (long long)u * stride;
    for (int p = 0; p < l1; p++) {
        const uint64_t base = fp_dup_hash_val(r[p]) + (uint64_t)p;
        for (int i = 0; i < S.buf_num; i++) acc[i] += S.primes[(uint64_t)(p * S.buf_num + i) & S.offset_mask] * base;
    }
    if (paired) {
        const int l2 = len2[u];
        r = seq2 + (long long)u * stride;
        for (int p = 0; p < l2; p++) {
            const int q = p + l1;
            const uint64_t base = fp_dup_hash_val(r[p]) + (uint64_t)q;
            for (int i = 0; i < S.buf_num; i++) acc[i] += S.primes[(uint64_t)(q * S.buf_num + i) & S.offset_mask] * base;
        }
    }
    for (int i = 0; i < S.buf_num; i++) S.pos[u * S.buf_num + i] = acc[i] % S.buf_bits;
}

/* pass 1, one thread per (unit, array) */
FP_DUP_HD void fp_dup_first(const fp_dup_state& S, long long t) {
    const long long u = t / S.buf_num; const int a = (int)(t % S.buf_num);
    const uint64_t key = fp_dup_key(a, S.pos[t]);
    uint64_t slot = fp_dup_slot(key, S.table_mask);
    for (uint64_t probes = 0; probes <= S.table_mask; probes++) {          /* the table is at most half full: ends after a few probes */
        const uint64_t prev = FP_DUP_CAS64(&S.keys[slot], FP_DUP_EMPTY, key);
        if (prev == FP_DUP_EMPTY || prev == key) { FP_DUP_MIN32(&S.vals[slot], (uint32_t)u); return; }
        slot = (slot + 1) & S.table_mask;
    }
}

/* pass 2, one thread per unit: returns 1 iff the reference would call it a duplicate */
FP_DUP_HD int fp_dup_decide(const fp_dup_state& S, long long u) {
    int dup = 1;
    for (int a = 0; a < S.buf_num; a++) {
        const uint64_t pos = S.pos[u * S.buf_num + a];
        const uint32_t word = S.bits[(uint64_t)a * (S.buf_bits >> 5) + (pos >> 5)];
        if ((word >> (pos & 31)) & 1u) continue;                          /* set by an earlier batch */
        const uint64_t key = fp_dup_key(a, pos);
        uint64_t slot = fp_dup_slot(key, S.table_mask);
        uint64_t probes = 0;
        while (S.keys[slot] != key && probes++ <= S.table_mask) slot = (slot + 1) & S.table_mask;     /* present: pass 1 inserted it */
        if (S.keys[slot] != key || !(S.vals[slot] < (uint32_t)u)) dup = 0;  /* nobody before me touched it */
    }
    return dup;
}

/* pass 3, one thread per (unit, array) */
FP_DUP_HD void fp_dup_commit(const fp_dup_state& S, long long t) {
    const int a = (int)(t % S.buf_num);
    const uint64_t pos = S.pos[t];
    FP_DUP_OR32(&S.bits[(uint64_t)a * (S.buf_bits >> 5) + (pos >> 5)], 1u << (pos & 31));
}

/* Duplicate::Duplicate sizing (duplicate.cpp:9-50) and initPrimeArrays (:68-86); host functions */
#include <math.h>
static inline void fp_dup_sizes(int accuracy_level, uint64_t* buf_bytes, int* buf_num) {
    uint64_t b = 1ull << 29; int n = 2;
    switch (accuracy_level) {
        case 2: b *= 2; break;
        case 3: b *= 2; n *= 2; break;
        case 4: b *= 4; n *= 2; break;
        case 5: b *= 8; n *= 2; break;
        case 6: b *= 8; n *= 4; break;
        default: break;
    }
    *buf_bytes = b; *buf_num = n;
}
static inline void fp_dup_primes(uint64_t* out, int count_wanted) {
    uint64_t number = 10000; int count = 0;
    while (count < count_wanted) {
        number++;
        int is_prime = 1;
        for (uint64_t i = 2; i <= sqrt((double)number); i++) if (number % i == 0) { is_prime = 0; break; }
        if (is_prime) { out[count++] = number; number += 10000; }
    }
}
#endif
