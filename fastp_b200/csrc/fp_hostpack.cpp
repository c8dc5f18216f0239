/*
 * fp_hostpack.cpp -- 2-bit packing of read rows on the host (see fp_hostpack.h).  The end-to-end path is PCIe-bound, so the bases
 * cross the link at 2 bits each; this is the code that has to keep up with the link (about 40 GB/s of bases for a Gen5 x16 slot).
 */
#include "fp_hostpack.h"
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

/* four bases, byte by byte: the only place that sees an 'N' (or rejects a byte) */
inline int pack4_scalar(const uint8_t* s, int n, int pos0, uint8_t* out, uint32_t unit, int which, std::vector<fp_npos>& nl) {
    uint8_t o = 0;
    for (int i = 0; i < n; i++) {
        const uint8_t ch = s[i];
        if (ch == 'N') nl.push_back(fp_npos{unit, (uint16_t)(pos0 + i), (uint8_t)which, 0});
        else if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') return 1;
        else o |= (uint8_t)(((ch >> 1) & 3) << (2 * i));
    }
    *out = o;
    return 0;
}

/* 8 bases per step: a SWAR test "all of A/C/G/T", then one multiply per four codes */
inline int pack_swar(const uint8_t* s, int k, int L, uint8_t* d, uint32_t unit, int which, std::vector<fp_npos>& nl) {
    for (; k + 8 <= L; k += 8) {
        uint64_t x; memcpy(&x, s + k, 8);
        const uint64_t K = 0x0101010101010101ull;
        const uint64_t common = (x & 0xE8E8E8E8E8E8E8E8ull) ^ 0x4040404040404040ull;          /* bits 7,6,5 = 010, bit 3 = 0 */
        const uint64_t v1 = ((x >> 4) ^ x) & K;                                                  /* bit4 != bit0  (T <-> bit4) */
        const uint64_t v2 = ((((x >> 2) & ~(x >> 1)) ^ x)) & K;                                  /* (bit2 & !bit1) != bit0 */
        if (common == 0 && v1 == K && v2 == K) {
            const uint64_t c2 = (x >> 1) & 0x0303030303030303ull;                                 /* 2-bit codes, one per byte */
            const uint32_t lo4 = (uint32_t)c2, hi4 = (uint32_t)(c2 >> 32);
            d[k >> 2] = (uint8_t)((lo4 * 0x01041040u) >> 24);
            d[(k >> 2) + 1] = (uint8_t)((hi4 * 0x01041040u) >> 24);
            continue;
        }
        if (pack4_scalar(s + k, 4, k, d + (k >> 2), unit, which, nl)) return 1;
        if (pack4_scalar(s + k + 4, 4, k + 4, d + (k >> 2) + 1, unit, which, nl)) return 1;
    }
    for (; k < L; k += 4)
        if (pack4_scalar(s + k, L - k < 4 ? L - k : 4, k, d + (k >> 2), unit, which, nl)) return 1;
    return 0;
}

#if defined(__x86_64__)
/* 32 bases per step: code = (x >> 1) & 3; the block is clean iff "ACTG"[code] == x for every byte; then two multiply-adds fold four
 * codes into one byte (c0 + 4 c1, then v0 + 16 v1) and one shuffle collects the eight bytes */
__attribute__((target("avx2")))
int pack_avx2(const uint8_t* s, int L, uint8_t* d, uint32_t unit, int which, std::vector<fp_npos>& nl) {
    const __m256i tbl = _mm256_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 'A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i m3 = _mm256_set1_epi8(3);
    const __m256i w1 = _mm256_set1_epi16(0x0401);
    const __m256i w2 = _mm256_set1_epi32(0x00100001);
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    int k = 0;
    for (; k + 32 <= L; k += 32) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + k));
        const __m256i c = _mm256_and_si256(_mm256_srli_epi16(x, 1), m3);
        const __m256i y = _mm256_shuffle_epi8(tbl, c);
        if (_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, y)) != -1) {
            for (int j = 0; j < 32; j += 4)
                if (pack4_scalar(s + k + j, 4, k + j, d + ((k + j) >> 2), unit, which, nl)) return 1;
            continue;
        }
        const __m256i r = _mm256_shuffle_epi8(_mm256_madd_epi16(_mm256_maddubs_epi16(c, w1), w2), gather);
        const uint32_t lo = (uint32_t)_mm256_extract_epi32(r, 0), hi = (uint32_t)_mm256_extract_epi32(r, 4);
        memcpy(d + (k >> 2), &lo, 4); memcpy(d + (k >> 2) + 4, &hi, 4);
    }
    const int tail = L - k;
    if (tail > 0) {
        /* the row's last 1..31 bases: same block on a copy padded with 'A' (never reads past the row), only the bytes of the tail are stored */
        alignas(32) uint8_t buf[32];
        memset(buf, 'A', 32); memcpy(buf, s + k, (size_t)tail);
        const __m256i x = _mm256_load_si256(reinterpret_cast<const __m256i*>(buf));
        const __m256i c = _mm256_and_si256(_mm256_srli_epi16(x, 1), m3);
        const __m256i y = _mm256_shuffle_epi8(tbl, c);
        if (_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, y)) != -1) return pack_swar(s, k, L, d, unit, which, nl);
        const __m256i r = _mm256_shuffle_epi8(_mm256_madd_epi16(_mm256_maddubs_epi16(c, w1), w2), gather);
        const uint32_t lo = (uint32_t)_mm256_extract_epi32(r, 0), hi = (uint32_t)_mm256_extract_epi32(r, 4);
        uint8_t o[8]; memcpy(o, &lo, 4); memcpy(o + 4, &hi, 4);
        memcpy(d + (k >> 2), o, (size_t)((tail + 3) >> 2));
    }
    return 0;
}
#endif

#if defined(__x86_64__)
/* 64 bases per step (AVX-512 BW + VL): the same arithmetic on a zmm register; VPMOVDB takes the sixteen packed bytes out of the 32-bit lanes
 * in one instruction, and byte-masked loads / stores do the row's tail (a 150-base row is 64 + 64 + 22, no scalar remainder, no copy). */
__attribute__((target("avx512f,avx512bw,avx512vl")))
int pack_avx512(const uint8_t* s, int L, uint8_t* d, uint32_t unit, int which, std::vector<fp_npos>& nl) {
    const __m512i tbl = _mm512_broadcast_i32x4(_mm_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i m3 = _mm512_set1_epi8(3);
    const __m512i w1 = _mm512_set1_epi16(0x0401);
    const __m512i w2 = _mm512_set1_epi32(0x00100001);
    const __m512i isN = _mm512_set1_epi8('N');
    const __m512i padA = _mm512_set1_epi8('A');                            /* bases past the row pack as code 0, like the other paths */
    for (int k = 0; k < L; k += 64) {
        const int n = L - k < 64 ? L - k : 64;
        const __mmask64 lm = n == 64 ? ~0ull : ((1ull << n) - 1ull);
        const __m512i x = _mm512_mask_loadu_epi8(padA, lm, s + k);        /* masked-off bytes are not touched (no read past the row) */
        __m512i c = _mm512_and_si512(_mm512_srli_epi16(x, 1), m3);
        const __mmask64 okm = _mm512_cmpeq_epi8_mask(x, _mm512_shuffle_epi8(tbl, c));
        if (okm != ~0ull) {
            /* 'N' bases pack as code 0 and go on the list (in position order, like the byte-wise paths); any other byte: not representable */
            __mmask64 nm = _mm512_cmpeq_epi8_mask(x, isN);
            if ((okm | nm) != ~0ull) return 1;
            c = _mm512_maskz_mov_epi8(~nm, c);
            for (; nm; nm &= nm - 1) nl.push_back(fp_npos{unit, (uint16_t)(k + __builtin_ctzll(nm)), (uint8_t)which, 0});
        }
        const __m128i r = _mm512_cvtepi32_epi8(_mm512_madd_epi16(_mm512_maddubs_epi16(c, w1), w2));
        if (n == 64) _mm_storeu_si128(reinterpret_cast<__m128i*>(d + (k >> 2)), r);
        else _mm_mask_storeu_epi8(d + (k >> 2), (__mmask16)((1u << ((n + 3) >> 2)) - 1u), r);
    }
    return 0;
}

#endif

/* 3 = AVX-512 BW + VL, 2 = AVX2, 1 = SWAR; FP_HOSTPACK_ISA=avx2 | swar lowers it (tests run every path on one machine) */
int isa_level() {
    static const int v = []() {
        int lv = 1;
#if defined(__x86_64__)
        if (__builtin_cpu_supports("avx2")) lv = 2;
        if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl")) lv = 3;
#endif
        if (const char* e = getenv("FP_HOSTPACK_ISA")) {
            if (!strcmp(e, "swar")) lv = 1;
            else if (!strcmp(e, "avx2") && lv > 2) lv = 2;
        }
        return lv;
    }();
    return v;
}

}  // namespace

int fp_pack_bases_row(const uint8_t* s, int L, uint8_t* d, uint32_t unit, int which, std::vector<fp_npos>& nl) {
#if defined(__x86_64__)
    const int lv = isa_level();
    if (lv == 3) return pack_avx512(s, L, d, unit, which, nl);
    if (lv == 2) return pack_avx2(s, L, d, unit, which, nl);
#endif
    return pack_swar(s, 0, L, d, unit, which, nl);
}
