/*
 * synth.h -- deterministic synthetic FASTQ read-pair generator, compiled for BOTH the device
 * (fp_synth_fill in libfastp_b200.so, so benchmark inputs are born in HBM) and the host
 * (oracle/synth_host.c, so the CPU oracle can regenerate any batch bit-for-bit).
 * Integer arithmetic only: identical bytes on host and device by construction.
 *
 * Counter-based: pair `index` of stream `seed` depends on (seed, index) only.
 *
 * profile 0  "ref-style": mirrors the reference's own benchmark generator
 *            scripts/bench_e2e.sh:41-86 -- uniform ACGT, qualities from 512 pooled templates
 *            (positions <5 or >L-10 in Q20..35, else Q30..40), R1/R2 independent, fixed length.
 * profile 1  "enriched" (SURVEY.md 8d): fragment model with insert ~ N(300,80) (L<=200) or
 *            N(400,120), R1 = frag[0:L], R2 = revcomp(frag)[0:L]; insert < L => TruSeq adapter
 *            read-through (src/knownadapters.h:14-15) then poly-G fill; per-base substitution errors
 *            with p = 10^(-q/10); N bases; polyG / polyA tails; low-quality 3' tails; planted
 *            correctable mismatches (Q>=30 vs Q<=14); short / empty reads; adapter dimers.
 * profile 2  profile 1 plus single-base deletions (8 % of R2) / insertions (4 % of R1).
 * profile 3  profile 1 plus 20 planted over-represented sequences (40..150 bases, each in about 1 % of the fragments), so that
 *            Evaluator::computeOverRepSeq finds a non-empty candidate list (BASELINE.json configs[4]).
 */
#ifndef FP_SYNTH_H
#define FP_SYNTH_H
#include <stdint.h>

#ifdef __CUDACC__
#define FP_HD __host__ __device__ static __forceinline__
#else
#define FP_HD static inline
#endif

#define FP_SYNTH_MAXFRAG 1024

typedef struct { uint64_t s; } fp_rng;

FP_HD uint64_t fp_mix64(uint64_t z) {   /* splitmix64 finaliser */
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
FP_HD void fp_rng_seed(fp_rng* r, uint64_t seed, uint64_t index, uint64_t stream) {
    r->s = fp_mix64(seed + 0x9E3779B97F4A7C15ULL * (index * 4 + stream + 1));
}
FP_HD uint32_t fp_rng_u32(fp_rng* r) {
    r->s += 0x9E3779B97F4A7C15ULL;
    return (uint32_t)(fp_mix64(r->s) >> 32);
}
FP_HD uint32_t fp_rng_below(fp_rng* r, uint32_t n) {
    return (uint32_t)(((uint64_t)fp_rng_u32(r) * n) >> 32);
}
/* true with probability permille/1000 */
FP_HD int fp_rng_permille(fp_rng* r, uint32_t permille) { return fp_rng_below(r, 1000) < permille; }

FP_HD uint8_t fp_synth_base(uint32_t k) { return (uint8_t)("ACGT"[k & 3]); }
FP_HD uint8_t fp_synth_comp(uint8_t b) {
    return b == 'A' ? 'T' : b == 'T' ? 'A' : b == 'C' ? 'G' : b == 'G' ? 'C' : 'N';
}

/* round(2^32 * 10^(-q/10)) for q = 0..41 (q=0 saturated) */
FP_HD uint32_t fp_synth_err_thr(int q) {
    const uint32_t T[42] = {
        4294967295u, 3411628761u, 2709942490u, 2152575368u, 1709843830u, 1358170214u, 1078826017u, 856941961u,
        680692893u, 540693542u, 429496730u, 341162876u, 270994249u, 215257537u, 170984383u, 135817021u,
        107882602u, 85694196u, 68069289u, 54069354u, 42949673u, 34116288u, 27099425u, 21525754u,
        17098438u, 13581702u, 10788260u, 8569420u, 6806929u, 5406935u, 4294967u, 3411629u,
        2709942u, 2152575u, 1709844u, 1358170u, 1078826u, 856942u, 680693u, 540694u, 429497u, 341163u };
    if (q < 0) q = 0;
    if (q > 41) q = 41;
    return T[q];
}

FP_HD const char* fp_synth_adapter(int which) {   /* TruSeq R1 / R2, src/knownadapters.h:14-15 */
    return which ? "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT" : "AGATCGGAAGAGCACACGTCTGAACTCCAGTCA";
}

/* pooled quality template t (profile 0), position j */
FP_HD uint8_t fp_synth_pool_qual(uint64_t seed, uint32_t t, int j, int L) {
    fp_rng r; fp_rng_seed(&r, seed ^ 0x5151515151ULL, ((uint64_t)t << 16) + (uint64_t)j, 3);
    if (j < 5 || j > L - 10) return (uint8_t)(33 + 20 + fp_rng_below(&r, 16));
    return (uint8_t)(33 + 30 + fp_rng_below(&r, 11));
}

/*
 * Generate pair `index`.  seq*, qual* point at rows of `stride` bytes; bytes beyond the read
 * length are zero-filled.  For SE callers pass seq2 = qual2 = len2 = NULL.
 */
FP_HD void fp_synth_pair(uint64_t seed, uint64_t index, int profile, int L, int stride,
                         uint8_t* seq1, uint8_t* qual1, uint16_t* len1,
                         uint8_t* seq2, uint8_t* qual2, uint16_t* len2) {
    fp_rng r; fp_rng_seed(&r, seed, index, 0);
    int n1 = L, n2 = L, ins = L;
    if (profile == 0) {
        uint32_t t1 = fp_rng_below(&r, 512), t2 = fp_rng_below(&r, 512);
        for (int j = 0; j < L; j++) { seq1[j] = fp_synth_base(fp_rng_u32(&r) >> 13); qual1[j] = fp_synth_pool_qual(seed, t1, j, L); }
        if (seq2) for (int j = 0; j < L; j++) { seq2[j] = fp_synth_base(fp_rng_u32(&r) >> 13); qual2[j] = fp_synth_pool_qual(seed, t2, j, L); }
    } else {
        uint8_t frag[FP_SYNTH_MAXFRAG];
        /* insert size: mean + sd * z, z ~ Irwin-Hall(12) - 6 */
        int mean = L <= 200 ? 300 : 400, sd = L <= 200 ? 80 : 120;
        int64_t z = 0;
        for (int k = 0; k < 12; k++) z += (int64_t)(fp_rng_u32(&r) >> 16);
        z -= 6 * 65536;
        ins = mean + (int)((z * sd) / 65536);
        if (ins < 35) ins = 35;
        if (fp_rng_permille(&r, 3)) ins = (int)fp_rng_below(&r, 3);          /* adapter dimer: insert 0..2 */
        else if (fp_rng_permille(&r, 50)) ins = 20 + (int)fp_rng_below(&r, (uint32_t)L);   /* extra short inserts */
        if (profile == 2) {   /* many short inserts: only there can a gap near the overlap's end decide the outcome */
            fp_rng rj; fp_rng_seed(&rj, seed, index, 2);
            if (fp_rng_permille(&rj, 150)) ins = 30 + (int)fp_rng_below(&rj, 40);
        }
        if (ins > FP_SYNTH_MAXFRAG) ins = FP_SYNTH_MAXFRAG;
        int lowcomplex = fp_rng_permille(&r, 5);
        for (int j = 0; j < ins; j++) {
            uint32_t u = fp_rng_u32(&r);
            frag[j] = lowcomplex ? fp_synth_base((u >> 13) & 1 ? 0 : (u >> 20)) : fp_synth_base(u >> 13);
        }
        if (profile == 3) {   /* planted over-represented sequence k: content depends on (seed, k) only */
            fp_rng rp; fp_rng_seed(&rp, seed ^ 0x0E44E4ULL, index, 3);
            if (fp_rng_permille(&rp, 200)) {
                const uint32_t k = fp_rng_below(&rp, 20);
                const int Lk = 40 + (int)((k * 37u) % 111u);
                if (ins > Lk) {
                    const int pos = (int)fp_rng_below(&rp, (uint32_t)(ins - Lk + 1));
                    fp_rng rk; fp_rng_seed(&rk, seed ^ 0x51A17EDULL, (uint64_t)k, 1);
                    for (int j = 0; j < Lk; j++) frag[pos + j] = fp_synth_base(fp_rng_u32(&rk) >> 13);
                }
            }
        }
        /* read lengths: mostly L; 1% arbitrary in [0, L] */
        if (fp_rng_permille(&r, 10)) n1 = (int)fp_rng_below(&r, (uint32_t)L + 1);
        if (fp_rng_permille(&r, 10)) n2 = (int)fp_rng_below(&r, (uint32_t)L + 1);
        for (int which = 0; which < 2; which++) {
            uint8_t* s = which ? seq2 : seq1;
            uint8_t* q = which ? qual2 : qual1;
            int n = which ? n2 : n1;
            if (!s) { continue; }
            const char* ad = fp_synth_adapter(which);
            /* quality model */
            int binned = fp_rng_permille(&r, 500);
            int lowtail = fp_rng_permille(&r, 300) ? (L / 2 + (int)fp_rng_below(&r, (uint32_t)(L - L / 2))) : L + 1;
            int allbad = fp_rng_permille(&r, 10);
            for (int j = 0; j < n; j++) {
                /* true base */
                uint8_t b;
                if (j < ins) b = which ? fp_synth_comp(frag[ins - 1 - j]) : frag[j];
                else if (j - ins < 33) b = (uint8_t)ad[j - ins];
                else b = 'G';
                /* quality */
                int qv;
                uint32_t u = fp_rng_u32(&r);
                if (allbad) qv = 2 + (int)(u % 14);
                else if (j >= lowtail) qv = 2 + (int)(u % 14);
                else if (binned) { uint32_t k = u % 100; qv = k < 80 ? 37 : k < 92 ? 25 : k < 98 ? 11 : 2; }
                else if (j < 5 || j > L - 10) qv = 20 + (int)(u % 16);
                else qv = 30 + (int)(u % 11);
                /* substitution error with p = 10^(-q/10) */
                if (fp_rng_u32(&r) < fp_synth_err_thr(qv)) {
                    uint32_t k = fp_rng_below(&r, 3);
                    const char* alt = b == 'A' ? "CGT" : b == 'C' ? "AGT" : b == 'G' ? "ACT" : "ACG";
                    b = (uint8_t)alt[k];
                }
                if (fp_rng_permille(&r, 1)) b = 'N';
                s[j] = b; q[j] = (uint8_t)(33 + qv);
            }
            /* poly tails */
            uint32_t pt = fp_rng_below(&r, 1000);
            if (pt < 30 && n > 0) {
                int k = 10 + (int)fp_rng_below(&r, 41);
                uint8_t pb = pt < 20 ? 'G' : 'A';
                for (int j = n - k < 0 ? 0 : n - k; j < n; j++) if (fp_rng_below(&r, 16) != 0) s[j] = pb;
            }
            /* heavy-N reads */
            if (fp_rng_permille(&r, 3)) for (int k = 0; k < 10 && n > 0; k++) s[fp_rng_below(&r, (uint32_t)n)] = 'N';
        }
        /* planted correctable mismatch inside the overlapped part of the fragment */
        if (seq2 && fp_rng_permille(&r, 20) && ins >= 40 && ins < 2 * L - 35) {
            int lo = ins - L < 0 ? 0 : ins - L, hi = ins < L ? ins : L;       /* fragment coords present in both reads */
            if (hi > lo) {
                int f = lo + (int)fp_rng_below(&r, (uint32_t)(hi - lo));
                int p1 = f, p2 = ins - 1 - f;
                if (p1 < n1 && p2 < n2 && p2 >= 0) {
                    int flip1 = (int)fp_rng_below(&r, 2);
                    uint8_t* sb = flip1 ? &seq1[p1] : &seq2[p2];
                    uint8_t old = *sb;
                    *sb = old == 'A' ? 'C' : old == 'C' ? 'G' : old == 'G' ? 'T' : 'A';
                    uint8_t badq = (uint8_t)(33 + 2 + fp_rng_below(&r, 14));     /* Q2..15: 15 is NOT <= 14 */
                    uint8_t goodq = (uint8_t)(33 + 29 + fp_rng_below(&r, 12));   /* Q29..40: 29 is NOT >= 30 */
                    if (flip1) { qual1[p1] = badq; qual2[p2] = goodq; } else { qual2[p2] = badq; qual1[p1] = goodq; }
                }
            }
        }
    }
    if (profile == 2 && seq2) {
        /* profile 2 = enriched + single-base indels, so the one-gap overlap passes (--allow_gap_overlap_trimming) have work */
        fp_rng ri; fp_rng_seed(&ri, seed, index, 1);
        if (fp_rng_permille(&ri, 80) && n2 > 20) {                       /* deletion in R2 */
            /* half of them within the first 8 bases: the reference only accepts a gap whose shifted remainder stays
               within the mismatch limit (Matcher::diffWithOneInsertion's early return), i.e. gaps near the overlap's end */
            const int q = fp_rng_permille(&ri, 500) ? 1 + (int)fp_rng_below(&ri, 7) : 5 + (int)fp_rng_below(&ri, (uint32_t)(n2 - 10));
            for (int j = q; j + 1 < n2; j++) { seq2[j] = seq2[j + 1]; qual2[j] = qual2[j + 1]; }
            seq2[n2 - 1] = fp_synth_base(fp_rng_u32(&ri) >> 13);
        }
        if (fp_rng_permille(&ri, 40) && n1 > 20) {                       /* insertion in R1 */
            const int e1 = ins < n1 ? ins : n1;
            const int q = (e1 > 12 && fp_rng_permille(&ri, 500)) ? e1 - 2 - (int)fp_rng_below(&ri, 7) : 5 + (int)fp_rng_below(&ri, (uint32_t)(n1 - 10));
            for (int j = n1 - 1; j > q; j--) { seq1[j] = seq1[j - 1]; qual1[j] = qual1[j - 1]; }
            seq1[q] = fp_synth_base(fp_rng_u32(&ri) >> 13);
        }
    }
    for (int j = n1; j < stride; j++) { seq1[j] = 0; qual1[j] = 0; }
    *len1 = (uint16_t)n1;
    if (seq2) {
        for (int j = n2; j < stride; j++) { seq2[j] = 0; qual2[j] = 0; }
        *len2 = (uint16_t)n2;
    }
}

#endif /* FP_SYNTH_H */
