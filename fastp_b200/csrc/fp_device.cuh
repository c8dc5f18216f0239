/*
 * fp_device.cuh -- hand-written sm_100a device code for the per-read hot path.
 *
 * One persistent CTA per SM slot loops over tiles of TILE reads/pairs:
 *   phase 0  the tile's rows (contiguous in HBM) are pulled into shared memory by the TMA bulk-copy
 *            engine (cp.async.bulk + mbarrier complete_tx) -- one copy per array, each HBM byte read once
 *   phase 1  pre-filter Stats::statRead as a COLUMN pass: a thread owns 4 consecutive cycles of one read
 *            side, walks the tile's rows 4 at a time, transposes 4x4 bytes in registers (PRMT) and
 *            accumulates the per-cycle counters with dp4a into registers that live across all tiles
 *   phase 2  the per-read operator chain, one warp per read/pair, every scan in ballot / popc / ffs
 *            form (trimAndCut, trimPolyG, analyze, correction, adapter trimming, trimPolyX, passFilter)
 *   phase 3  post-filter statRead: column pass again over the (corrected, trimmed) rows that passed
 * Counters are block-privatised (registers / shared memory) and land in the global int64 block with
 * one atomicAdd per counter per CTA.
 *
 * Reference semantics (file:line under /root/reference/src) are cited at each operator; SURVEY.md
 * App. A lists the quirks that are reproduced on purpose.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fastp_b200.h"

#define FP_THREADS 256
#define FP_WARPS (FP_THREADS / 32)
#define FULL_MASK 0xffffffffu
#define FP_MAX_ISIZE_SMEM 1025

/* device-side parameter block (constant memory): fp_params + precomputed integer LUTs so that the
 * reference's three `double` predicates are evaluated on the host with its own expressions
 *  and the device stays pure integer. */
struct fp_dev_params {
    int paired, thread0;
    int trim_front1, trim_tail1, trim_front2, trim_tail2, max_len1, max_len2;
    int cut_front, cut_tail, cut_right;
    int cf_w, cf_thr, ct_w, ct_thr, cr_w, cr_thr, cr_q;      /* thr = w*(33+Q); cr_q = 33+Q */
    int polyg, polyg_min, polyx, polyx_min;
    int adapter_enabled, has_r1, has_r2, n_fasta, fasta_match_req, dimer_max_len;
    int correction, ov_require;
    int qual_filter, qualified_qual, n_base_limit, avg_qual_req;
    int length_filter, length_required, length_limit;
    int complexity_filter;
    int isize_max;
    int stride, cycles, tile, n_stats;
    int adapter_r1_off, adapter_r1_len, adapter_r2_off, adapter_r2_len;   /* into adapters blob */
    /* global-memory tables */
    const int16_t* lut_ovlimit;     /* [stride+1]  min(diffLimit, (int)(ol * (pct/100.0)))  overlapanalysis.cpp:51 */
    const int16_t* lut_lowq;        /* [stride+1]  floor(unqualifiedPercentLimit*rlen/100.0) filter.cpp:37          */
    const int16_t* lut_mindiff;     /* [stride+1]  smallest diff with diff/(len-1) >= threshold filter.cpp:65      */
    const uint8_t* adapters;        /* blob of adapter strings, each padded to a multiple of 4 + 8               */
    const int32_t* fasta_off;       /* [n_fasta] offsets into blob */
    const int32_t* fasta_len;       /* [n_fasta] */
    const uint32_t* adapter_planes; /* [2 + n_fasta][3][8] lo/hi/nn bit planes of each adapter (index 0 = r1, 1 = r2, 2+i = fasta i) */
    const uint8_t* adapter_clean;   /* [2 + n_fasta] 1 if the adapter holds only A,C,G,T,N */
    /* counter layout */
    fp_counter_layout L;
};

__constant__ fp_dev_params c_p;

/* ------------------------------------------------------------------ small helpers */
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* bytes of x that are non-zero -> bit 7 of each byte */
__device__ __forceinline__ uint32_t nz_bytes(uint32_t x) {
    return (x | ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu)) & 0x80808080u;
}
/* unaligned 32-bit little-endian load (two aligned loads + funnel shift); reads up to 7 bytes past p */
__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t* p) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    uint32_t lo = w[0], hi = w[1];
    return __funnelshift_r(lo, hi, (unsigned)(a & 3) * 8);
}
/* fire-and-forget 64-bit add to the global counter block (two's complement for negative deltas): RED, not ATOM */
__device__ __forceinline__ void red_add64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "l"(v) : "memory");
}
__device__ __forceinline__ uint8_t dev_complement(uint8_t b) {   /* util.h:16-33 */
    switch (b) {
        case 'A': case 'a': return 'T';
        case 'T': case 't': return 'A';
        case 'C': case 'c': return 'G';
        case 'G': case 'g': return 'C';
        default: return 'N';
    }
}
__device__ __forceinline__ int warp_sum(int v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ int warp_min(int v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

/* first i in [0,n) with pred(i) (lanes evaluate 32 consecutive i at a time), else n. warp-uniform result. */
template <class F>
__device__ __forceinline__ int warp_find_first(int n, F pred) {
    const int lane = lane_id();
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        bool p = (i < n) && pred(i);
        unsigned m = __ballot_sync(FULL_MASK, p);
        if (m) return base + __ffs(m) - 1;
    }
    return n;
}

/* a read as the warp sees it: smem row pointers (already advanced by front), length, NULL flag */
struct WRead {
    uint8_t* seq;
    uint8_t* qual;
    int len;
    int front;     /* bytes the pointers were advanced from the row start */
    bool null;
};

/* shared-memory block-level counters */
struct BlockCounters {
    unsigned int fr[FP_FR_WORDS];
    unsigned int isize[FP_MAX_ISIZE_SMEM];
};

/* ------------------------------------------------------------------------------------------------
 * Filter::trimAndCut  (filter.cpp:68-207).  P = per-warp scratch for prefix sums of quals, int[len+1].
 * ------------------------------------------------------------------------------------------------ */
__device__ __noinline__ void dev_trim_and_cut(WRead& r, int front, int tail, int* P) {
    const int lane = lane_id();
    const bool anycut = c_p.cut_front || c_p.cut_tail || c_p.cut_right;
    if (front == 0 && tail == 0 && !anycut) return;                       /* :71-72 */
    int rlen = r.len - front - tail;                                      /* :75 */
    if (rlen < 0) { r.null = true; return; }
    if (!anycut) {                                                        /* :79-89 */
        r.seq += front; r.qual += front; r.front += front; r.len = rlen;
        return;
    }
    const int l = r.len;
    const signed char* q = reinterpret_cast<const signed char*>(r.qual);
    const uint8_t* s = r.seq;
    /* prefix sums P[k] = sum_{j<k} q[j] */
    {
        int carry = 0;
        if (lane == 0) P[0] = 0;
        for (int base = 0; base < l; base += 32) {
            int i = base + lane;
            int v = (i < l) ? (int)q[i] : 0;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL_MASK, v, o); if (lane >= o) v += t; }
            if (i < l) P[i + 1] = carry + v;
            carry += __shfl_sync(FULL_MASK, v, 31);
        }
        __syncwarp();
    }
    if (c_p.cut_front) {                                                  /* :97-127 */
        const int w = c_p.cf_w, thr = c_p.cf_thr;
        if (l - front - tail - w <= 0) { r.null = true; return; }
        const int nwin = (l - tail - w) - front;                          /* s = front .. l-tail-w-1 */
        int k = warp_find_first(nwin, [&](int i) { int sp = front + i; return P[sp + w] - P[sp] >= thr; });
        int sp = front + k;                                               /* not found: k == nwin -> s = l-tail-w */
        if (sp > 0) sp = sp + w - 1;
        /* while (s<l && seq[s]=='N') s++ */
        int nn = warp_find_first(l - sp > 0 ? l - sp : 0, [&](int i) { return s[sp + i] != 'N'; });
        sp += nn;
        front = sp;
        rlen = l - front - tail;
    }
    if (c_p.cut_right) {                                                  /* :130-163 */
        const int w = c_p.cr_w, thr = c_p.cr_thr, qthr = c_p.cr_q;
        if (l - front - tail - w <= 0) { r.null = true; return; }
        const int nwin = (l - tail - w) - front;
        int k = warp_find_first(nwin, [&](int i) { int sp = front + i; return P[sp + w] - P[sp] < thr; });
        if (k < nwin) {
            int sp = front + k;
            /* while (s < l-1 && q[s] >= 33+Q) s++ */
            int span = l - 1 - sp;
            int nn = warp_find_first(span > 0 ? span : 0, [&](int i) { return (int)q[sp + i] < qthr; });
            sp += nn;
            rlen = sp - front;
        }
    }
    if (!c_p.cut_right && c_p.cut_tail) {                                 /* :166-194 */
        const int w = c_p.ct_w, thr = c_p.ct_thr;
        if (l - front - tail - w <= 0) { r.null = true; return; }
        const int t0 = l - tail - 1;
        const int nwin = t0 - (front + w) + 1;                            /* t = t0 .. front+w */
        int k = warp_find_first(nwin, [&](int i) { int t = t0 - i; return P[t + 1] - P[t - w + 1] >= thr; });
        int t = t0 - k;                                                   /* not found: t = front+w-1 */
        if (t < l - 1) t = t - w + 1;
        /* while (t>=0 && seq[t]=='N') t-- */
        int nn = warp_find_first(t + 1 > 0 ? t + 1 : 0, [&](int i) { return s[t - i] != 'N'; });
        t -= nn;
        rlen = t - front + 1;
    }
    if (rlen <= 0 || front >= l - 1) { r.null = true; return; }           /* :196-197 */
    r.seq += front; r.qual += front; r.front += front; r.len = rlen;      /* :199-204 */
}

/* ------------------------------------------------------------------------------------------------
 * PolyX::trimPolyG  (polyx.cpp:16-42).  Returns true if the read was shortened.
 * ------------------------------------------------------------------------------------------------ */
__device__ __noinline__ bool dev_trim_polyg(WRead& r, int minLen) {
    const int lane = lane_id();
    const int rlen = r.len;
    const uint8_t* data = r.seq;
    int cum = 0;                 /* mismatches before this chunk */
    int lastG = -1;              /* highest scan index i (<= break) holding a G */
    int ibreak = rlen;
    for (int base = 0; base < rlen; base += 32) {
        int i = base + lane;
        bool valid = i < rlen;
        bool isG = valid && data[rlen - 1 - i] == 'G';
        unsigned mm = __ballot_sync(FULL_MASK, valid && !isG);
        unsigned gm = __ballot_sync(FULL_MASK, isG);
        int mismatch = cum + __popc(mm & (0xffffffffu >> (31 - lane)));
        bool brk = valid && (mismatch > 5 || (mismatch > (i + 1) / 8 && i >= minLen - 1));
        unsigned bm = __ballot_sync(FULL_MASK, brk);
        if (bm) {
            int bl = __ffs(bm) - 1;
            ibreak = base + bl;
            unsigned g = gm & (0xffffffffu >> (31 - bl));
            if (g) lastG = base + 31 - __clz(g);
            break;
        }
        if (gm) lastG = base + 31 - __clz(gm);
        cum += __popc(mm);
    }
    if (ibreak >= minLen) {
        int firstGPos = lastG >= 0 ? rlen - 1 - lastG : rlen - 1;
        if (firstGPos >= 0 && firstGPos <= r.len && firstGPos != r.len) { r.len = firstGPos; return true; }
    }
    return false;
}

/* ------------------------------------------------------------------------------------------------
 * PolyX::trimPolyX  (polyx.cpp:49-116).  Returns true if addPolyXTrimmed is called (poly, n out).
 * ------------------------------------------------------------------------------------------------ */
__device__ __noinline__ bool dev_trim_polyx(WRead& r, int minLen, int& polyOut, int& nOut) {
    const int lane = lane_id();
    const int rlen = r.len;
    const uint8_t* data = r.seq;
    int cumA = 0, cumT = 0, cumC = 0, cumG = 0;
    int cntA = 0, cntT = 0, cntC = 0, cntG = 0;          /* counts at the final pos */
    int pos = rlen;
    for (int base = 0; base < rlen; base += 32) {
        int i = base + lane;
        bool valid = i < rlen;
        uint8_t c = valid ? data[rlen - 1 - i] : 0;
        bool n = (c == 'N');
        unsigned mA = __ballot_sync(FULL_MASK, c == 'A' || n);
        unsigned mT = __ballot_sync(FULL_MASK, c == 'T' || n);
        unsigned mC = __ballot_sync(FULL_MASK, c == 'C' || n);
        unsigned mG = __ballot_sync(FULL_MASK, c == 'G' || n);
        unsigned le = 0xffffffffu >> (31 - lane);
        int a = cumA + __popc(mA & le), t = cumT + __popc(mT & le), cc = cumC + __popc(mC & le), g = cumG + __popc(mG & le);
        int cmp = i + 1;
        int allowed = min(5, cmp / 8);
        bool need = (cmp - a > allowed) && (cmp - t > allowed) && (cmp - cc > allowed) && (cmp - g > allowed);
        bool brk = valid && need && (i >= 8 || i + 1 >= minLen - 1);
        unsigned bm = __ballot_sync(FULL_MASK, brk);
        if (bm) {
            int bl = __ffs(bm) - 1;
            pos = base + bl;
            cntA = __shfl_sync(FULL_MASK, a, bl); cntT = __shfl_sync(FULL_MASK, t, bl);
            cntC = __shfl_sync(FULL_MASK, cc, bl); cntG = __shfl_sync(FULL_MASK, g, bl);
            break;
        }
        cumA += __popc(mA); cumT += __popc(mT); cumC += __popc(mC); cumG += __popc(mG);
        cntA = cumA; cntT = cumT; cntC = cumC; cntG = cumG;
    }
    if (pos + 1 >= minLen) {                                              /* :96-115 */
        int poly = 0, mx = cntA;
        if (cntT > mx) { mx = cntT; poly = 1; }
        if (cntC > mx) { mx = cntC; poly = 2; }
        if (cntG > mx) { mx = cntG; poly = 3; }
        const uint8_t polyBase = (poly == 0) ? 'A' : (poly == 1) ? 'T' : (poly == 2) ? 'C' : 'G';
        /* largest scan index i <= min(pos, rlen-1) with data[rlen-1-i] == polyBase, else -1 */
        int top = min(pos, rlen - 1);
        int found = -1;
        for (int base = (top >= 0 ? (top & ~31) : -32); base >= 0; base -= 32) {
            int i = base + lane;
            bool hit = (i <= top) && data[rlen - 1 - i] == polyBase;
            unsigned hm = __ballot_sync(FULL_MASK, hit);
            if (hm) { found = base + 31 - __clz(hm); break; }
        }
        int newlen = rlen - found - 1;                                    /* resize(rlen - pos - 1) */
        if (newlen >= 0 && newlen <= r.len) r.len = newlen;
        polyOut = poly; nOut = found + 1;
        return true;
    }
    return false;
}

/* ------------------------------------------------------------------------------------------------
 * Bit planes of a read ROW (clean rows only: every base in {A,C,G,T,N}).  Bit p = row position p.
 * code = (base>>1)&3 : A0 C1 T2 G3;  lo = code bit0, hi = code bit1, both 0 under N;  nn = N mask;
 * lq = (qual < qualified_qual).  Two bases differ iff (lo^lo') | (hi^hi') | (nn^nn').
 * Each plane holds pw = stride/32 + 2 words, zero beyond the row, so any 32-bit field is one funnel
 * shift (plane_bits).  Planes of every row of a tile are built once, word-parallel (phase 0.5); users
 * address the trimmed window as bit (front + k) and mask by the current length.
 * ------------------------------------------------------------------------------------------------ */
struct Planes { uint32_t *lo, *hi, *nn, *lq; };

__device__ __forceinline__ uint32_t plane_bits(const uint32_t* P, int bit) {
    const int w = bit >> 5;
    return __funnelshift_r(P[w], P[w + 1], bit & 31);
}
__device__ __forceinline__ uint32_t low_mask(int nbits) {     /* mask of min(max(nbits,0),32) low bits */
    return nbits >= 32 ? 0xffffffffu : (nbits <= 0 ? 0u : ((1u << nbits) - 1u));
}
/* 4 bytes holding 0/1 -> 4-bit nibble, byte k -> bit k */
__device__ __forceinline__ uint32_t pack_nibble(uint32_t v01) { return ((v01 * 0x01020408u) >> 24) & 0xFu; }

/* One plane word (32 bases = 8 seq words + 8 qual words) of one row: returns false if a valid byte is outside
 * {A,C,G,T,N} or a quality has bit 7 set.  x/q: the 8 words; n = number of valid bases in this word (0..32). */
__device__ __forceinline__ bool plane_word_from_bytes(const uint32_t (&x)[8], const uint32_t (&q)[8], int n, uint32_t qq4,
                                                      uint32_t& lo, uint32_t& hi, uint32_t& nn, uint32_t& lq) {
    const uint32_t K = 0x01010101u;
    lo = hi = nn = lq = 0;
    uint32_t bad = 0;
    #pragma unroll
    for (int k = 0; k < 8; k++) {
        const int nv = n - 4 * k;                                   /* valid bytes of this word */
        const uint32_t vm = nv >= 4 ? K : (nv <= 0 ? 0u : (K & ((1u << (8 * nv)) - 1u)));
        const uint32_t w = x[k];
        const uint32_t c0 = w & K, c1 = (w >> 1) & K, c2 = (w >> 2) & K, b3 = (w >> 3) & K, b4 = (w >> 4) & K;
        const uint32_t up = (~(w >> 5)) & (w >> 6) & (~(w >> 7)) & K;         /* bits 7..5 == 010 */
        const uint32_t isACG = c0 & (~c2 | c1) & ~b3 & ~b4;                     /* 0x41 0x43 0x47 */
        const uint32_t isT = ~c0 & ~c1 & c2 & ~b3 & b4;                          /* 0x54 */
        const uint32_t isN = ~c0 & c1 & c2 & b3 & ~b4 & up & vm;                 /* 0x4E */
        const uint32_t acgt = (isACG | isT) & up & vm;
        bad |= vm & ~(acgt | isN);
        bad |= (q[k] >> 7) & vm;
        /* q < qualified_qual  <=>  bit7 of (q | 0x80) - qq is clear (q, qq < 128) */
        const uint32_t ql = (~(((q[k] | 0x80808080u) - qq4) >> 7)) & vm;
        lo |= pack_nibble(c1 & acgt) << (4 * k);
        hi |= pack_nibble(c2 & acgt) << (4 * k);
        nn |= pack_nibble(isN) << (4 * k);
        lq |= pack_nibble(ql) << (4 * k);
    }
    return bad == 0;
}

/* ballot-based rebuild of one row's planes (used after base correction rewrote the row; rare) */
__device__ __noinline__ void dev_rebuild_planes(const uint8_t* seq, const uint8_t* qual, int len, int pw, Planes P) {
    const int lane = lane_id();
    const uint8_t qq = (uint8_t)c_p.qualified_qual;
    for (int w = 0; w < pw; w++) {
        const int i = w * 32 + lane;
        const bool valid = i < len;
        uint8_t b = 0, q = 255;
        if (valid) { b = seq[i]; q = qual[i]; }
        const bool isN = (b == 'N');
        const int c2 = (b >> 1) & 3;
        const bool acgt = valid && !isN;
        const unsigned mlo = __ballot_sync(FULL_MASK, acgt && (c2 & 1));
        const unsigned mhi = __ballot_sync(FULL_MASK, acgt && (c2 & 2));
        const unsigned mn = __ballot_sync(FULL_MASK, isN);
        const unsigned mq = __ballot_sync(FULL_MASK, valid && q < qq);
        if (lane == 0) { P.lo[w] = mlo; P.hi[w] = mhi; P.nn[w] = mn; P.lq[w] = mq; }
    }
    __syncwarp();
}

/* reverseComplement(r2 window) as planes RC (relative to bit 0), from r2's row planes by bit reversal:
 * RC[k] = complement(row2[e - k]), e = front2 + len2 - 1.  complement flips code bit1 (A0<->T2, C1<->G3), N stays N. */
__device__ __forceinline__ void dev_rc_planes(const Planes& P2, int front2, int len2, int pw, Planes RC) {
    const int lane = lane_id();
    if (lane < pw) {
        const int e = front2 + len2 - 1;
        const int s0 = e - 32 * lane - 31;                           /* row position of the field's bit 0 (before reversal) */
        uint32_t flo, fhi, fnn;
        if (s0 >= 0) { flo = plane_bits(P2.lo, s0); fhi = plane_bits(P2.hi, s0); fnn = plane_bits(P2.nn, s0); }
        else if (s0 > -32) { flo = P2.lo[0] << (-s0); fhi = P2.hi[0] << (-s0); fnn = P2.nn[0] << (-s0); }
        else { flo = fhi = fnn = 0; }
        const uint32_t vm = low_mask(len2 - 32 * lane);
        const uint32_t n = __brev(fnn) & vm;
        RC.nn[lane] = n;
        RC.lo[lane] = __brev(flo) & vm & ~n;
        RC.hi[lane] = ~__brev(fhi) & vm & ~n;
    }
    __syncwarp();
}

/* OverlapAnalysis::analyze on bit planes (both rows clean).  Same candidate order and acceptance rule as the byte
 * version below (overlapanalysis.cpp:34-89); per candidate offset: funnel shifts + xor/or + 2 popc, no loop. */
__device__ __noinline__ fp_ov_result dev_analyze_planes(int len1, int front1, int len2, int front2, Planes A, Planes P2, Planes RC, int pw, const int16_t* lut) {
    const int lane = lane_id();
    dev_rc_planes(P2, front2, len2, pw, RC);
    const int req = c_p.ov_require;
    fp_ov_result ov; ov.overlapped = 0; ov.has_gap = 0; ov.offset = 0; ov.overlap_len = 0; ov.diff = 0;
    const int nfwd = max(len1 - req, 0), nbwd = max(len2 - req, 0);
    /* constant sides: first 64 bases of rc(r2) (forward scan) and of r1 (backward scan) */
    const uint32_t blo0 = RC.lo[0], blo1 = RC.lo[1], bhi0 = RC.hi[0], bhi1 = RC.hi[1], bnn0 = RC.nn[0], bnn1 = RC.nn[1];
    const uint32_t alo0 = plane_bits(A.lo, front1), alo1 = plane_bits(A.lo, front1 + 32), ahi0 = plane_bits(A.hi, front1), ahi1 = plane_bits(A.hi, front1 + 32),
                   ann0 = plane_bits(A.nn, front1), ann1 = plane_bits(A.nn, front1 + 32);
    for (int dir = 0; dir < 2; dir++) {
        const int ncand = dir == 0 ? nfwd : nbwd;
        for (int base = 0; base < ncand; base += 32) {
            const int o = base + lane;
            const bool valid = o < ncand;
            int ol = 0, limit = -1, mm = 0;
            if (valid) {
                ol = dir == 0 ? min(len1 - o, len2) : min(len1, len2 - o);
                limit = lut[ol];
                const int pp = min(ol, 50);                                /* complete_compare_require :29 */
                uint32_t x0, x1;
                if (dir == 0) {
                    const int bit = front1 + o, w = bit >> 5, sh = bit & 31;
                    x0 = (__funnelshift_r(A.lo[w], A.lo[w + 1], sh) ^ blo0) | (__funnelshift_r(A.hi[w], A.hi[w + 1], sh) ^ bhi0) | (__funnelshift_r(A.nn[w], A.nn[w + 1], sh) ^ bnn0);
                    x1 = (__funnelshift_r(A.lo[w + 1], A.lo[w + 2], sh) ^ blo1) | (__funnelshift_r(A.hi[w + 1], A.hi[w + 2], sh) ^ bhi1) | (__funnelshift_r(A.nn[w + 1], A.nn[w + 2], sh) ^ bnn1);
                } else {
                    const int w = o >> 5, sh = o & 31;
                    x0 = (__funnelshift_r(RC.lo[w], RC.lo[w + 1], sh) ^ alo0) | (__funnelshift_r(RC.hi[w], RC.hi[w + 1], sh) ^ ahi0) | (__funnelshift_r(RC.nn[w], RC.nn[w + 1], sh) ^ ann0);
                    x1 = (__funnelshift_r(RC.lo[w + 1], RC.lo[w + 2], sh) ^ alo1) | (__funnelshift_r(RC.hi[w + 1], RC.hi[w + 2], sh) ^ ahi1) | (__funnelshift_r(RC.nn[w + 1], RC.nn[w + 2], sh) ^ ann1);
                }
                mm = __popc(x0 & low_mask(pp)) + __popc(x1 & low_mask(pp - 32));
            }
            const unsigned am = __ballot_sync(FULL_MASK, valid && mm <= limit);
            if (am) {
                const int wl = __ffs(am) - 1;
                const int wo = base + wl;
                const int wol = __shfl_sync(FULL_MASK, ol, wl);
                int diff = __shfl_sync(FULL_MASK, mm, wl);
                if (wol > 50) {                                            /* :41-43 full recount, lanes over words */
                    int d = 0;
                    const int abit = front1 + (dir == 0 ? wo : 0), bbit = dir == 0 ? 0 : wo;
                    for (int w = lane; w * 32 < wol; w += 32) {
                        const uint32_t x = (plane_bits(A.lo, abit + 32 * w) ^ plane_bits(RC.lo, bbit + 32 * w)) | (plane_bits(A.hi, abit + 32 * w) ^ plane_bits(RC.hi, bbit + 32 * w)) |
                                           (plane_bits(A.nn, abit + 32 * w) ^ plane_bits(RC.nn, bbit + 32 * w));
                        d += __popc(x & low_mask(wol - 32 * w));
                    }
                    diff = warp_sum(d);
                }
                ov.overlapped = 1; ov.offset = (int16_t)(dir == 0 ? wo : -wo); ov.overlap_len = (int16_t)wol; ov.diff = (int16_t)diff;
                return ov;
            }
        }
    }
    return ov;
}

/* Filter::passFilter on bit planes (clean row; window = bits [front, front+rlen)).  filter.cpp:15-57 */
__device__ __noinline__ int dev_pass_filter_planes(const uint8_t* qual, int rlen, bool null, Planes P, int front, int pw, const int16_t* lut) {
    if (null || rlen == 0) return FP_FAIL_LENGTH;
    const int lane = lane_id();
    int packed = 0;                                                        /* lowq | nb << 10 | adj << 20 (each <= 512) */
    if (lane * 32 < rlen) {
        const int bit = front + 32 * lane;
        const uint32_t m = low_mask(rlen - 32 * lane);
        const uint32_t nn = plane_bits(P.nn, bit);
        packed = __popc(plane_bits(P.lq, bit) & m) | (__popc(nn & m) << 10);
        if (c_p.complexity_filter) {
            const uint32_t lo = plane_bits(P.lo, bit), hi = plane_bits(P.hi, bit);
            const uint32_t m1 = low_mask(rlen - 1 - 32 * lane);            /* pairs (i, i+1), i < rlen-1 */
            const uint32_t d = (lo ^ plane_bits(P.lo, bit + 1)) | (hi ^ plane_bits(P.hi, bit + 1)) | (nn ^ plane_bits(P.nn, bit + 1));
            packed |= __popc(d & m1) << 20;
        }
    }
    packed = warp_sum(packed);
    const int lowq = packed & 0x3FF, nb = (packed >> 10) & 0x3FF, adj = packed >> 20;
    if (c_p.qual_filter) {
        if (lowq > (int)lut[(c_p.stride + 2) + rlen]) return FP_FAIL_QUALITY;
        if (c_p.avg_qual_req > 0) {
            int tq = 0;
            for (int i = lane; i < rlen; i += 32) tq += (int)qual[i] - 33;
            tq = warp_sum(tq);
            if ((tq / rlen) < c_p.avg_qual_req) return FP_FAIL_QUALITY;
        }
        if (nb > c_p.n_base_limit) return FP_FAIL_N_BASE;
    }
    if (c_p.length_filter) {
        if (rlen < c_p.length_required) return FP_FAIL_LENGTH;
        if (c_p.length_limit > 0 && rlen > c_p.length_limit) return FP_FAIL_TOO_LONG;
    }
    if (c_p.complexity_filter) {
        if (rlen <= 1) return FP_FAIL_COMPLEXITY;
        if (adj < (int)lut[2 * (c_p.stride + 2) + rlen]) return FP_FAIL_COMPLEXITY;
    }
    return FP_PASS_FILTER;
}

/* ------------------------------------------------------------------------------------------------
 * OverlapAnalysis::analyze    if (c_p.length_filter) {
        if (rlen < c_p.length_required) return FP_FAIL_LENGTH;
        if (c_p.length_limit > 0 && rlen > c_p.length_limit) return FP_FAIL_TOO_LONG;
    }
    if (c_p.complexity_filter) {
        if (rlen <= 1) return FP_FAIL_COMPLEXITY;
        adj = warp_sum(adj);
        if (adj < (int)lut[2 * (c_p.stride + 2) + rlen]) return FP_FAIL_COMPLEXITY;
    }
    return FP_PASS_FILTER;
}

/* ------------------------------------------------------------------------------------------------
 * OverlapAnalysis::analyze  (overlapanalysis.cpp:17-146, allowGap=false).
 * rc = per-warp scratch holding reverseComplement(r2) (simd.cpp:297-310), padded by 8 readable bytes.
 * Lanes enumerate 32 candidate offsets at a time in the reference's order (forward 0,1,.. then
 * backward 0,-1,..); each lane counts mismatches over the protected prefix 4 bytes per step with
 * early exit; the lowest accepting lane wins.
 * ------------------------------------------------------------------------------------------------ */
__device__ __forceinline__ int dev_count_mismatch_coop(const uint8_t* a, const uint8_t* b, int n) {
    /* warp-cooperative Hamming distance over n bytes (countMismatches simd.cpp:320-324) */
    int d = 0;
    for (int i = lane_id(); i < n; i += 32) d += (a[i] != b[i]);
    return warp_sum(d);
}

__device__ __noinline__ fp_ov_result dev_analyze(const WRead& r1, const WRead& r2, uint8_t* rc, const int16_t* lut) {
    const int lane = lane_id();
    const int len1 = r1.len, len2 = r2.len;
    for (int i = lane; i < len2; i += 32) rc[i] = dev_complement(r2.seq[len2 - 1 - i]);
    __syncwarp();
    const int req = c_p.ov_require;
    fp_ov_result ov; ov.overlapped = 0; ov.has_gap = 0; ov.offset = 0; ov.overlap_len = 0; ov.diff = 0;
    const uint8_t* s1 = r1.seq;
    const uint32_t* rcw = reinterpret_cast<const uint32_t*>(rc);
    const int nfwd = max(len1 - req, 0);           /* forward offsets 0 .. len1-req-1  (:48) */
    const int nbwd = max(len2 - req, 0);           /* backward offsets 0 .. -(len2-req-1) (:73) */
    for (int dir = 0; dir < 2; dir++) {
        const int ncand = dir == 0 ? nfwd : nbwd;
        for (int base = 0; base < ncand; base += 32) {
            const int o = base + lane;
            const bool valid = o < ncand;
            int ol = 0, limit = 0, pp = 0, mm = 0;
            if (valid) {
                ol = dir == 0 ? min(len1 - o, len2) : min(len1, len2 - o);
                limit = lut[ol];
                pp = min(ol, 50);                                          /* complete_compare_require :29 */
            }
            bool active = valid && pp > 0;
            int k = 0;
            while (__any_sync(FULL_MASK, active)) {
                if (active) {
                    uint32_t a, b;
                    if (dir == 0) { a = ld_u32_unaligned(s1 + o + k); b = rcw[k >> 2]; }
                    else          { a = ld_u32_unaligned(s1 + k); b = ld_u32_unaligned(rc + o + k); }
                    uint32_t x = a ^ b;
                    int rem = pp - k;
                    if (rem < 4) x &= (1u << (rem * 8)) - 1u;
                    mm += __popc(nz_bytes(x));
                    k += 4;
                    if (mm > limit || k >= pp) active = false;
                }
            }
            unsigned am = __ballot_sync(FULL_MASK, valid && mm <= limit);
            if (am) {
                const int wl = __ffs(am) - 1;
                const int wo = base + wl;
                const int wol = __shfl_sync(FULL_MASK, ol, wl);
                int diff = __shfl_sync(FULL_MASK, mm, wl);
                if (wol > 50)                                              /* :41-43 full recount */
                    diff = dir == 0 ? dev_count_mismatch_coop(s1 + wo, rc, wol) : dev_count_mismatch_coop(s1, rc + wo, wol);
                ov.overlapped = 1;
                ov.offset = (int16_t)(dir == 0 ? wo : -wo);
                ov.overlap_len = (int16_t)wol;
                ov.diff = (int16_t)diff;
                return ov;
            }
        }
    }
    return ov;
}

/* ------------------------------------------------------------------------------------------------
 * BaseCorrector::correctByOverlapAnalysis  (basecorrector.cpp:21-83).
 * Bases/quals are overwritten in the shared-memory tile AND in the HBM rows (g1/g2 = global row
 * pointers already advanced by front); each overwritten base is also appended to the patch list.
 * ------------------------------------------------------------------------------------------------ */
struct PatchSink { fp_patch* patches; unsigned int cap; unsigned int* count; };

__device__ __noinline__ void dev_correct(WRead& r1, WRead& r2, const fp_ov_result ov, uint8_t* g1s, uint8_t* g1q,
                                         uint8_t* g2s, uint8_t* g2q, unsigned int pair_index, const PatchSink& sink,
                                         BlockCounters* bc, bool& c1, bool& c2) {
    c1 = c2 = false;
    if (ov.diff == 0 || !ov.overlapped) return;                           /* :23-24 */
    const int lane = lane_id();
    const int ol = ov.overlap_len;
    const int start1 = max(0, (int)ov.offset);
    const int start2 = r2.len - max(0, -(int)ov.offset) - 1;
    const signed char GOOD = 33 + 30, BAD = 33 + 14;                       /* :35-36 */
    int corrected = 0;
    bool my1 = false, my2 = false;
    for (int i = lane; i < ol; i += 32) {
        const int p1 = start1 + i, p2 = start2 - i;
        const uint8_t b1 = r1.seq[p1], b2 = r2.seq[p2];
        if (b1 != dev_complement(b2)) {
            const signed char q1 = (signed char)r1.qual[p1], q2 = (signed char)r2.qual[p2];
            if (q1 >= GOOD && q2 <= BAD) {                                 /* use R1 :42-50 */
                const uint8_t nb = dev_complement(b1);
                r2.seq[p2] = nb; r2.qual[p2] = (uint8_t)q1;
                g2s[p2] = nb; g2q[p2] = (uint8_t)q1;
                corrected++; my2 = true;
                /* addCorrection(seq2[p2] (already overwritten), complement(seq1[p1])) -> diagonal entry */
                atomicAdd(&bc->fr[FP_FR_CORRECTION + (nb & 7) * 8 + (nb & 7)], 1u);
                if (sink.count) {
                    unsigned int slot = atomicAdd(sink.count, 1u);
                    if (slot < sink.cap) { fp_patch pt; pt.pair = pair_index; pt.pos = (uint16_t)(r2.front + p2); pt.which = 1; pt.base = nb; pt.qual = (uint8_t)q1; pt._pad[0] = pt._pad[1] = pt._pad[2] = 0; sink.patches[slot] = pt; }
                }
            } else if (q2 >= GOOD && q1 <= BAD) {                          /* use R2 :51-59 */
                const uint8_t nb = dev_complement(b2);
                r1.seq[p1] = nb; r1.qual[p1] = (uint8_t)q2;
                g1s[p1] = nb; g1q[p1] = (uint8_t)q2;
                corrected++; my1 = true;
                atomicAdd(&bc->fr[FP_FR_CORRECTION + (nb & 7) * 8 + (nb & 7)], 1u);
                if (sink.count) {
                    unsigned int slot = atomicAdd(sink.count, 1u);
                    if (slot < sink.cap) { fp_patch pt; pt.pair = pair_index; pt.pos = (uint16_t)(r1.front + p1); pt.which = 0; pt.base = nb; pt.qual = (uint8_t)q2; pt._pad[0] = pt._pad[1] = pt._pad[2] = 0; sink.patches[slot] = pt; }
                }
            }
        }
    }
    c1 = __any_sync(FULL_MASK, my1);
    c2 = __any_sync(FULL_MASK, my2);
    corrected = warp_sum(corrected);
    if (corrected > 0 && lane == 0) atomicAdd(&bc->fr[FP_FR_CORRECTED_READS], (c1 && c2) ? 2u : 1u);   /* :75-80 */
    __syncwarp();
}

/* ------------------------------------------------------------------------------------------------
 * AdapterTrimmer::trimBySequence  (adaptertrimmer.cpp:64-157) with Matcher::matchWithOneInsertion
 * (matcher.cpp:10-54) in closed form.
 *
 * matchWithOneInsertion(ins, norm, c, limit) == exists i in [1,c-1] with L[i-1] + R[i] <= limit where
 * L/R are the true prefix / suffix mismatch counts (the early breaks and the diffLimit+1 sentinel never
 * change the verdict, tests/test_closed_forms.py).  With D1[j] = ins[j]!=norm[j], D2[j] = ins[j+1]!=norm[j],
 * P1/P2 their exclusive prefix sums:  min_{1<=i<=c-1}(P1[i]-P2[i]) + P2[c] <= limit.
 * The insertion / deletion scans (:105-135) align the adapter to the read START for every `pos`
 * (rdata is not advanced, :110) -- only cmplen varies -- so one pass over j < alen decides all `pos`.
 * A = per-warp int scratch (>= 2*(FP_MAX_ADAPTER_LEN+2) ints).
 * ------------------------------------------------------------------------------------------------ */
__device__ __noinline__ int dev_gap_scan(const uint8_t* ins, int ins_n, const uint8_t* norm, int norm_n, int cmax, int cmin, int* A) {
    /* returns the largest c in [cmin, cmax] for which the one-insertion match holds, else -1.
       ins readable up to index cmax (ins_n > cmax), norm readable up to cmax-1. */
    const int lane = lane_id();
    if (cmax < cmin) return -1;
    int* M = A;                 /* M[c]  = min_{1<=i<=c-1} (P1[i]-P2[i]) for c>=2            */
    int* P2 = A + (FP_MAX_ADAPTER_LEN + 2);
    int c1 = 0, c2 = 0;         /* running prefix sums P1[base], P2[base] */
    int runmin = 1 << 20;       /* min over i in [1, base-1] */
    if (lane == 0) { P2[0] = 0; M[0] = runmin; M[1] = runmin; }
    for (int base = 0; base <= cmax; base += 32) {
        int j = base + lane;                       /* computes P1[j+1], P2[j+1] */
        int d1 = (j < cmax) ? (ins[j] != norm[j]) : 0;
        int d2 = (j < cmax) ? (ins[j + 1] != norm[j]) : 0;
        int s1 = d1, s2 = d2;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t1 = __shfl_up_sync(FULL_MASK, s1, o), t2 = __shfl_up_sync(FULL_MASK, s2, o);
            if (lane >= o) { s1 += t1; s2 += t2; }
        }
        int p1 = c1 + s1, p2 = c2 + s2;            /* P1[j+1], P2[j+1] */
        /* m[i] for i = j+1 (valid for i>=1): P1[i]-P2[i]; prefix-min over i in [1, j+1] */
        int m = (j + 1 <= cmax) ? (p1 - p2) : (1 << 20);
        int pm = m;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL_MASK, pm, o); if (lane >= o) pm = min(pm, t); }
        pm = min(pm, runmin);                      /* min over i in [1, j+1] */
        if (j + 1 <= cmax) { P2[j + 1] = p2; M[j + 2] = pm; }   /* M[c] with c-1 = j+1 */
        c1 += __shfl_sync(FULL_MASK, s1, 31); c2 += __shfl_sync(FULL_MASK, s2, 31);
        runmin = __shfl_sync(FULL_MASK, pm, 31);
    }
    __syncwarp();
    /* largest c in [cmin,cmax] with M[c] + P2[c] <= c/8 - 1 */
    int best = -1;
    for (int top = cmax; top >= cmin; top -= 32) {
        int c = top - lane;
        bool ok = (c >= cmin) && (c >= 2) && (M[c] + P2[c] <= c / 8 - 1);
        unsigned m = __ballot_sync(FULL_MASK, ok);
        if (m) { best = top - (__ffs(m) - 1); break; }
    }
    __syncwarp();
    return best;
}

__device__ __noinline__ bool dev_trim_by_sequence(WRead& r, const uint8_t* adata, int alen, int matchReq, int* A,
                                                  int& posOut, int& basesOut, BlockCounters* bc, int aidx, const Planes* RP) {
    const int lane = lane_id();
    const int rlen = r.len;
    const uint8_t* rdata = r.seq;
    if (alen < matchReq) return false;                                    /* :73-74 */
    int start = 0;
    if (alen >= 16) start = -4; else if (alen >= 12) start = -3; else if (alen >= 8) start = -2;
    bool found = false;
    int pos = 0;
    /* scan 1, negative positions (:87-100): adapter[-pos ..) vs read[0 ..), length min(rlen, alen+pos) */
    for (int p = start; p < 0 && p < rlen - matchReq; p++) {
        int cmplen = min(rlen - p, alen);
        int allowed = cmplen / 8;
        int so = -p;
        int mism = dev_count_mismatch_coop(adata + so, rdata, cmplen - so);
        if (mism <= allowed) { found = true; pos = p; break; }
    }
    /* scan 1, pos >= 0 on bit planes (clean read, clean adapter): lanes over pos, 32 bases per popc */
    if (!found && RP != nullptr && c_p.adapter_clean[aidx]) {
        const int npos = rlen - matchReq;
        const uint32_t* alo = c_p.adapter_planes + aidx * 24; const uint32_t* ahi = alo + 8; const uint32_t* ann = alo + 16;
        const int nw = (alen + 31) >> 5;
        for (int base = 0; base < npos; base += 32) {
            const int p = base + lane;
            const bool valid = p < npos;
            int allowed = -1, mm = 0;
            if (valid) {
                const int cmplen = min(rlen - p, alen);
                allowed = cmplen / 8;
                for (int k = 0; k < nw; k++) {
                    if (32 * k >= cmplen) break;
                    const int bit = r.front + p + 32 * k;
                    const uint32_t x = (plane_bits(RP->lo, bit) ^ __ldg(alo + k)) | (plane_bits(RP->hi, bit) ^ __ldg(ahi + k)) |
                                       (plane_bits(RP->nn, bit) ^ __ldg(ann + k));
                    mm += __popc(x & low_mask(cmplen - 32 * k));
                }
            }
            const unsigned am = __ballot_sync(FULL_MASK, valid && mm <= allowed);
            if (am) { found = true; pos = base + __ffs(am) - 1; break; }
        }
    } else if (!found) {
        const int npos = rlen - matchReq;                                  /* pos in [0, npos) */
        const uint32_t* aw = reinterpret_cast<const uint32_t*>(adata);
        for (int base = 0; base < npos; base += 32) {
            const int p = base + lane;
            const bool valid = p < npos;
            int cmplen = 0, allowed = 0, mm = 0;
            if (valid) { cmplen = min(rlen - p, alen); allowed = cmplen / 8; }
            bool active = valid;
            int k = 0;
            while (__any_sync(FULL_MASK, active)) {
                if (active) {
                    uint32_t a = __ldg(aw + (k >> 2));
                    uint32_t b = ld_u32_unaligned(rdata + p + k);
                    uint32_t x = a ^ b;
                    int rem = cmplen - k;
                    if (rem < 4) x &= (1u << (rem * 8)) - 1u;
                    mm += __popc(nz_bytes(x));
                    k += 4;
                    if (mm > allowed || k >= cmplen) active = false;
                }
            }
            unsigned am = __ballot_sync(FULL_MASK, valid && mm <= allowed);
            if (am) { found = true; pos = base + __ffs(am) - 1; break; }
        }
    }
    /* scan 2 (:105-118): insertion in the read. pos in [0, rlen-matchReq-1), cmplen = min(rlen-pos-1, alen) */
    if (!found && rlen - matchReq - 1 > 0) {
        int cmax = min(rlen - 1, alen);
        int cmin = matchReq + 1;                                          /* pos = rlen-matchReq-2 */
        int c = dev_gap_scan(rdata, rlen, adata, alen, cmax, cmin, A);
        if (c >= 0) { found = true; pos = (c == cmax) ? 0 : rlen - 1 - c; }
    }
    /* scan 3 (:122-135): deletion in the read. pos in [0, rlen-matchReq), cmplen = min(rlen-pos, alen-1) */
    if (!found && rlen - matchReq > 0) {
        int cmax = min(rlen, alen - 1);
        int cmin = matchReq + 1;                                          /* pos = rlen-matchReq-1 */
        int c = dev_gap_scan(adata, alen, rdata, rlen, cmax, cmin, A);
        if (c >= 0) { found = true; pos = (c == cmax) ? 0 : rlen - c; }
    }
    if (found) {                                                          /* :137-154 */
        int abases;
        if (pos < 0) { abases = alen + pos; r.len = 0; }
        else { abases = rlen - pos; if (pos <= r.len) r.len = pos; }
        if (abases > 0 && lane == 0) atomicAdd(&bc->fr[FP_FR_ADAPTER_BASES], (unsigned)abases);
        posOut = pos; basesOut += max(abases, 0);
        return true;
    }
    return false;
}

__device__ __forceinline__ bool dev_trim_by_multi(WRead& r, int* A, int& posOut, int& basesOut, BlockCounters* bc, const Planes* RP) {
    bool trimmed = false;                                                 /* adaptertrimmer.cpp:48-62 */
    for (int i = 0; i < c_p.n_fasta; i++)
        trimmed |= dev_trim_by_sequence(r, c_p.adapters + c_p.fasta_off[i], c_p.fasta_len[i], c_p.fasta_match_req, A, posOut, basesOut, bc, 2 + i, RP);
    return trimmed;
}

/* ------------------------------------------------------------------------------------------------
 * Filter::passFilter  (filter.cpp:15-57) + countQualityMetrics (simd.cpp:281-295) +
 * passLowComplexityFilter (filter.cpp:59-66)
 * ------------------------------------------------------------------------------------------------ */
__device__ __noinline__ int dev_pass_filter(const WRead& r, const int16_t* lut) {
    if (r.null || r.len == 0) return FP_FAIL_LENGTH;
    const int lane = lane_id();
    const int rlen = r.len;
    int lowq = 0, nb = 0, tq = 0, adj = 0;
    const bool need_metrics = c_p.qual_filter || c_p.length_filter;
    const uint8_t qq = (uint8_t)c_p.qualified_qual;
    for (int i = lane; i < rlen; i += 32) {
        uint8_t q = r.qual[i], b = r.seq[i];
        tq += (int)q - 33;
        lowq += (q < qq);
        nb += (b == 'N');
        if (i + 1 < rlen) adj += (b != r.seq[i + 1]);
    }
    if (need_metrics) { lowq = warp_sum(lowq); nb = warp_sum(nb); tq = warp_sum(tq); }
    else { lowq = nb = tq = 0; }
    if (c_p.qual_filter) {
        if (lowq > (int)lut[(c_p.stride + 2) + rlen]) return FP_FAIL_QUALITY;
        else if (c_p.avg_qual_req > 0 && (tq / rlen) < c_p.avg_qual_req) return FP_FAIL_QUALITY;
        else if (nb > c_p.n_base_limit) return FP_FAIL_N_BASE;
    }
    if (c_p.length_filter) {
        if (rlen < c_p.length_required) return FP_FAIL_LENGTH;
        if (c_p.length_limit > 0 && rlen > c_p.length_limit) return FP_FAIL_TOO_LONG;
    }
    if (c_p.complexity_filter) {
        if (rlen <= 1) return FP_FAIL_COMPLEXITY;
        adj = warp_sum(adj);
        if (adj < (int)lut[2 * (c_p.stride + 2) + rlen]) return FP_FAIL_COMPLEXITY;
    }
    return FP_PASS_FILTER;
}

/* per-read outcome kept in shared memory between phase 2 and phase 3 */
struct ReadOutcome { uint16_t front, len; };    /* len == 0xFFFF: not counted by the post stats */

__device__ __forceinline__ fp_read_result make_result(const WRead& r, int verdict, int pv, int flags, int apos, int abases, int pbase, int plen) {
    fp_read_result o;
    if (r.null) { o.front = 0; o.len = 0; flags |= FP_F_DROPPED; }
    else { o.front = (uint16_t)r.front; o.len = (uint16_t)r.len; }
    o.verdict = (uint8_t)verdict; o.flags = (uint8_t)flags; o.adapter_pos = (int16_t)apos; o.adapter_len = (uint16_t)abases;
    o.polyx_base = (uint8_t)pbase; o.pair_verdict = (uint8_t)pv; o.polyx_len = (uint16_t)plen; o.reserved = 0;
    return o;
}

/* ------------------------------------------------------------------------------------------------
 * Column-pass statistics (Stats::statRead per-base part, stats.cpp:204-268).
 *
 * Thread = (side, word column w, row group g).  It owns cycles 4w..4w+3 of that side for its whole
 * life; acc[cyc][bin][kind] are 32-bit registers: kind 0 = count, 1 = count(q>='5'), 2 = count(q>='?'),
 * 3 = sum of raw quality chars (qualsum = kind3 - 33*kind0).  Bins: A C T N G (base&7 = 1 3 4 6 7);
 * any other byte goes through a slow global-atomic path.
 * ------------------------------------------------------------------------------------------------ */
#define NB 5
struct ColAcc { unsigned int v[4][NB][4]; };

__device__ __forceinline__ void transpose4x4(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3,
                                             uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3) {
    uint32_t t0 = __byte_perm(r0, r1, 0x5140);   /* r0b0 r1b0 r0b1 r1b1 */
    uint32_t t1 = __byte_perm(r2, r3, 0x5140);
    uint32_t t2 = __byte_perm(r0, r1, 0x7362);   /* r0b2 r1b2 r0b3 r1b3 */
    uint32_t t3 = __byte_perm(r2, r3, 0x7362);
    c0 = __byte_perm(t0, t1, 0x5410);
    c1 = __byte_perm(t0, t1, 0x7632);
    c2 = __byte_perm(t2, t3, 0x5410);
    c3 = __byte_perm(t2, t3, 0x7632);
}

/* keep bytes [lo, hi) of the word covering positions [4w, 4w+4) */
__device__ __forceinline__ uint32_t window_mask(int w4, int lo, int hi) {
    int a = max(lo - w4, 0), b = min(hi - w4, 4);
    if (b <= a) return 0u;
    uint32_t m = (b >= 4) ? 0xffffffffu : ((1u << (8 * b)) - 1u);
    if (a > 0) m &= ~((1u << (8 * a)) - 1u);
    return m;
}

__device__ __forceinline__ void acc_cycle(unsigned int (&a)[NB][4], uint32_t s, uint32_t q) {
    /* s: bases of 4 rows at one cycle (0 = masked out), q: their quality chars */
    const uint32_t K = 0x01010101u;
    uint32_t p0 = s & K, p1 = (s >> 1) & K, p2 = (s >> 2) & K;
    uint32_t mA = p0 & ~p1 & ~p2;          /* 001 */
    uint32_t mC = p0 & p1 & ~p2;           /* 011 */
    uint32_t mT = ~p0 & ~p1 & p2;          /* 100 */
    uint32_t mN = ~p0 & p1 & p2;           /* 110 */
    uint32_t mG = p0 & p1 & p2;            /* 111 */
    /* q >= '5' (53) and q >= '?' (63): q in [0,127] -> (q + 128 - thr) bit 7 */
    uint32_t q7 = q & 0x7F7F7F7Fu;
    uint32_t t20 = ((q7 + 0x4B4B4B4Bu) >> 7) & K;   /* 128-53 = 75 = 0x4B */
    uint32_t t30 = ((q7 + 0x41414141u) >> 7) & K;   /* 128-63 = 65 = 0x41 */
    uint32_t m[NB] = {mA, mC, mT, mN, mG};
    #pragma unroll
    for (int b = 0; b < NB; b++) {
        a[b][0] = __dp4a(m[b], K, a[b][0]);
        a[b][1] = __dp4a(m[b], t20, a[b][1]);
        a[b][2] = __dp4a(m[b], t30, a[b][2]);
        a[b][3] = __dp4a(m[b], q, a[b][3]);
    }
}

/* slow path for one byte that is not A/C/G/T/N (or a quality >= 128): global atomics */
__device__ __noinline__ void slow_cycle_byte(unsigned long long* G, int stats, int cycle, uint8_t base, uint8_t q) {
    const fp_counter_layout& L = c_p.L;
    int b = base & 7;
    if (cycle >= L.cycles) return;
    if (q >= '?') { red_add64(&G[fp_off_cycle(&L, stats, 0 * 8 + b, cycle)], 1ull); red_add64(&G[fp_off_cycle(&L, stats, 1 * 8 + b, cycle)], 1ull); }
    else if (q >= '5') red_add64(&G[fp_off_cycle(&L, stats, 1 * 8 + b, cycle)], 1ull);
    red_add64(&G[fp_off_cycle(&L, stats, 2 * 8 + b, cycle)], 1ull);
    red_add64(&G[fp_off_cycle(&L, stats, 3 * 8 + b, cycle)], (unsigned long long)(long long)((int)q - 33));
}

/* 2-bit value of a base for the 5-mer code (stats.cpp:293-318): A0 T1 C2 G3, -1 otherwise */
__device__ __forceinline__ int dev_base2val(uint8_t b) {
    return b == 'A' ? 0 : b == 'T' ? 1 : b == 'C' ? 2 : b == 'G' ? 3 : -1;
}

/* ------------------------------------------------------------------------------------------------
 * Exact per-position statistics engine (warp per row, global atomics).  Adds `sign` times the
 * contribution Stats::statRead (stats.cpp:204-268) makes for positions i in [lo,hi) of a read whose
 * first counted base is at row index ctx0 (cycle = i - ctx0; the 5-mer ending at i counts iff
 * i-4 >= ctx0 and all five bases are in ACGT).  Used for
 *   - rows holding bytes outside {A,C,G,T,N} (excluded from the dense column pass),
 *   - the post-filter Stats as a DELTA against the pre-filter Stats: post = pre - (bases removed by
 *     trimming / failed reads) + (front-shifted or corrected reads re-added).
 * ------------------------------------------------------------------------------------------------ */
__device__ __noinline__ void dev_stat_positions(unsigned long long* G, int stats, const uint8_t* seq, const uint8_t* qual,
                                               int ctx0, int lo, int hi, int sign) {
    const fp_counter_layout& L = c_p.L;
    const unsigned long long one = (unsigned long long)(long long)sign;
    for (int i = lo + lane_id(); i < hi; i += 32) {
        const uint8_t base = seq[i], q = qual[i];
        const int b = base & 7;
        const int cyc = i - ctx0;
        if (q < FP_QUAL_BINS) red_add64(&G[fp_off_qualhist(&L, stats, q)], one);
        if (cyc < L.cycles) {
            const unsigned long long qv = (unsigned long long)((long long)sign * ((int)q - 33));
            if (q >= '?') { red_add64(&G[fp_off_cycle(&L, stats, 0 * 8 + b, cyc)], one); red_add64(&G[fp_off_cycle(&L, stats, 1 * 8 + b, cyc)], one); }
            else if (q >= '5') red_add64(&G[fp_off_cycle(&L, stats, 1 * 8 + b, cyc)], one);
            red_add64(&G[fp_off_cycle(&L, stats, 2 * 8 + b, cyc)], one);
            red_add64(&G[fp_off_cycle(&L, stats, 3 * 8 + b, cyc)], qv);
        }
        if (i - 4 >= ctx0) {
            int code = 0; bool ok = true;
            #pragma unroll
            for (int k = 0; k < 5; k++) { int v = dev_base2val(seq[i - 4 + k]); ok = ok && (v >= 0); code = (code << 2) | (v & 3); }
            if (ok) red_add64(&G[fp_off_kmer(&L, stats, code)], one);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Block-private form of the same engine for CLEAN rows (bases in {A,C,G,T,N}): signed 32-bit shared-memory
 * accumulators instead of global atomics (the global block would serialise on its few hot addresses).
 *   D.cyc [side][cycle][bin A,C,T,N,G][kind count,q20,q30,qualsum]   D.kmer [side][1024]   D.qh [side][128]
 * Flushed once per CTA into the POST stats of that side.
 * ------------------------------------------------------------------------------------------------ */
struct DeltaAcc { int* cyc; int* kmer; int* qh; int cycles; };

__device__ __noinline__ void dev_stat_positions_smem(const DeltaAcc D, int side, const uint8_t* seq, const uint8_t* qual,
                                                    int ctx0, int lo, int hi, int sign) {
    const int lane = lane_id();
    int* cy = D.cyc + side * D.cycles * 20;
    int* km = D.kmer + side * FP_KMER_BINS;
    int* qh = D.qh + side * FP_QUAL_BINS;
    for (int base = lo; base < hi; base += 32) {
        const int i = base + lane;
        const bool valid = i < hi;
        uint8_t b = 0, q = 0;
        if (valid) { b = seq[i]; q = qual[i]; }
        /* quality histogram: aggregate equal values inside the warp first (few distinct qualities) */
        const unsigned peers = __match_any_sync(FULL_MASK, valid ? (int)q : -1);
        if (valid && lane == __ffs(peers) - 1) atomicAdd(&qh[q], sign * __popc(peers));
        if (valid) {
            const int cyc = i - ctx0;
            if (cyc < D.cycles) {
                const int bin = (0x43F21F0Fu >> (4 * (b & 7))) & 0xF;       /* base&7: A1 C3 T4 N6 G7 -> 0..4 */
                int* c4 = cy + (cyc * 5 + bin) * 4;
                atomicAdd(&c4[0], sign);
                if (q >= '5') atomicAdd(&c4[1], sign);
                if (q >= '?') atomicAdd(&c4[2], sign);
                atomicAdd(&c4[3], sign * ((int)q - 33));
            }
            if (i - 4 >= ctx0) {
                /* 5-mer ending at i: two unaligned 32-bit loads cover bases i-4..i */
                const uint32_t w0 = ld_u32_unaligned(seq + i - 4);             /* bases i-4..i-1 */
                const uint32_t K = 0x01010101u;
                const uint32_t c0 = w0 & K, c1 = (w0 >> 1) & K, c2 = (w0 >> 2) & K;
                const uint32_t ok4 = (c0 & ~c1 & ~c2) | (c0 & c1) | (~c0 & ~c1 & c2);       /* A C G T by base&7 */
                const uint32_t v4 = (w0 & 0x02020202u) | c2;
                const int vb = ((b >> 1) & 1) * 2 + ((b >> 2) & 1);
                const bool okb = (b == 'A') | (b == 'C') | (b == 'G') | (b == 'T');
                if (ok4 == K && okb) atomicAdd(&km[((((v4 * 0x40100401u) >> 24) << 2) | vb) & 0x3FF], sign);
            }
        }
    }
}

/* post-filter delta: block-private for clean rows, exact global path otherwise */
__device__ __forceinline__ void post_delta(bool clean, const DeltaAcc& D, unsigned long long* G, int side, const uint8_t* seq, const uint8_t* qual,
                                           int ctx0, int lo, int hi, int sign) {
    if (hi <= lo) return;
    if (clean) dev_stat_positions_smem(D, side, seq, qual, ctx0, lo, hi, sign);
    else dev_stat_positions(G, side * 2 + 1, seq, qual, ctx0, lo, hi, sign);
}

/* ------------------------------------------------------------------------------------------------
 * Shared-memory layout of one CTA (offsets computed on the host, fp_api.cu)
 * ------------------------------------------------------------------------------------------------ */
struct fp_smem_layout {
    int off_mbar, off_tile, tile_array_bytes, off_len, off_clean, off_rc, rc_bytes, off_scratch, scratch_ints,
        off_kmer, off_qhist, off_bc, off_rl, off_next, off_planes, off_rcplanes, off_lut, off_delta, off_queue, plane_words, plane_stride, total;
};

struct fp_launch_args {
    fp_batch b;                      /* device pointers */
    fp_read_result* out1;
    fp_read_result* out2;
    fp_ov_result* ov;
    PatchSink sink;
    unsigned long long* counters;    /* global int64 block (two's complement adds) */
    long long n_tiles;
    fp_smem_layout sl;
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
/* TMA bulk copy global -> shared (1-D), completion signalled on the mbarrier */
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

/* ------------------------------------------------------------------------------------------------
 * The fused kernel.
 * ------------------------------------------------------------------------------------------------ */
template <bool PAIRED>
__global__ void __launch_bounds__(FP_THREADS, 2) fp_chain_kernel(const fp_launch_args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int SIDES = PAIRED ? 2 : 1;
    const fp_smem_layout& sl = a.sl;
    const int S = c_p.stride, T = c_p.tile;
    const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    unsigned long long* G = a.counters;
    const fp_counter_layout& L = c_p.L;

    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + sl.off_mbar);
    uint8_t* tile_seq[2]; uint8_t* tile_qual[2];
    tile_seq[0] = smem + sl.off_tile;
    tile_qual[0] = tile_seq[0] + sl.tile_array_bytes;
    tile_seq[1] = tile_qual[0] + sl.tile_array_bytes;
    tile_qual[1] = tile_seq[1] + sl.tile_array_bytes;
    uint16_t* s_len = reinterpret_cast<uint16_t*>(smem + sl.off_len);       /* [SIDES][T] */
    uint8_t* s_clean = smem + sl.off_clean;                                 /* [SIDES][T] */
    uint8_t* my_rc = smem + sl.off_rc + warp * sl.rc_bytes;
    int* my_scratch = reinterpret_cast<int*>(smem + sl.off_scratch) + warp * sl.scratch_ints;
    unsigned int* s_kmer = reinterpret_cast<unsigned int*>(smem + sl.off_kmer);    /* [SIDES][1024] */
    unsigned int* s_qhist = reinterpret_cast<unsigned int*>(smem + sl.off_qhist);  /* [SIDES][128]  */
    BlockCounters* bc = reinterpret_cast<BlockCounters*>(smem + sl.off_bc);
    DeltaAcc D;
    D.cycles = S;
    D.cyc = reinterpret_cast<int*>(smem + sl.off_delta);                                   /* [SIDES][S][5][4] */
    D.kmer = D.cyc + SIDES * S * 20;                                                       /* [SIDES][1024]    */
    D.qh = D.kmer + SIDES * FP_KMER_BINS;                                                  /* [SIDES][128]     */
    int16_t* s_lut = reinterpret_cast<int16_t*>(smem + sl.off_lut);                        /* [3][S+2]: ovlimit, lowq, mindiff */
    int* s_next = reinterpret_cast<int*>(smem + sl.off_next);                              /* dynamic row claim of phase 2 */
    const int PW = sl.plane_words;
    uint32_t* tile_planes = reinterpret_cast<uint32_t*>(smem + sl.off_planes);             /* [SIDES][T][4][PW] */
    uint32_t* my_rcp = reinterpret_cast<uint32_t*>(smem + sl.off_rcplanes) + warp * (3 * PW);
    Planes PRC = {my_rcp, my_rcp + PW, my_rcp + 2 * PW, nullptr};

    /* zero block-level accumulators */
    for (int i = tid; i < SIDES * FP_KMER_BINS; i += FP_THREADS) s_kmer[i] = 0;
    for (int i = tid; i < SIDES * FP_QUAL_BINS; i += FP_THREADS) s_qhist[i] = 0;
    for (int i = tid; i < SIDES * (S * 20 + FP_KMER_BINS + FP_QUAL_BINS); i += FP_THREADS) D.cyc[i] = 0;
    for (int i = tid; i < (int)(sizeof(BlockCounters) / 4); i += FP_THREADS) reinterpret_cast<unsigned int*>(bc)[i] = 0;
    for (int i = tid; i < S + 2; i += FP_THREADS) { s_lut[i] = c_p.lut_ovlimit[i]; s_lut[(S + 2) + i] = c_p.lut_lowq[i]; s_lut[2 * (S + 2) + i] = c_p.lut_mindiff[i]; }
    if (tid == 0) { mbar_init(mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

    /* column-pass ownership */
    const int WPR = S >> 2;                       /* words per row */
    const int ncols = SIDES * WPR;
    const int ngroups = FP_THREADS / ncols > 0 ? FP_THREADS / ncols : 1;
    const bool col_active = tid < ngroups * ncols;      /* host guarantees ncols <= FP_THREADS */
    const int my_col = tid % ncols, my_group = tid / ncols;
    const int my_side = my_col / WPR, my_w = my_col % WPR;
    ColAcc acc;
    #pragma unroll
    for (int c = 0; c < 4; c++)
        #pragma unroll
        for (int b = 0; b < NB; b++)
            #pragma unroll
            for (int k = 0; k < 4; k++) acc.v[c][b][k] = 0;

    unsigned long long rl[8] = {0, 0, 0, 0, 0, 0, 0, 0};   /* per-warp (uniform) reads / lengthSum of pre1 post1 pre2 post2 */
    __syncthreads();
    uint32_t parity = 0;

    for (long long tix = blockIdx.x; tix < a.n_tiles; tix += gridDim.x) {
        const long long row0 = tix * T;
        const int rows = (int)min((long long)T, a.b.n - row0);
        /* ---------------- phase 0: TMA bulk loads ---------------- */
        if (tid == 0) {
            *s_next = 0;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t bytes = (uint32_t)rows * (uint32_t)S;
            mbar_expect_tx(mbar, bytes * 2 * SIDES);
            tma_bulk_g2s(tile_seq[0], a.b.seq1 + row0 * S, bytes, mbar);
            tma_bulk_g2s(tile_qual[0], a.b.qual1 + row0 * S, bytes, mbar);
            if (PAIRED) {
                tma_bulk_g2s(tile_seq[1], a.b.seq2 + row0 * S, bytes, mbar);
                tma_bulk_g2s(tile_qual[1], a.b.qual2 + row0 * S, bytes, mbar);
            }
        }
        for (int i = tid; i < SIDES * T; i += FP_THREADS) {
            int sd = i / T, r = i % T;
            uint16_t ln = 0;
            if (r < rows) { ln = (sd == 0 ? a.b.len1 : a.b.len2)[row0 + r]; if (ln > S) ln = (uint16_t)S; }
            s_len[i] = ln;
        }
        mbar_wait(mbar, parity);
        parity ^= 1;
        __syncthreads();

        /* ---------------- phase 0.5: bit planes of every row + validation (clean = only A,C,G,T,N, quals < 128) ---------------- */
        for (int i = tid; i < SIDES * T; i += FP_THREADS) s_clean[i] = 1;
        __syncthreads();
        {
            const int nwords = (S + 31) >> 5;                     /* plane words holding bases; the rest of PW is zero padding */
            const uint32_t qq4 = (uint32_t)(c_p.qualified_qual & 0x7F) * 0x01010101u;
            for (int it = tid; it < SIDES * T * PW; it += FP_THREADS) {
                const int j = it % PW, rr = (it / PW) % T, sd = it / (PW * T);
                uint32_t lo = 0, hi = 0, nn = 0, lq = 0;
                if (j < nwords && rr < rows) {
                    const int ln = s_len[sd * T + rr];
                    const int n = ln - 32 * j;
                    if (n > 0) {
                        const uint4* sp = reinterpret_cast<const uint4*>(tile_seq[sd] + rr * S + 32 * j);
                        const uint4* qp = reinterpret_cast<const uint4*>(tile_qual[sd] + rr * S + 32 * j);
                        const uint4 s0 = sp[0], s1 = sp[1], q0 = qp[0], q1 = qp[1];
                        const uint32_t x[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                        const uint32_t q[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                        if (!plane_word_from_bytes(x, q, n, qq4, lo, hi, nn, lq)) s_clean[sd * T + rr] = 0;
                    }
                }
                uint32_t* pr = tile_planes + ((sd * T + rr) * 4) * PW + j;
                pr[0] = lo; pr[PW] = hi; pr[2 * PW] = nn; pr[3 * PW] = lq;
            }
        }
        __syncthreads();

        /* ---------------- phase 1: dense column pass (pre-filter stats of clean rows) ---------------- */
        if (col_active) {
            const uint8_t* ts = tile_seq[my_side]; const uint8_t* tq = tile_qual[my_side];
            const uint16_t* lens = s_len + my_side * T; const uint8_t* cl = s_clean + my_side * T;
            unsigned int* kh = s_kmer + my_side * FP_KMER_BINS; unsigned int* qh = s_qhist + my_side * FP_QUAL_BINS;
            const int w4 = my_w * 4;
            for (int r0 = my_group * 4; r0 < rows; r0 += ngroups * 4) {
                uint32_t xs[4], xq[4];
                #pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int r = r0 + k;
                    int hi = (r < rows && cl[r]) ? lens[r] : 0;
                    uint32_t m = window_mask(w4, 0, hi);
                    uint32_t x = 0, q = 0;
                    if (m) {
                        x = *reinterpret_cast<const uint32_t*>(ts + r * S + w4) & m;
                        q = *reinterpret_cast<const uint32_t*>(tq + r * S + w4) & m;
                        /* quality histogram (stats.cpp:213) */
                        int nb = min(hi - w4, 4);
                        #pragma unroll
                        for (int j = 0; j < 4; j++) if (j < nb) atomicAdd(&qh[(q >> (8 * j)) & 0xFF], 1u);
                        /* 5-mers ending in this word (stats.cpp:228-266) */
                        if (my_w > 0) {
                            uint32_t xp = *reinterpret_cast<const uint32_t*>(ts + r * S + w4 - 4);
                            const uint32_t K = 0x01010101u;
                            /* class masks from base&7: valid = A(001) C(011) T(100) G(111) */
                            uint32_t c0 = x & K, c1 = (x >> 1) & K, c2 = (x >> 2) & K;
                            uint32_t okc = (c0 & ~c1 & ~c2) | (c0 & c1) | (~c0 & ~c1 & c2);
                            uint32_t p0 = xp & K, p1 = (xp >> 1) & K, p2 = (xp >> 2) & K;
                            uint32_t okp = (p0 & ~p1 & ~p2) | (p0 & p1) | (~p0 & ~p1 & p2);
                            uint32_t vc = (x & 0x02020202u) | c2, vp = (xp & 0x02020202u) | p2;     /* val = bit1*2 + bit2 */
                            uint32_t s16 = (((vp * 0x40100401u) >> 24) << 8) | ((vc * 0x40100401u) >> 24);
                            uint32_t ok8 = ((((okp * 0x08040201u) >> 24) & 0xF) << 4) | (((okc * 0x08040201u) >> 24) & 0xF);
                            #pragma unroll
                            for (int j = 0; j < 4; j++)
                                if (((ok8 >> (3 - j)) & 0x1F) == 0x1F) atomicAdd(&kh[(s16 >> (2 * (3 - j))) & 0x3FF], 1u);
                        }
                    }
                    xs[k] = x; xq[k] = q;
                }
                uint32_t cs0, cs1, cs2, cs3, cq0, cq1, cq2, cq3;
                transpose4x4(xs[0], xs[1], xs[2], xs[3], cs0, cs1, cs2, cs3);
                transpose4x4(xq[0], xq[1], xq[2], xq[3], cq0, cq1, cq2, cq3);
                acc_cycle(acc.v[0], cs0, cq0);
                acc_cycle(acc.v[1], cs1, cq1);
                acc_cycle(acc.v[2], cs2, cq2);
                acc_cycle(acc.v[3], cs3, cq3);
            }
        }
        __syncthreads();

        /* ---------------- phase 2: per-read operator chain, one warp per read / pair (dynamic claim) ---------------- */
        for (;;) {
            int r = 0;
            if (lane == 0) r = atomicAdd(s_next, 1);
            r = __shfl_sync(FULL_MASK, r, 0);
            if (r >= rows) break;
            const long long gi = row0 + r;
            if (!PAIRED) {
                /* SingleEndProcessor::processSingleEnd loop body  seprocessor.cpp:204-296 */
                uint8_t* rs = tile_seq[0] + r * S; uint8_t* rq = tile_qual[0] + r * S;
                const int len0 = s_len[r];
                const bool clean = s_clean[r];
                rl[0] += 1; rl[1] += len0;
                if (!clean) dev_stat_positions(G, FP_STATS_PRE1, rs, rq, 0, 0, len0, +1);
                WRead r1 = {rs, rq, len0, 0, false};
                int flags = 0, apos = 0, abases = 0, pbase = 255, plen = 0;
                dev_trim_and_cut(r1, c_p.trim_front1, c_p.trim_tail1, my_scratch);               /* :235 */
                if (!r1.null && c_p.polyg) { if (dev_trim_polyg(r1, c_p.polyg_min)) flags |= FP_F_POLYG_TRIMMED; }   /* :237-240 */
                const bool usep = clean && !r1.null;
                uint32_t* pr1 = tile_planes + (r * 4) * PW;
                Planes P1 = {pr1, pr1 + PW, pr1 + 2 * PW, pr1 + 3 * PW};
                bool dimer = false;
                if (!r1.null && c_p.adapter_enabled) {                                            /* :243-260 */
                    bool trimmed = false;
                    if (c_p.has_r1) trimmed = dev_trim_by_sequence(r1, c_p.adapters + c_p.adapter_r1_off, c_p.adapter_r1_len, 4, my_scratch, apos, abases, bc, 0, usep ? &P1 : nullptr);
                    if (c_p.n_fasta > 0) trimmed |= dev_trim_by_multi(r1, my_scratch, apos, abases, bc, usep ? &P1 : nullptr);
                    if (trimmed) { if (lane == 0) atomicAdd(&bc->fr[FP_FR_ADAPTER_READS], 1u); flags |= FP_F_ADAPTER_TRIMMED; }
                    if (trimmed && r1.len <= c_p.dimer_max_len) dimer = true;
                }
                if (!r1.null && c_p.polyx) {                                                      /* :263-266 */
                    if (dev_trim_polyx(r1, c_p.polyx_min, pbase, plen)) {
                        if (lane == 0) { atomicAdd(&bc->fr[FP_FR_POLYX_READS + pbase], 1u); atomicAdd(&bc->fr[FP_FR_POLYX_BASES + pbase], (unsigned)plen); }
                        flags |= FP_F_POLYX_TRIMMED;
                    }
                }
                if (!r1.null && c_p.max_len1 > 0 && c_p.max_len1 < r1.len) r1.len = c_p.max_len1;   /* :268-271 */
                int result = usep ? dev_pass_filter_planes(r1.qual, r1.len, r1.null, P1, r1.front, PW, s_lut) : dev_pass_filter(r1, s_lut);   /* :273 */
                if (dimer) { result = FP_FAIL_ADAPTER_DIMER; flags |= FP_F_ADAPTER_DIMER; }
                if (lane == 0) atomicAdd(&bc->fr[FP_FR_READSTATS + result], 1u);                   /* :278 */
                const bool counted = !r1.null && result == FP_PASS_FILTER;                        /* :281-286 */
                /* post stats as a delta against pre */
                if (counted) {
                    rl[2] += 1; rl[3] += r1.len;
                    if (r1.front == 0 && clean) post_delta(true, D, G, 0, rs, rq, 0, r1.len, len0, -1);
                    else {
                        if (clean) post_delta(true, D, G, 0, rs, rq, 0, 0, len0, -1);
                        post_delta(clean, D, G, 0, rs, rq, r1.front, r1.front, r1.front + r1.len, +1);
                    }
                } else if (clean) post_delta(true, D, G, 0, rs, rq, 0, 0, len0, -1);
                if (lane == 0) a.out1[gi] = make_result(r1, result, result, flags, apos, abases, pbase, plen);
            } else {
                /* PairEndProcessor::processPairEnd loop body  peprocessor.cpp:383-643 */
                uint8_t* rs1 = tile_seq[0] + r * S; uint8_t* rq1 = tile_qual[0] + r * S;
                uint8_t* rs2 = tile_seq[1] + r * S; uint8_t* rq2 = tile_qual[1] + r * S;
                const int l1 = s_len[r], l2 = s_len[T + r];
                const bool clean1 = s_clean[r], clean2 = s_clean[T + r];
                rl[0] += 1; rl[1] += l1; rl[4] += 1; rl[5] += l2;
                if (!clean1) dev_stat_positions(G, FP_STATS_PRE1, rs1, rq1, 0, 0, l1, +1);
                if (!clean2) dev_stat_positions(G, FP_STATS_PRE2, rs2, rq2, 0, 0, l2, +1);
                WRead r1 = {rs1, rq1, l1, 0, false}, r2 = {rs2, rq2, l2, 0, false};
                int flags1 = 0, flags2 = 0, apos1 = 0, apos2 = 0, ab1 = 0, ab2 = 0, pb1 = 255, pb2 = 255, pl1 = 0, pl2 = 0;
                dev_trim_and_cut(r1, c_p.trim_front1, c_p.trim_tail1, my_scratch);               /* :425-426 */
                dev_trim_and_cut(r2, c_p.trim_front2, c_p.trim_tail2, my_scratch);
                const bool both = !r1.null && !r2.null;
                if (both && c_p.polyg) {                                                          /* :428-431 */
                    if (dev_trim_polyg(r1, c_p.polyg_min)) flags1 |= FP_F_POLYG_TRIMMED;
                    if (dev_trim_polyg(r2, c_p.polyg_min)) flags2 |= FP_F_POLYG_TRIMMED;
                }
                const bool usep1 = clean1 && !r1.null, usep2 = clean2 && !r2.null;
                uint32_t* pr1 = tile_planes + (r * 4) * PW; uint32_t* pr2 = tile_planes + ((T + r) * 4) * PW;
                Planes P1 = {pr1, pr1 + PW, pr1 + 2 * PW, pr1 + 3 * PW}, P2 = {pr2, pr2 + PW, pr2 + 2 * PW, pr2 + 3 * PW};
                bool dimer = false;
                bool removed1 = false, removed2 = false;      /* whole read already subtracted from post (before correction) */
                fp_ov_result ov; ov.overlapped = 0; ov.has_gap = 0; ov.offset = 0; ov.overlap_len = 0; ov.diff = 0;
                if (both && (c_p.adapter_enabled || c_p.correction || c_p.thread0)) {             /* :438-441 */
                    ov = (usep1 && usep2) ? dev_analyze_planes(r1.len, r1.front, r2.len, r2.front, P1, P2, PRC, PW, s_lut) : dev_analyze(r1, r2, my_rc, s_lut);
                    if (c_p.thread0) {                                                            /* statInsertSize :449-452 / :497-504, :710-723 */
                        int isize = c_p.isize_max;
                        if (ov.overlapped) {
                            if (ov.offset > 0) isize = r1.len + r2.len - ov.overlap_len + r1.front + r2.front;
                            else isize = ov.overlap_len + r1.front + r2.front;
                        }
                        if (isize > c_p.isize_max) isize = c_p.isize_max;
                        if (lane == 0) {
                            if (c_p.isize_max < FP_MAX_ISIZE_SMEM) atomicAdd(&bc->isize[isize], 1u);
                            else red_add64(&G[L.off_isize + isize], 1ull);
                        }
                    }
                }
                if (both && (c_p.adapter_enabled || c_p.correction)) {                            /* :443 */
                    if (c_p.correction && ov.overlapped && ov.diff != 0) {                        /* :453-456 */
                        /* the post stats are kept as a delta against the ORIGINAL bases: take the two reads out
                           before any base is overwritten; they are re-added below if the pair passes */
                        if (clean1) post_delta(true, D, G, 0, rs1, rq1, 0, 0, l1, -1);
                        if (clean2) post_delta(true, D, G, 1, rs2, rq2, 0, 0, l2, -1);
                        removed1 = removed2 = true;
                        __syncwarp();
                        bool c1, c2;
                        dev_correct(r1, r2, ov, a.b.seq1 + gi * S + r1.front, a.b.qual1 + gi * S + r1.front,
                                    a.b.seq2 + gi * S + r2.front, a.b.qual2 + gi * S + r2.front, (unsigned int)gi, a.sink, bc, c1, c2);
                        if (c1) { flags1 |= FP_F_CORRECTED; if (usep1) dev_rebuild_planes(rs1, rq1, l1, PW, P1); }
                        if (c2) { flags2 |= FP_F_CORRECTED; if (usep2) dev_rebuild_planes(rs2, rq2, l2, PW, P2); }
                    }
                    if (c_p.adapter_enabled) {                                                    /* :457-485 */
                        bool trimmed = false;
                        if (ov.overlapped && ov.offset < 0) {                                     /* trimByOverlapAnalysis adaptertrimmer.cpp:17-46 */
                            const int ol = ov.overlap_len;
                            const int nl1 = min(r1.len, ol + r2.front), nl2 = min(r2.len, ol + r1.front);
                            const int a1 = r1.len - nl1, a2 = r2.len - nl2;
                            r1.len = nl1; r2.len = nl2;
                            if (lane == 0) atomicAdd(&bc->fr[FP_FR_ADAPTER_BASES], (unsigned)(a1 + a2));
                            ab1 += a1; ab2 += a2;
                            trimmed = true;
                        }
                        bool t1 = trimmed, t2 = trimmed;
                        if (!trimmed) {                                                           /* :461-466 */
                            if (c_p.has_r1) t1 = dev_trim_by_sequence(r1, c_p.adapters + c_p.adapter_r1_off, c_p.adapter_r1_len, 4, my_scratch, apos1, ab1, bc, 0, usep1 ? &P1 : nullptr);
                            if (c_p.has_r2) t2 = dev_trim_by_sequence(r2, c_p.adapters + c_p.adapter_r2_off, c_p.adapter_r2_len, 4, my_scratch, apos2, ab2, bc, 1, usep2 ? &P2 : nullptr);
                        }
                        if (c_p.n_fasta > 0) {                                                    /* :467-470 */
                            t1 |= dev_trim_by_multi(r1, my_scratch, apos1, ab1, bc, usep1 ? &P1 : nullptr);
                            t2 |= dev_trim_by_multi(r2, my_scratch, apos2, ab2, bc, usep2 ? &P2 : nullptr);
                        }
                        if (t1) { if (lane == 0) atomicAdd(&bc->fr[FP_FR_ADAPTER_READS], 1u); flags1 |= FP_F_ADAPTER_TRIMMED; }   /* :472-475 */
                        if (t2) { if (lane == 0) atomicAdd(&bc->fr[FP_FR_ADAPTER_READS], 1u); flags2 |= FP_F_ADAPTER_TRIMMED; }
                        if ((t1 || t2) && r1.len <= c_p.dimer_max_len && r2.len <= c_p.dimer_max_len) dimer = true;   /* :480-484 */
                    }
                }
                if (both && c_p.polyx) {                                                          /* :506-509 */
                    if (dev_trim_polyx(r1, c_p.polyx_min, pb1, pl1)) {
                        if (lane == 0) { atomicAdd(&bc->fr[FP_FR_POLYX_READS + pb1], 1u); atomicAdd(&bc->fr[FP_FR_POLYX_BASES + pb1], (unsigned)pl1); }
                        flags1 |= FP_F_POLYX_TRIMMED;
                    }
                    if (dev_trim_polyx(r2, c_p.polyx_min, pb2, pl2)) {
                        if (lane == 0) { atomicAdd(&bc->fr[FP_FR_POLYX_READS + pb2], 1u); atomicAdd(&bc->fr[FP_FR_POLYX_BASES + pb2], (unsigned)pl2); }
                        flags2 |= FP_F_POLYX_TRIMMED;
                    }
                }
                if (both) {                                                                       /* :511-516 */
                    if (c_p.max_len1 > 0 && c_p.max_len1 < r1.len) r1.len = c_p.max_len1;
                    if (c_p.max_len2 > 0 && c_p.max_len2 < r2.len) r2.len = c_p.max_len2;
                }
                int res1 = usep1 ? dev_pass_filter_planes(r1.qual, r1.len, r1.null, P1, r1.front, PW, s_lut) : dev_pass_filter(r1, s_lut);   /* :565-566 */
                int res2 = usep2 ? dev_pass_filter_planes(r2.qual, r2.len, r2.null, P2, r2.front, PW, s_lut) : dev_pass_filter(r2, s_lut);
                if (dimer) { res1 = res2 = FP_FAIL_ADAPTER_DIMER; flags1 |= FP_F_ADAPTER_DIMER; flags2 |= FP_F_ADAPTER_DIMER; }
                const int pv = max(res1, res2);
                if (lane == 0) atomicAdd(&bc->fr[FP_FR_READSTATS + pv], 2u);                       /* :573 */
                const bool counted = !r1.null && res1 == FP_PASS_FILTER && !r2.null && res2 == FP_PASS_FILTER;   /* :577-591 */
                if (counted) { rl[2] += 1; rl[3] += r1.len; rl[6] += 1; rl[7] += r2.len; }
                /* post stats as a delta against pre (per side) */
                #pragma unroll
                for (int sd = 0; sd < 2; sd++) {
                    const WRead& rr = sd ? r2 : r1;
                    uint8_t* rs = sd ? rs2 : rs1; uint8_t* rq = sd ? rq2 : rq1;
                    const int l0 = sd ? l2 : l1;
                    const bool clean = sd ? clean2 : clean1, removed = sd ? removed2 : removed1;
                    if (counted) {
                        if (rr.front == 0 && clean && !removed) post_delta(true, D, G, sd, rs, rq, 0, rr.len, l0, -1);
                        else {
                            if (clean && !removed) post_delta(true, D, G, sd, rs, rq, 0, 0, l0, -1);
                            post_delta(clean, D, G, sd, rs, rq, rr.front, rr.front, rr.front + rr.len, +1);
                        }
                    } else if (clean && !removed) post_delta(true, D, G, sd, rs, rq, 0, 0, l0, -1);
                }
                if (lane == 0) {
                    a.out1[gi] = make_result(r1, res1, pv, flags1, apos1, ab1, pb1, pl1);
                    a.out2[gi] = make_result(r2, res2, pv, flags2, apos2, ab2, pb2, pl2);
                    if (a.ov) a.ov[gi] = ov;
                }
            }
        }
        __syncthreads();
    }

    /* ---------------- flush block-level accumulators ---------------- */
    if (col_active) {
        const int BIN_SLOT[NB] = {1, 3, 4, 6, 7};      /* base & 7 of A C T N G */
        #pragma unroll
        for (int c = 0; c < 4; c++) {
            const int cyc = my_w * 4 + c;
            if (cyc >= L.cycles) continue;
            #pragma unroll
            for (int b = 0; b < NB; b++) {
                const unsigned int n = acc.v[c][b][0], n20 = acc.v[c][b][1], n30 = acc.v[c][b][2], sq = acc.v[c][b][3];
                if (n == 0) continue;
                const long long qs = (long long)sq - 33ll * (long long)n;
                #pragma unroll
                for (int pp = 0; pp < 2; pp++) {       /* dense pass feeds pre AND post (post gets deltas on top) */
                    const int st = my_side * 2 + pp;
                    if (n30) red_add64(&G[fp_off_cycle(&L, st, 0 * 8 + BIN_SLOT[b], cyc)], (unsigned long long)n30);
                    if (n20) red_add64(&G[fp_off_cycle(&L, st, 1 * 8 + BIN_SLOT[b], cyc)], (unsigned long long)n20);
                    red_add64(&G[fp_off_cycle(&L, st, 2 * 8 + BIN_SLOT[b], cyc)], (unsigned long long)n);
                    red_add64(&G[fp_off_cycle(&L, st, 3 * 8 + BIN_SLOT[b], cyc)], (unsigned long long)qs);
                }
            }
        }
    }
    for (int i = tid; i < SIDES * FP_KMER_BINS; i += FP_THREADS) {
        unsigned int v = s_kmer[i];
        if (v) { int sd = i / FP_KMER_BINS, k = i % FP_KMER_BINS; red_add64(&G[fp_off_kmer(&L, sd * 2, k)], (unsigned long long)v); red_add64(&G[fp_off_kmer(&L, sd * 2 + 1, k)], (unsigned long long)v); }
    }
    for (int i = tid; i < SIDES * FP_QUAL_BINS; i += FP_THREADS) {
        unsigned int v = s_qhist[i];
        if (v) { int sd = i / FP_QUAL_BINS, k = i % FP_QUAL_BINS; red_add64(&G[fp_off_qualhist(&L, sd * 2, k)], (unsigned long long)v); red_add64(&G[fp_off_qualhist(&L, sd * 2 + 1, k)], (unsigned long long)v); }
    }
    {   /* block-private post-filter deltas -> POST stats of each side */
        const int BIN_SLOT[NB] = {1, 3, 4, 6, 7};
        for (int i = tid; i < SIDES * S * 20; i += FP_THREADS) {
            const int v = D.cyc[i];
            if (v == 0) continue;
            const int sd = i / (S * 20), rem = i % (S * 20), cyc = rem / 20, bin = (rem % 20) / 4, kind = rem & 3;
            if (cyc >= L.cycles) continue;
            const int gk = kind == 0 ? 2 : kind == 1 ? 1 : kind == 2 ? 0 : 3;       /* count->content, q20, q30, qualsum */
            red_add64(&G[fp_off_cycle(&L, sd * 2 + 1, gk * 8 + BIN_SLOT[bin], cyc)], (unsigned long long)(long long)v);
        }
        for (int i = tid; i < SIDES * FP_KMER_BINS; i += FP_THREADS) { const int v = D.kmer[i]; if (v) red_add64(&G[fp_off_kmer(&L, (i / FP_KMER_BINS) * 2 + 1, i % FP_KMER_BINS)], (unsigned long long)(long long)v); }
        for (int i = tid; i < SIDES * FP_QUAL_BINS; i += FP_THREADS) { const int v = D.qh[i]; if (v) red_add64(&G[fp_off_qualhist(&L, (i / FP_QUAL_BINS) * 2 + 1, i % FP_QUAL_BINS)], (unsigned long long)(long long)v); }
    }
    for (int i = tid; i < FP_FR_WORDS; i += FP_THREADS) { unsigned int v = bc->fr[i]; if (v) red_add64(&G[L.off_filter + i], (unsigned long long)v); }
    if (c_p.isize_max < FP_MAX_ISIZE_SMEM)
        for (int i = tid; i <= c_p.isize_max; i += FP_THREADS) { unsigned int v = bc->isize[i]; if (v) red_add64(&G[L.off_isize + i], (unsigned long long)v); }
    if (lane == 0) {
        #pragma unroll
        for (int k = 0; k < 8; k++)
            if (k < 2 * L.n_stats && rl[k]) red_add64(&G[(k & 1) ? fp_off_length_sum(&L, k >> 1) : fp_off_reads(&L, k >> 1)], rl[k]);
    }
}

/* totals per cycle (kinds 32, 33 = sum over the 8 base slots of kinds 16..23 and 24..31; stats.cpp:225-226) */
__global__ void fp_finalize_kernel(const long long* __restrict__ raw, long long* __restrict__ fin, fp_counter_layout L) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= L.total) return;
    long long v = raw[i];
    if (i < L.off_filter) {
        long long in_stats = i % L.stats_stride;
        int st = (int)(i / L.stats_stride);
        if (in_stats < (long long)FP_CYCLE_KINDS * L.cycles) {
            int kind = (int)(in_stats / L.cycles), cyc = (int)(in_stats % L.cycles);
            if (kind == 32 || kind == 33) {
                v = 0;
                int k0 = kind == 32 ? 16 : 24;
                for (int b = 0; b < 8; b++) v += raw[fp_off_cycle(&L, st, k0 + b, cyc)];
            }
        }
    }
    fin[i] = v;
}

/* synthetic generator: one thread per read/pair */
#include "synth.h"
__global__ void fp_synth_kernel(fp_batch b, long long first_index, unsigned long long seed, int profile, int read_len) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    long long o = i * b.stride;
    fp_synth_pair(seed, (uint64_t)(first_index + i), profile, read_len, b.stride, b.seq1 + o, b.qual1 + o, b.len1 + i,
                  b.seq2 ? b.seq2 + o : nullptr, b.seq2 ? b.qual2 + o : nullptr, b.seq2 ? b.len2 + i : nullptr);
}
