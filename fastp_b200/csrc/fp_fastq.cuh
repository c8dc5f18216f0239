/*
 * fp_fastq.cuh -- FASTQ text <-> SoA rows on the device (the data formats either side of the hot path).
 *
 * decode  = FastqReader::getLine + FastqReader::read   src/fastqreader.cpp:240-368
 * encode  = Read::appendToString                        src/read.cpp:119-134, called for pairs/reads that pass
 *           (src/peprocessor.cpp:583-584, src/seprocessor.cpp:268)
 *
 * Semantics restated from the reference (tests/test_fastq_codec.py pins them against FastqReader itself):
 *   - a line ends at '\n' or at a '\r' that is not followed by '\n'; "\r\n" is ONE terminator and the '\r' is not content
 *     (getLine :245-266);
 *   - a record starts at the first line that is non-empty and begins with '@' (lines before it are skipped, :338-341);
 *     the next three lines are sequence, strand, quality WHATEVER they contain;
 *   - strand empty or not starting with '+' (:349), or |quality| != |sequence| (:356): the reader returns NULL, i.e. the
 *     input ENDS at that record -- the device reports the index of the first such record and the caller drops the rest;
 *   - a final line without terminator still counts (bufferFinished, :256).
 * The "which line is a record start" question is a 4-state automaton over lines (state = line position in the record);
 * it is evaluated in parallel as a scan over transition functions, so blank lines / junk between records behave
 * exactly like the sequential reader.
 *
 * Layout of the work: byte blocks of FQ_BB bytes (terminator positions), line blocks of FQ_LB lines (automaton),
 * one warp per record (scatter / gather copies).  All offsets are 32-bit: a chunk is < 4 GiB.
 */
#pragma once
#include "fp_device.cuh"

#define FQ_T 256
#define FQ_BPT 64                      /* bytes per thread in the terminator passes */
#define FQ_BB (FQ_T * FQ_BPT)          /* bytes per block */
#define FQ_LPT 8                       /* lines per thread in the automaton passes */
#define FQ_LB (FQ_T * FQ_LPT)          /* lines per block */

#define FQ_ERR_NONE 0
#define FQ_ERR_STRAND 1                /* "Expected '+'"                        fastqreader.cpp:349 */
#define FQ_ERR_LENGTH 2                /* sequence and quality differ in length fastqreader.cpp:356 */
#define FQ_ERR_STRIDE 3                /* read longer than the row stride (not a reference error: the rows are ours) */

struct fq_rec { unsigned int name_off, name_len, strand_off, strand_len; };      /* 16 B per record, offsets into the chunk */

__device__ __forceinline__ bool fq_is_term(const uint8_t* t, long long n, long long i) {
    const uint8_t c = t[i];
    return c == '\n' || (c == '\r' && (i + 1 >= n || t[i + 1] != '\n'));
}

/* ---- pass 1/2: terminator positions ---- */
__device__ __forceinline__ unsigned int fq_thread_terms(const uint8_t* text, long long n, long long b0, unsigned long long& bits) {
    /* terminator mask of this thread's FQ_BPT bytes (bit k = byte b0 + k) */
    bits = 0;
    if (b0 >= n) return 0;
    const long long e = min(b0 + (long long)FQ_BPT, n);
    if (e - b0 == FQ_BPT && ((reinterpret_cast<uintptr_t>(text + b0) & 15) == 0)) {
        const uint4* p = reinterpret_cast<const uint4*>(text + b0);
        uint8_t nextc = (e < n) ? text[e] : 0;
        #pragma unroll
        for (int v = 0; v < FQ_BPT / 16; v++) {
            const uint4 w = p[v];
            const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                #pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int idx = v * 16 + k * 4 + b;
                    const uint8_t c = (uint8_t)(ws[k] >> (8 * b));
                    uint8_t nx;
                    if (b < 3) nx = (uint8_t)(ws[k] >> (8 * b + 8));
                    else if (k < 3) nx = (uint8_t)ws[k + 1];
                    else if (v + 1 < FQ_BPT / 16) nx = reinterpret_cast<const uint8_t*>(p + v + 1)[0];
                    else nx = nextc;
                    const bool last = (b0 + idx + 1 >= n);
                    if (c == '\n' || (c == '\r' && (last || nx != '\n'))) bits |= 1ull << idx;
                }
            }
        }
    } else {
        for (long long i = b0; i < e; i++) if (fq_is_term(text, n, i)) bits |= 1ull << (int)(i - b0);
    }
    return (unsigned int)__popcll(bits);
}

__device__ __forceinline__ unsigned int fq_block_excl_scan(unsigned int v, unsigned int* s_w, unsigned int& total) {
    /* exclusive prefix of v over the block's threads (FQ_T), total = block sum */
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned int inc = v;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(FULL_MASK, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    unsigned int before = 0, tot = 0;
    #pragma unroll
    for (int k = 0; k < FQ_T / 32; k++) { const unsigned int t = s_w[k]; if (k < w) before += t; tot += t; }
    __syncthreads();
    total = tot;
    return before + inc - v;
}

__global__ void __launch_bounds__(FQ_T) fq_term_count_kernel(const uint8_t* text, long long n, unsigned int* block_cnt) {
    __shared__ unsigned int s_w[FQ_T / 32];
    unsigned long long bits;
    const unsigned int c = fq_thread_terms(text, n, (long long)blockIdx.x * FQ_BB + (long long)threadIdx.x * FQ_BPT, bits);
    unsigned int tot;
    fq_block_excl_scan(c, s_w, tot);
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = tot;
}
/* single thread: exclusive offsets of the byte blocks; info[0] = number of lines (incl. a final unterminated line when `final`) */
__global__ void fq_term_scan_kernel(unsigned int* block_cnt, int nblocks, const uint8_t* text, long long n, int final_chunk,
                                    unsigned int* term, unsigned int term_cap, unsigned int* info) {
    if (blockIdx.x || threadIdx.x) return;
    unsigned int run = 0;
    for (int i = 0; i < nblocks; i++) { const unsigned int v = block_cnt[i]; block_cnt[i] = run; run += v; }
    info[1] = run;                                              /* terminated lines */
    unsigned int nl = run;
    if (final_chunk && n > 0 && !fq_is_term(text, n, n - 1) && !(text[n - 1] == '\r')) {
        if (run < term_cap) term[run] = (unsigned int)n;        /* virtual terminator after the last byte */
        nl = run + 1;
    }
    info[0] = nl;
}
__global__ void __launch_bounds__(FQ_T) fq_term_fill_kernel(const uint8_t* text, long long n, const unsigned int* block_off,
                                                            unsigned int* term, unsigned int term_cap) {
    __shared__ unsigned int s_w[FQ_T / 32];
    unsigned long long bits;
    const long long b0 = (long long)blockIdx.x * FQ_BB + (long long)threadIdx.x * FQ_BPT;
    const unsigned int c = fq_thread_terms(text, n, b0, bits);
    unsigned int tot;
    unsigned int k = block_off[blockIdx.x] + fq_block_excl_scan(c, s_w, tot);
    while (bits) {
        const int b = __ffsll((long long)bits) - 1;
        bits &= bits - 1;
        if (k < term_cap) term[k] = (unsigned int)(b0 + b);
        k++;
    }
}

/* ---- line helpers ---- */
__device__ __forceinline__ void fq_line_span(const uint8_t* text, long long n, const unsigned int* term, unsigned int k,
                                             unsigned int& start, unsigned int& end) {
    start = k == 0 ? 0u : term[k - 1] + 1u;
    end = term[k];                                              /* position of the terminator (== n for the virtual one) */
    if ((long long)end < n && text[end] == '\n' && end > start && text[end - 1] == '\r') end--;     /* "\r\n": '\r' is not content */
}
__device__ __forceinline__ bool fq_namelike(const uint8_t* text, long long n, const unsigned int* term, unsigned int k) {
    unsigned int s, e;
    fq_line_span(text, n, term, k, s, e);
    return e > s && text[s] == '@';
}

/* transition function of the record automaton as 4 x 2 bits (F >> 2s) & 3 = next state from s; record counts per start state */
struct fq_elem { unsigned int F; unsigned int C[4]; };
__device__ __forceinline__ fq_elem fq_identity() { fq_elem e; e.F = 0xE4u; e.C[0] = e.C[1] = e.C[2] = e.C[3] = 0; return e; }   /* 3,2,1,0 */
__device__ __forceinline__ fq_elem fq_line_elem(bool namelike) {
    fq_elem e;
    e.F = namelike ? 0x39u : 0x38u;          /* s0->1 (or 0), s1->2, s2->3, s3->0 */
    e.C[0] = namelike ? 1u : 0u; e.C[1] = e.C[2] = e.C[3] = 0;
    return e;
}
/* a then b */
__device__ __forceinline__ fq_elem fq_compose(const fq_elem& a, const fq_elem& b) {
    fq_elem r; r.F = 0;
    #pragma unroll
    for (int s = 0; s < 4; s++) {
        const unsigned int m = (a.F >> (2 * s)) & 3u;
        r.F |= ((b.F >> (2 * m)) & 3u) << (2 * s);
        r.C[s] = a.C[s] + (m == 0 ? b.C[0] : m == 1 ? b.C[1] : m == 2 ? b.C[2] : b.C[3]);
    }
    return r;
}
__device__ __forceinline__ fq_elem fq_shfl_up(const fq_elem& e, int o) {
    fq_elem r;
    r.F = __shfl_up_sync(FULL_MASK, e.F, o);
    #pragma unroll
    for (int s = 0; s < 4; s++) r.C[s] = __shfl_up_sync(FULL_MASK, e.C[s], o);
    return r;
}

/* mode 0: block aggregates; mode 1: emit record start lines given the state / record base at the block start */
template <int MODE>
__global__ void __launch_bounds__(FQ_T) fq_fsm_kernel(const uint8_t* text, long long n, const unsigned int* term, unsigned int nlines,
                                                      fq_elem* block_agg, const unsigned int* block_state, const unsigned int* block_rec,
                                                      unsigned int* rec_line, unsigned int rec_cap) {
    __shared__ fq_elem s_w[FQ_T / 32];
    const unsigned int l0 = (unsigned int)blockIdx.x * FQ_LB + (unsigned int)threadIdx.x * FQ_LPT;
    unsigned int nl_mask = 0;
    fq_elem mine = fq_identity();
    #pragma unroll
    for (int k = 0; k < FQ_LPT; k++) {
        const unsigned int l = l0 + k;
        if (l < nlines) {
            const bool nm = fq_namelike(text, n, term, l);
            if (nm) nl_mask |= 1u << k;
            mine = fq_compose(mine, fq_line_elem(nm));
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    fq_elem inc = mine;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const fq_elem t = fq_shfl_up(inc, o); if (lane >= o) inc = fq_compose(t, inc); }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (MODE == 0) {
        if (threadIdx.x == 0) {
            fq_elem a = s_w[0];
            for (int k = 1; k < FQ_T / 32; k++) a = fq_compose(a, s_w[k]);
            block_agg[blockIdx.x] = a;
        }
        return;
    }
    /* exclusive prefix of this thread = (warps before) then (lanes before) */
    fq_elem pre = fq_identity();
    for (int k = 0; k < w; k++) pre = fq_compose(pre, s_w[k]);
    fq_elem lanes_before = fq_shfl_up(inc, 1);
    if (lane == 0) lanes_before = fq_identity();
    pre = fq_compose(pre, lanes_before);
    const unsigned int s0 = block_state[blockIdx.x];
    unsigned int st = (pre.F >> (2 * s0)) & 3u;
    unsigned int rec = block_rec[blockIdx.x] + pre.C[s0];
    #pragma unroll
    for (int k = 0; k < FQ_LPT; k++) {
        const unsigned int l = l0 + k;
        if (l < nlines) {
            const bool nm = (nl_mask >> k) & 1u;
            if (st == 0) { if (nm) { if (rec < rec_cap) rec_line[rec] = l; rec++; st = 1; } }
            else st = (st + 1) & 3u;
        }
    }
}
/* single thread: state and record count at every block start; info[2] = records started, info[3] = complete records */
__global__ void fq_fsm_scan_kernel(const fq_elem* block_agg, int nblocks, unsigned int* block_state, unsigned int* block_rec, unsigned int* info) {
    if (blockIdx.x || threadIdx.x) return;
    unsigned int st = 0, rec = 0;
    for (int b = 0; b < nblocks; b++) {
        block_state[b] = st; block_rec[b] = rec;
        const fq_elem a = block_agg[b];
        rec += a.C[st];
        st = (a.F >> (2 * st)) & 3u;
    }
    info[2] = rec;
    /* the last record is complete iff its quality line exists, i.e. the automaton is back in state 0 */
    info[3] = (st == 0) ? rec : (rec > 0 ? rec - 1 : 0);
}

/* ---- scatter: one warp per record ---- */
__global__ void __launch_bounds__(FQ_T) fq_scatter_kernel(const uint8_t* text, long long n, const unsigned int* term, const unsigned int* rec_line,
                                                          unsigned int nrec, int stride, int phred64,
                                                          uint8_t* seq, uint8_t* qual, uint16_t* len, fq_rec* recs,
                                                          unsigned int* rec_end, unsigned int* first_bad, unsigned int* bad_code) {
    const unsigned int r = blockIdx.x * (FQ_T / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= nrec) return;
    const unsigned int l = rec_line[r];
    unsigned int ns, ne, ss, se, ps, pe, qs, qe;
    fq_line_span(text, n, term, l, ns, ne);
    fq_line_span(text, n, term, l + 1, ss, se);
    fq_line_span(text, n, term, l + 2, ps, pe);
    fq_line_span(text, n, term, l + 3, qs, qe);
    int code = FQ_ERR_NONE;
    if (pe == ps || text[ps] != '+') code = FQ_ERR_STRAND;
    else if (qe - qs != se - ss) code = FQ_ERR_LENGTH;
    else if ((int)(se - ss) > stride) code = FQ_ERR_STRIDE;
    if (code != FQ_ERR_NONE) {
        if (lane == 0) { const unsigned int old = atomicMin(first_bad, r); if (r < old) atomicExch(bad_code, (unsigned int)code); }
        /* a later, smaller r may overwrite the code again: the host re-reads the code of first_bad through rec codes below */
    }
    const int L = code == FQ_ERR_NONE ? (int)(se - ss) : 0;
    uint8_t* srow = seq + (size_t)r * stride; uint8_t* qrow = qual + (size_t)r * stride;
    for (int i = lane; i < stride; i += 32) {
        uint8_t b = 0, q = 0;
        if (i < L) {
            b = text[ss + i]; q = text[qs + i];
            if (phred64) { const int v = (int)(signed char)q - 31; q = (uint8_t)(v < 33 ? 33 : v); }      /* read.cpp:35-39 */
        }
        srow[i] = b; qrow[i] = q;
    }
    if (lane == 0) {
        len[r] = (uint16_t)L;
        fq_rec rc; rc.name_off = ns; rc.name_len = ne - ns; rc.strand_off = ps; rc.strand_len = pe - ps;
        recs[r] = rc;
        /* first byte after this record's quality line (and its terminator) */
        const unsigned int t = term[l + 3];
        rec_end[r] = (long long)t < n ? t + 1u : (unsigned int)n;
        if (code != FQ_ERR_NONE) recs[r].name_len |= 0x80000000u | ((unsigned int)code << 28);     /* marks the record as bad */
    }
}

/* ---- finish: everything the host needs from a decode, written to (mapped) host memory by one thread ----
 * out[0] records kept, out[1] error code, out[2] error record (0xFFFFFFFF none), out[3] more, out[4] consumed bytes */
__global__ void fq_finish_kernel(const unsigned int* term, unsigned int nlines, unsigned int nterm, long long nbytes,
                                 const unsigned int* rec_line, unsigned int nstarted, unsigned int ncomplete, unsigned int nrec,
                                 const fq_rec* recs, const unsigned int* first_bad, unsigned int* out) {
    if (blockIdx.x || threadIdx.x) return;
    const unsigned int fb = nrec > 0 ? *first_bad : 0xFFFFFFFFu;
    unsigned int keep = nrec, err = FQ_ERR_NONE, more = 0;
    long long consumed;
    auto line_start = [&](unsigned int l) -> long long { return l == 0 ? 0ll : (long long)term[l - 1] + 1; };
    if (fb != 0xFFFFFFFFu) {                                   /* the reference reader stops here: fastqreader.cpp:349-364 */
        err = (recs[fb].name_len >> 28) & 7u; keep = fb; consumed = nbytes;
    } else if (ncomplete > nrec) {                             /* capacity reached: resume at the next record's name line */
        more = 1; consumed = line_start(rec_line[nrec]);
    } else if (nstarted > ncomplete) {                         /* the last record is not complete in this chunk: resume at its name line */
        consumed = line_start(rec_line[ncomplete]);
    } else {                                                   /* every complete line was a record line or skipped */
        consumed = nlines > nterm ? nbytes : (long long)term[nlines - 1] + 1;
    }
    out[0] = keep; out[1] = err; out[2] = fb; out[3] = more;
    out[4] = (unsigned int)(consumed & 0xFFFFFFFFll); out[5] = (unsigned int)(consumed >> 32);
}
__global__ void fq_set_u32_kernel(unsigned int* p, unsigned int v) { if (!blockIdx.x && !threadIdx.x) *p = v; }
__global__ void fq_copy_u32_kernel(const unsigned int* src, unsigned int* dst) { if (!blockIdx.x && !threadIdx.x) *dst = *src; }

/* ---- encode ---- */
#define FQ_SCAN_ITEMS 2048
__global__ void __launch_bounds__(FQ_T) fq_size_blocksum_kernel(const fq_rec* recs, const fp_read_result* res, long long n, unsigned long long* blocksum) {
    __shared__ unsigned long long s[FQ_T / 32];
    const long long b0 = (long long)blockIdx.x * FQ_SCAN_ITEMS;
    unsigned long long c = 0;
    for (int k = threadIdx.x; k < FQ_SCAN_ITEMS; k += FQ_T) {
        const long long i = b0 + k;
        if (i < n && res[i].pair_verdict == FP_PASS_FILTER && !(res[i].flags & FP_F_DUPLICATE)) c += (unsigned long long)(recs[i].name_len & 0x0FFFFFFFu) + recs[i].strand_len + 2ull * res[i].len + 4ull;
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL_MASK, c, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long t = 0; for (int w = 0; w < FQ_T / 32; w++) t += s[w]; blocksum[blockIdx.x] = t; }
}
__global__ void fq_size_scan_kernel(unsigned long long* blocksum, int nblocks, unsigned long long* total) {
    if (blockIdx.x || threadIdx.x) return;
    unsigned long long run = 0;
    for (int i = 0; i < nblocks; i++) { const unsigned long long v = blocksum[i]; blocksum[i] = run; run += v; }
    *total = run;
}
/* one warp per record, FQ_SCAN_ITEMS records per block in order: the block's warps walk the records 8 at a time */
__global__ void __launch_bounds__(FQ_T) fq_encode_kernel(const uint8_t* text, const fq_rec* recs, const fp_read_result* res,
                                                         const uint8_t* seq, const uint8_t* qual, int stride, long long n,
                                                         const unsigned long long* blockoff, uint8_t* out, unsigned long long out_cap) {
    __shared__ unsigned long long s_run;
    __shared__ unsigned long long s_sz[FQ_T];
    const long long b0 = (long long)blockIdx.x * FQ_SCAN_ITEMS;
    if (threadIdx.x == 0) s_run = blockoff[blockIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k0 = 0; k0 < FQ_SCAN_ITEMS; k0 += FQ_T) {
        /* sizes of 256 consecutive records, exclusive scan inside the block */
        const long long i = b0 + k0 + threadIdx.x;
        unsigned long long sz = 0;
        if (i < n && res[i].pair_verdict == FP_PASS_FILTER && !(res[i].flags & FP_F_DUPLICATE)) sz = (unsigned long long)(recs[i].name_len & 0x0FFFFFFFu) + recs[i].strand_len + 2ull * res[i].len + 4ull;
        unsigned long long inc = sz;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(FULL_MASK, inc, o); if (lane >= o) inc += t; }
        __shared__ unsigned long long s_w[FQ_T / 32];
        if (lane == 31) s_w[w] = inc;
        __syncthreads();
        unsigned long long before = s_run;
        for (int k = 0; k < w; k++) before += s_w[k];
        s_sz[threadIdx.x] = before + inc - sz;                       /* output offset of record i */
        __syncthreads();
        /* copies: warp w takes records w, w+8, ... of this group */
        for (int j = w; j < FQ_T; j += FQ_T / 32) {
            const long long ri = b0 + k0 + j;
            if (ri >= n) break;
            const fp_read_result rr = res[ri];
            if (rr.pair_verdict != FP_PASS_FILTER || (rr.flags & FP_F_DUPLICATE)) continue;
            const fq_rec rc = recs[ri];
            const unsigned int nl = rc.name_len & 0x0FFFFFFFu;
            unsigned long long o = s_sz[j];
            const unsigned long long need = (unsigned long long)nl + rc.strand_len + 2ull * rr.len + 4ull;
            if (o + need > out_cap) continue;                         /* caller sees total > cap */
            uint8_t* d = out + o;
            for (unsigned int t = lane; t < nl; t += 32) d[t] = text[rc.name_off + t];
            if (lane == 0) d[nl] = '\n';
            d += nl + 1;
            const uint8_t* srow = seq + (size_t)ri * stride + rr.front;
            for (unsigned int t = lane; t < rr.len; t += 32) d[t] = srow[t];
            if (lane == 0) d[rr.len] = '\n';
            d += rr.len + 1;
            for (unsigned int t = lane; t < rc.strand_len; t += 32) d[t] = text[rc.strand_off + t];
            if (lane == 0) d[rc.strand_len] = '\n';
            d += rc.strand_len + 1;
            const uint8_t* qrow = qual + (size_t)ri * stride + rr.front;
            for (unsigned int t = lane; t < rr.len; t += 32) d[t] = qrow[t];
            if (lane == 0) d[rr.len] = '\n';
        }
        __syncthreads();
        if (threadIdx.x == FQ_T - 1) s_run = s_sz[FQ_T - 1] + sz;
        __syncthreads();
    }
}
