/*
 * fp_gzio.cpp -- the wire format either side of the text path: gzip / BGZF in, gzip members out.  Host code.
 *
 * The reference inflates with isa-l on its reader thread (src/fastqreader.cpp:88-209; BGZF block sizes from the BC extra field,
 * src/bgzf.h:165-195) and deflates every output pack as ONE gzip member with libdeflate on its writer threads
 * (src/writerthread.cpp:118-168).  Same shapes here on zlib (the library this image has):
 *   fp_gz_inflate   whole buffer -> text.  BGZF input is indexed by its BSIZE fields and inflated block-parallel by `threads`
 *                   host threads; any other gzip stream (one or many members) is inflated sequentially, member after member.
 *   fp_gz_deflate   text -> concatenated gzip members of `member_bytes` input bytes each, compressed in parallel (a gzip reader
 *                   sees one stream; the decompressed bytes are what matter, member borders are free -- the reference's own
 *                   output differs from run to run with --thread).
 *   fp_gz_open / fp_gz_read / fp_gz_close   streaming reader for callers that feed fp_fastq_process_host chunk by chunk
 *                   (plain files pass through unchanged).
 */
#include <zlib.h>
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include "fastp_b200.h"

namespace {
struct Block { int64_t in_off, in_len, out_off, out_len; };

/* BGZF: gzip member whose extra field holds subfield 'B','C' with BSIZE = total block size - 1 (src/bgzf.h:165-195) */
bool bgzf_block_size(const uint8_t* p, int64_t n, int64_t* bsize) {
    if (n < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return false;
    const int xlen = p[10] | (p[11] << 8);
    if (12 + xlen > n) return false;
    for (int o = 12; o + 4 <= 12 + xlen;) {
        const int slen = p[o + 2] | (p[o + 3] << 8);
        if (p[o] == 'B' && p[o + 1] == 'C' && slen == 2) { *bsize = (int64_t)(p[o + 4] | (p[o + 5] << 8)) + 1; return true; }
        o += 4 + slen;
    }
    return false;
}

int inflate_block(const uint8_t* in, int64_t n_in, uint8_t* out, int64_t n_out) {      /* one complete gzip member */
    z_stream z; memset(&z, 0, sizeof(z));
    if (inflateInit2(&z, 15 + 16) != Z_OK) return -1;
    z.next_in = const_cast<Bytef*>(in); z.avail_in = (uInt)n_in; z.next_out = out; z.avail_out = (uInt)n_out;
    const int rc = inflate(&z, Z_FINISH);
    const bool ok = rc == Z_STREAM_END && (int64_t)z.total_out == n_out;
    inflateEnd(&z);
    return ok ? 0 : -1;
}
}  // namespace

extern "C" {

int fp_gz_is_bgzf(const uint8_t* in, int64_t n) { int64_t b; return in && bgzf_block_size(in, n, &b) ? 1 : 0; }

int fp_gz_inflate(const uint8_t* in, int64_t n_in, uint8_t* out, int64_t cap, int64_t* n_out, int threads) {
    if (!in || !n_out || n_in < 0 || (cap > 0 && !out)) return FP_E_INVAL;
    *n_out = 0;
    if (n_in == 0) return FP_OK;
    int64_t bs;
    if (bgzf_block_size(in, n_in, &bs)) {
        std::vector<Block> blocks;
        int64_t off = 0, oo = 0;
        while (off < n_in) {
            if (!bgzf_block_size(in + off, n_in - off, &bs) || off + bs > n_in || bs < 26) return FP_E_INVAL;
            const uint8_t* t = in + off + bs - 4;
            const int64_t isize = (int64_t)t[0] | ((int64_t)t[1] << 8) | ((int64_t)t[2] << 16) | ((int64_t)t[3] << 24);
            blocks.push_back(Block{off, bs, oo, isize});
            off += bs; oo += isize;
        }
        *n_out = oo;
        if (oo > cap) return FP_E_TOOLARGE;
        std::atomic<size_t> next(0); std::atomic<int> bad(0);
        auto work = [&]() {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= blocks.size()) return;
                const Block& b = blocks[i];
                if (b.out_len > 0 && inflate_block(in + b.in_off, b.in_len, out + b.out_off, b.out_len) != 0) bad = 1;
            }
        };
        std::vector<std::thread> th;
        const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(threads, 1), blocks.size()));
        for (int t = 1; t < nt; t++) th.emplace_back(work);
        work();
        for (auto& x : th) x.join();
        return bad ? FP_E_INVAL : FP_OK;
    }
    /* any other gzip stream: member after member (the reference's reader does the same, src/fastqreader.cpp:130-149) */
    z_stream z; memset(&z, 0, sizeof(z));
    if (inflateInit2(&z, 15 + 32) != Z_OK) return FP_E_NOMEM;
    int64_t ip = 0, op = 0;
    int rc = Z_OK;
    while (ip < n_in) {
        z.next_in = const_cast<Bytef*>(in + ip); z.avail_in = (uInt)std::min<int64_t>(n_in - ip, 1 << 30);
        z.next_out = out + op; z.avail_out = (uInt)std::min<int64_t>(cap - op, 1 << 30);
        const uInt ai = z.avail_in, ao = z.avail_out;
        rc = inflate(&z, Z_NO_FLUSH);
        ip += ai - z.avail_in; op += ao - z.avail_out;
        if (rc == Z_STREAM_END) { if (ip < n_in) inflateReset(&z); continue; }
        if (rc != Z_OK) { inflateEnd(&z); *n_out = op; return rc == Z_BUF_ERROR && op >= cap ? FP_E_TOOLARGE : FP_E_INVAL; }
        if (z.avail_out == 0 && op >= cap) { inflateEnd(&z); *n_out = op; return FP_E_TOOLARGE; }
    }
    inflateEnd(&z);
    *n_out = op;
    return FP_OK;
}

int64_t fp_gz_deflate_bound(int64_t n_in, int64_t member_bytes) {
    if (member_bytes <= 0) member_bytes = 1 << 20;
    const int64_t members = (n_in + member_bytes - 1) / member_bytes + 1;
    return n_in + n_in / 1000 + members * 64 + 1024;
}

int fp_gz_deflate(const uint8_t* in, int64_t n_in, uint8_t* out, int64_t cap, int64_t* n_out, int64_t member_bytes, int level, int threads) {
    if (!n_out || n_in < 0 || (n_in > 0 && !in) || (cap > 0 && !out)) return FP_E_INVAL;
    *n_out = 0;
    if (member_bytes <= 0) member_bytes = 1 << 20;
    const int64_t nm = (n_in + member_bytes - 1) / member_bytes;
    std::vector<std::vector<uint8_t>> parts((size_t)nm);
    std::atomic<int64_t> next(0); std::atomic<int> bad(0);
    auto work = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= nm) return;
            const int64_t lo = i * member_bytes, len = std::min(member_bytes, n_in - lo);
            z_stream z; memset(&z, 0, sizeof(z));
            if (deflateInit2(&z, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad = 1; return; }
            std::vector<uint8_t>& o = parts[(size_t)i];
            o.resize((size_t)deflateBound(&z, (uLong)len) + 32);
            z.next_in = const_cast<Bytef*>(in + lo); z.avail_in = (uInt)len; z.next_out = o.data(); z.avail_out = (uInt)o.size();
            if (deflate(&z, Z_FINISH) != Z_STREAM_END) bad = 1;
            o.resize(z.total_out);
            deflateEnd(&z);
        }
    };
    std::vector<std::thread> th;
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::max(threads, 1), nm));
    for (int t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
    if (bad) return FP_E_INVAL;
    int64_t total = 0;
    for (auto& o : parts) total += (int64_t)o.size();
    *n_out = total;
    if (total > cap) return FP_E_TOOLARGE;
    int64_t op = 0;
    for (auto& o : parts) { memcpy(out + op, o.data(), o.size()); op += (int64_t)o.size(); }
    return FP_OK;
}

void* fp_gz_open(const char* path) { return path ? (void*)gzopen(path, "rb") : nullptr; }
int64_t fp_gz_read(void* h, uint8_t* buf, int64_t cap) {
    if (!h || !buf || cap < 0) return -1;
    int64_t got = 0;
    while (got < cap) {
        const int r = gzread((gzFile)h, buf + got, (unsigned)std::min<int64_t>(cap - got, 1 << 30));
        if (r < 0) return -1;
        if (r == 0) break;
        got += r;
    }
    return got;
}
void fp_gz_close(void* h) { if (h) gzclose((gzFile)h); }

}  // extern "C"
