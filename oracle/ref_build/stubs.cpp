// Definitions for the stub headers: compressed I/O is unsupported in the oracle build.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "libdeflate.h"
#include "isa-l/igzip_lib.h"
static void die(const char* what) { fprintf(stderr, "oracle build: %s unsupported (plain FASTQ only)\n", what); abort(); }
extern "C" {
struct libdeflate_compressor* libdeflate_alloc_compressor(int) { return (struct libdeflate_compressor*)malloc(8); }
size_t libdeflate_gzip_compress_bound(struct libdeflate_compressor*, size_t n) { return n + 64; }
size_t libdeflate_gzip_compress(struct libdeflate_compressor*, const void*, size_t, void*, size_t) { die("gzip output"); return 0; }
void libdeflate_free_compressor(struct libdeflate_compressor* c) { free(c); }
void isal_inflate_init(struct inflate_state* s) { memset(s, 0, sizeof(*s)); }
void isal_inflate_reset(struct inflate_state* s) { memset(s, 0, sizeof(*s)); }
int  isal_inflate(struct inflate_state*) { die("gzip input"); return -1; }
int  isal_inflate_stateless(struct inflate_state*) { die("bgzf input"); return -1; }
void isal_gzip_header_init(struct isal_gzip_header* h) { memset(h, 0, sizeof(*h)); }
int  isal_read_gzip_header(struct inflate_state*, struct isal_gzip_header*) { die("gzip input"); return -1; }
}
