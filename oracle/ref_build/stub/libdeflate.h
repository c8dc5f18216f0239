/* Stub of libdeflate.h (libdeflate v1.23 is not vendored / installed here).  Only the four
 * declarations the reference's writers use (src/writer.cpp:71-138, src/writerthread.cpp:34-177).
 * gzip output is not on the per-read hot path; stubs.cpp makes them fail loudly. */
#ifndef LIBDEFLATE_STUB_H
#define LIBDEFLATE_STUB_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
struct libdeflate_compressor;
struct libdeflate_compressor* libdeflate_alloc_compressor(int level);
size_t libdeflate_gzip_compress_bound(struct libdeflate_compressor* c, size_t in_nbytes);
size_t libdeflate_gzip_compress(struct libdeflate_compressor* c, const void* in, size_t in_nbytes, void* out, size_t out_nbytes_avail);
void libdeflate_free_compressor(struct libdeflate_compressor* c);
#ifdef __cplusplus
}
#endif
#endif
