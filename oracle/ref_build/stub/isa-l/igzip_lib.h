/* Stub of isa-l/igzip_lib.h (isa-l v2.31.1 is not vendored / installed here).  Only the names the
 * reference's FASTQ reader touches (src/fastqreader.cpp:88-209, src/bgzf.h:165-195).  gzip input is
 * not on the per-read hot path; stubs.cpp makes the functions fail loudly (plain FASTQ only). */
#ifndef IGZIP_LIB_STUB_H
#define IGZIP_LIB_STUB_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#define ISAL_DECOMP_OK 0
#define ISAL_GZIP 1
#define ISAL_GZIP_NO_HDR_VER 3
#define ISAL_BLOCK_FINISH 11
struct inflate_state {
    uint8_t* next_out; uint32_t avail_out; uint32_t total_out;
    uint8_t* next_in;  uint32_t avail_in;
    uint32_t crc_flag; int block_state; uint32_t bfinal;
};
struct isal_gzip_header { uint32_t dummy; };
void isal_inflate_init(struct inflate_state* s);
void isal_inflate_reset(struct inflate_state* s);
int  isal_inflate(struct inflate_state* s);
int  isal_inflate_stateless(struct inflate_state* s);
void isal_gzip_header_init(struct isal_gzip_header* h);
int  isal_read_gzip_header(struct inflate_state* s, struct isal_gzip_header* h);
#ifdef __cplusplus
}
#endif
#endif
