/*
 * fp_common.c -- parameter defaults and counter-block layout shared by the device library
 * (libfastp_b200.so), the CPU oracle (test infrastructure) and the reference harness.
 * Plain C, no CUDA.
 */
#include "fastp_b200.h"
#include <string.h>

/* Defaults: Options::Options() src/options.cpp:9-32, nested constructors src/options.h:20-284,
 * and what main() sets when no flag is given (src/main.cpp:330-343: quality filter on with
 * num2qual(15)='0', 40 %, N<=5, avg 0; length filter on with 15; complexity off, 30 %). */
void fp_params_default(fp_params* p, int paired) {
    memset(p, 0, sizeof(*p));
    p->paired = paired ? 1 : 0;
    p->thread0_semantics = 1;
    p->cut_front_window = p->cut_tail_window = p->cut_right_window = 4;     /* options.h:138-145 */
    p->cut_front_quality = p->cut_tail_quality = p->cut_right_quality = 20;
    p->polyg_min_len = 10;                                                  /* options.h:86  */
    p->polyx_min_len = 10;                                                  /* options.h:97  */
    p->adapter_enabled = 1;                                                 /* options.h:200 */
    p->dimer_max_len = 2;                                                   /* options.h:205 */
    p->overlap_require = 30;                                                /* options.cpp:24 */
    p->overlap_diff_limit = 5;                                              /* options.cpp:25 */
    p->overlap_diff_percent_limit = 20;                                     /* options.cpp:26 */
    p->qual_filter_enabled = 1;                                             /* options.h:251 */
    p->qualified_qual = '0';                                                /* options.h:253 */
    p->unqualified_percent_limit = 40;                                      /* options.h:254 */
    p->n_base_limit = 5;                                                    /* options.h:255 */
    p->avg_qual_req = 0;                                                    /* main.cpp:333 default */
    p->length_filter_enabled = 1;                                           /* main.cpp:337  */
    p->length_required = 15;                                                /* options.h:274 */
    p->length_limit = 0;
    p->complexity_filter_enabled = 0;
    p->complexity_threshold = 30 / 100.0;                                   /* main.cpp:343  */
    p->insert_size_max = 512;                                               /* options.cpp:23 */
    p->seq_len1 = p->seq_len2 = 151;                                        /* options.cpp:28-29 */
    p->overrep_enabled = 0;                                                 /* options.h:74  */
    p->overrep_sampling = 20;                                               /* options.h:75  */
}

void fp_counter_layout_make(fp_counter_layout* L, int paired, int cycles, int insert_size_max) {
    fp_counter_layout_make_overrep(L, paired, cycles, insert_size_max, 0, 0, 0, 0);
}

void fp_counter_layout_make_overrep(fp_counter_layout* L, int paired, int cycles, int insert_size_max, int k1, int len1, int k2, int len2) {
    memset(L, 0, sizeof(*L));
    L->cycles = cycles;
    L->n_stats = paired ? 4 : 2;
    L->isize_bins = insert_size_max + 1;
    L->off_kmer = (int64_t)FP_CYCLE_KINDS * cycles;
    L->off_qualhist = L->off_kmer + FP_KMER_BINS;
    L->off_reads = L->off_qualhist + FP_QUAL_BINS;
    L->off_length_sum = L->off_reads + 1;
    L->stats_stride = L->off_length_sum + 1;
    L->off_filter = (int64_t)L->n_stats * L->stats_stride;
    L->off_isize = L->off_filter + FP_FR_WORDS;
    L->n_overrep[0] = k1; L->n_overrep[1] = paired ? k2 : 0;
    L->overrep_len[0] = len1; L->overrep_len[1] = len2;
    int64_t off = L->off_isize + L->isize_bins;
    for (int s = 0; s < 4; s++) {
        L->off_overrep[s] = off;
        if (s < L->n_stats) off += (int64_t)L->n_overrep[s >> 1] * (1 + L->overrep_len[s >> 1]);
    }
    L->total = off;
}

/* ABI self-check for foreign-language bindings (ctypes / cgo): sizeof of each public struct. */
size_t fp_abi_sizeof(int which) {
    switch (which) {
        case 0: return sizeof(fp_params);
        case 1: return sizeof(fp_batch);
        case 2: return sizeof(fp_read_result);
        case 3: return sizeof(fp_ov_result);
        case 4: return sizeof(fp_patch);
        case 5: return sizeof(fp_counter_layout);
        case 6: return sizeof(fp_fastq_rec);
        case 7: return sizeof(fp_fastq_info);
        case 8: return sizeof(fp_adapter_event);
        case 9: return sizeof(fp_packed_batch);
        case 10: return sizeof(fp_npos);
        default: return 0;
    }
}
