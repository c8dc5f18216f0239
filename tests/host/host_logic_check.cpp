/*
 * host_logic_check.cpp -- CPU-only check of the host side above the C-ABI (tests/test_host_logic.py builds and runs it):
 *   mode "params": Options -> fp_params mapping (Options::toParams), every field set to a distinctive value, printed name=value
 *   mode "stats <file>": Stats::fill / fillOverRep on a counter block produced by the CPU oracle (file = fp_counter_layout + int64 block)
 * No device call is made: only layout helpers and fp_params_default of libfastp_b200.so are used.
 */
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include "fastp_host.h"

using namespace fastp_b200;

static int doParams() {
    Options o;
    o.paired = true;
    o.trim.front1 = 1; o.trim.tail1 = 2; o.trim.front2 = 3; o.trim.tail2 = 4; o.trim.maxLen1 = 101; o.trim.maxLen2 = 102;
    o.qualityCut.enabledFront = true; o.qualityCut.enabledTail = false; o.qualityCut.enabledRight = true;
    o.qualityCut.windowSizeFront = 5; o.qualityCut.qualityFront = 21; o.qualityCut.windowSizeTail = 6; o.qualityCut.qualityTail = 22;
    o.qualityCut.windowSizeRight = 7; o.qualityCut.qualityRight = 23;
    o.polyGTrim.enabled = true; o.polyGTrim.minLen = 11; o.polyXTrim.enabled = true; o.polyXTrim.minLen = 12;
    o.adapter.enabled = true; o.adapter.sequence = "AGATCGGAAGAGC"; o.adapter.sequenceR2 = "CTGTCTCTTATA"; o.adapter.hasSeqR1 = true; o.adapter.hasSeqR2 = true;
    o.adapter.hasFasta = true; o.adapter.seqsInFasta = {"AAAACCCC", "GGGGTTTTAA"}; o.adapter.allowGapOverlapTrimming = true; o.adapter.dimerMaxLen = 3;
    o.correction.enabled = true;
    o.qualfilter.enabled = true; o.qualfilter.qualifiedQual = '5'; o.qualfilter.unqualifiedPercentLimit = 41; o.qualfilter.nBaseLimit = 6; o.qualfilter.avgQualReq = 13;
    o.lengthFilter.enabled = true; o.lengthFilter.requiredLength = 16; o.lengthFilter.maxLength = 140;
    o.complexityFilter.enabled = true; o.complexityFilter.threshold = 0.25;
    o.overRepAnalysis.enabled = true; o.overRepAnalysis.sampling = 7;
    o.overRepSeqs1["ACGTACGTAC"] = 0; o.overRepSeqs1["TTTTTTTTTTGG"] = 0; o.overRepSeqs2["CCCCCCCCCC"] = 0;
    o.insertSizeMax = 600; o.overlapRequire = 31; o.overlapDiffLimit = 4; o.overlapDiffPercentLimit = 19; o.seqLen1 = 151; o.seqLen2 = 149;
    fp_params p; std::vector<const char*> k0, k1, k2;
    o.toParams(&p, k0, k1, k2);
#define P(f) printf(#f "=%lld\n", (long long)p.f)
    P(paired); P(thread0_semantics); P(trim_front1); P(trim_tail1); P(trim_front2); P(trim_tail2); P(max_len1); P(max_len2);
    P(cut_front); P(cut_tail); P(cut_right); P(cut_front_window); P(cut_front_quality); P(cut_tail_window); P(cut_tail_quality);
    P(cut_right_window); P(cut_right_quality); P(polyg_enabled); P(polyg_min_len); P(polyx_enabled); P(polyx_min_len);
    P(adapter_enabled); P(has_seq_r1); P(has_seq_r2); P(n_fasta_adapters); P(allow_gap_overlap_trimming); P(dimer_max_len);
    P(correction_enabled); P(overlap_require); P(overlap_diff_limit); P(overlap_diff_percent_limit);
    P(qual_filter_enabled); P(qualified_qual); P(unqualified_percent_limit); P(n_base_limit); P(avg_qual_req);
    P(length_filter_enabled); P(length_required); P(length_limit); P(complexity_filter_enabled);
    P(insert_size_max); P(seq_len1); P(seq_len2); P(overrep_enabled); P(overrep_sampling); P(n_overrep1); P(n_overrep2);
    printf("complexity_threshold=%.6f\n", p.complexity_threshold);
    printf("adapter_seq_r1=%s\nadapter_seq_r2=%s\n", p.adapter_seq_r1, p.adapter_seq_r2);
    for (int i = 0; i < p.n_fasta_adapters; i++) printf("fasta%d=%s\n", i, p.fasta_adapters[i]);
    for (int i = 0; i < p.n_overrep1; i++) printf("ovr1_%d=%s\n", i, p.overrep_seqs1[i]);
    for (int i = 0; i < p.n_overrep2; i++) printf("ovr2_%d=%s\n", i, p.overrep_seqs2[i]);
    return 0;
}

static int doStats(const char* path) {
    std::ifstream f(path, std::ios::binary);
    fp_counter_layout L;
    f.read(reinterpret_cast<char*>(&L), sizeof(L));
    std::vector<int64_t> B(L.total);
    f.read(reinterpret_cast<char*>(B.data()), (std::streamsize)(L.total * 8));
    if (!f) { fprintf(stderr, "short file\n"); return 1; }
    for (int which = 0; which < L.n_stats; which++) {
        Stats s; s.fill(B.data(), L, which);
        long kmer = 0, qh = 0;
        for (long v : s.mKmer) kmer += v;
        for (int q = 0; q < 128; q++) qh += s.mBaseQualHistogram[q];
        printf("stats%d reads=%ld bases=%ld q20=%ld q30=%ld cycles=%d length_sum=%ld kmer=%ld qualhist=%ld tq0=%ld tb0=%ld A5=%ld\n", which, s.mReads, s.mBases,
               s.mQ20Total, s.mQ30Total, s.mCycles, s.mLengthSum, kmer, qh, s.mCycleTotalQual[0], s.mCycleTotalBase[0], s.mCycleBaseContents['A' & 7][5]);
    }
    FilterResult fr; fr.fill(B.data(), L);
    long corr = 0;
    for (int i = 0; i < 64; i++) corr += fr.mCorrectionMatrix[i];
    printf("filter pass=%ld lowq=%ld nbase=%ld tooshort=%ld adapter_reads=%ld adapter_bases=%ld corrected_reads=%ld corrections=%ld polyx_reads=%ld polyx_bases=%ld\n",
           fr.mFilterReadStats[FP_PASS_FILTER], fr.mFilterReadStats[FP_FAIL_QUALITY], fr.mFilterReadStats[FP_FAIL_N_BASE], fr.mFilterReadStats[FP_FAIL_LENGTH],
           fr.mTrimmedAdapterRead, fr.mTrimmedAdapterBases, fr.mCorrectedReads, corr,
           fr.mTrimmedPolyXReads[0] + fr.mTrimmedPolyXReads[1] + fr.mTrimmedPolyXReads[2] + fr.mTrimmedPolyXReads[3],
           fr.mTrimmedPolyXBases[0] + fr.mTrimmedPolyXBases[1] + fr.mTrimmedPolyXBases[2] + fr.mTrimmedPolyXBases[3]);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "params")) return doParams();
    if (argc >= 3 && !strcmp(argv[1], "stats")) return doStats(argv[2]);
    fprintf(stderr, "usage: host_logic_check params | stats <file>\n");
    return 2;
}
