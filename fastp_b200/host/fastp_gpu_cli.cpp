/*
 * fastp_gpu_cli.cpp -- minimal plain-FASTQ driver around GpuChainWorker (the caller either side of the hot path,
 * SURVEY.md 8(f) rank 1 kept on the host for now): reads R1[/R2], packs reads like the reference's reader
 * (src/peprocessor.cpp:760-813), runs the device chain through the drop-in bodies, writes the passing reads and a
 * small JSON summary.  Flags are the reference's own spellings (src/main.cpp:32-158) for the options the chain reads.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include "fastp_host.h"

using namespace fastp_b200;

static Read* readRecord(std::istream& in) {          /* FastqReader::read src/fastqreader.cpp:309-368 (plain text only) */
    std::string name, seq, strand, qual;
    if (!std::getline(in, name)) return nullptr;
    while (name.empty()) if (!std::getline(in, name)) return nullptr;
    if (!std::getline(in, seq) || !std::getline(in, strand) || !std::getline(in, qual)) return nullptr;
    if (name[0] != '@' || seq.size() != qual.size()) return nullptr;
    return new Read(new std::string(name), new std::string(seq), new std::string(strand), new std::string(qual));
}

int main(int argc, char** argv) {
    Options opt;
    std::string in1, in2, out1, out2, json;
    int packSize = 1 << 16, maxLen = 0;
    bool deviceFastq = false, phred64 = false;
    bool dedup = false, evalDup = true; int dupLevel = 0;     /* src/main.cpp:200-209 */
    int zlevel = 4, zthreads = 8;          /* -z / --compression (src/main.cpp:50), host threads for .gz output */
    size_t chunkBytes = 0;                 /* text path: bytes read per side per step (default: about one device batch) */
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "-i" || a == "--in1") in1 = next(); else if (a == "-I" || a == "--in2") in2 = next();
        else if (a == "-o" || a == "--out1") out1 = next(); else if (a == "-O" || a == "--out2") out2 = next();
        else if (a == "-j" || a == "--json") json = next();
        else if (a == "-A" || a == "--disable_adapter_trimming") opt.adapter.enabled = false;
        else if (a == "-a" || a == "--adapter_sequence") { opt.adapter.sequence = next(); opt.adapter.hasSeqR1 = true; }
        else if (a == "--adapter_sequence_r2") { opt.adapter.sequenceR2 = next(); opt.adapter.hasSeqR2 = true; }
        else if (a == "-f" || a == "--trim_front1") opt.trim.front1 = atoi(next()); else if (a == "-t" || a == "--trim_tail1") opt.trim.tail1 = atoi(next());
        else if (a == "-F" || a == "--trim_front2") opt.trim.front2 = atoi(next()); else if (a == "-T" || a == "--trim_tail2") opt.trim.tail2 = atoi(next());
        else if (a == "-b" || a == "--max_len1") opt.trim.maxLen1 = atoi(next()); else if (a == "-B" || a == "--max_len2") opt.trim.maxLen2 = atoi(next());
        else if (a == "-g" || a == "--trim_poly_g") opt.polyGTrim.enabled = true; else if (a == "--poly_g_min_len") opt.polyGTrim.minLen = atoi(next());
        else if (a == "-x" || a == "--trim_poly_x") opt.polyXTrim.enabled = true; else if (a == "--poly_x_min_len") opt.polyXTrim.minLen = atoi(next());
        else if (a == "-5" || a == "--cut_front") opt.qualityCut.enabledFront = true; else if (a == "-3" || a == "--cut_tail") opt.qualityCut.enabledTail = true;
        else if (a == "-r" || a == "--cut_right") opt.qualityCut.enabledRight = true;
        else if (a == "-W" || a == "--cut_window_size") { int w = atoi(next()); opt.qualityCut.windowSizeFront = opt.qualityCut.windowSizeTail = opt.qualityCut.windowSizeRight = w; }
        else if (a == "-M" || a == "--cut_mean_quality") { int q = atoi(next()); opt.qualityCut.qualityFront = opt.qualityCut.qualityTail = opt.qualityCut.qualityRight = q; }
        else if (a == "-Q" || a == "--disable_quality_filtering") opt.qualfilter.enabled = false;
        else if (a == "-q" || a == "--qualified_quality_phred") opt.qualfilter.qualifiedQual = (char)(33 + atoi(next()));
        else if (a == "-u" || a == "--unqualified_percent_limit") opt.qualfilter.unqualifiedPercentLimit = atoi(next());
        else if (a == "-n" || a == "--n_base_limit") opt.qualfilter.nBaseLimit = atoi(next()); else if (a == "-e" || a == "--average_qual") opt.qualfilter.avgQualReq = atoi(next());
        else if (a == "-L" || a == "--disable_length_filtering") opt.lengthFilter.enabled = false;
        else if (a == "-l" || a == "--length_required") opt.lengthFilter.requiredLength = atoi(next()); else if (a == "--length_limit") opt.lengthFilter.maxLength = atoi(next());
        else if (a == "-y" || a == "--low_complexity_filter") opt.complexityFilter.enabled = true;
        else if (a == "-Y" || a == "--complexity_threshold") opt.complexityFilter.threshold = std::min(100, std::max(0, atoi(next()))) / 100.0;
        else if (a == "-c" || a == "--correction") opt.correction.enabled = true;
        else if (a == "--overlap_len_require") opt.overlapRequire = atoi(next()); else if (a == "--overlap_diff_limit") opt.overlapDiffLimit = atoi(next());
        else if (a == "--overlap_diff_percent_limit") opt.overlapDiffPercentLimit = atoi(next());
        else if (a == "--device_fastq") deviceFastq = true; else if (a == "-6" || a == "--phred64") phred64 = true;
        else if (a == "--chunk_bytes") chunkBytes = (size_t)atoll(next());
        else if (a == "-D" || a == "--dedup") dedup = true; else if (a == "--dup_calc_accuracy") dupLevel = std::min(6, std::max(1, atoi(next())));
        else if (a == "--dont_eval_duplication") evalDup = false;
        else if (a == "-z" || a == "--compression") zlevel = std::min(9, std::max(1, atoi(next()))); else if (a == "--zthreads") zthreads = std::max(1, atoi(next()));
        else if (a == "--max_read_len") maxLen = atoi(next()); else if (a == "--pack_size") packSize = atoi(next());
        else { fprintf(stderr, "unknown flag %s\n", a.c_str()); return 2; }
    }
    if (in1.empty()) { fprintf(stderr, "usage: fastp_gpu_cli -i R1.fq [-I R2.fq] [-o out1.fq] [-O out2.fq] [-j summary.json] [fastp flags]\n"); return 2; }
    opt.paired = !in2.empty();
    std::ifstream f1(in1), f2;
    if (opt.paired) f2.open(in2);
    if (!f1 || (opt.paired && !f2)) { fprintf(stderr, "cannot open input\n"); return 1; }
    if (!deviceFastq && ((in1.size() > 3 && in1.compare(in1.size() - 3, 3, ".gz") == 0))) { fprintf(stderr, "compressed input needs --device_fastq\n"); return 2; }
    if (maxLen == 0) {                                  /* Evaluator::evaluateSeqLen peeks at the first records (src/evaluator.cpp:54-76) */
        std::vector<uint8_t> head(1 << 20);
        void* gp = fp_gz_open(in1.c_str());
        const int64_t got = gp ? fp_gz_read(gp, head.data(), (int64_t)head.size()) : 0;
        fp_gz_close(gp);
        maxLen = 151;
        int n = 0; size_t ls = 0;
        for (size_t i = 0; i < (size_t)std::max<int64_t>(got, 0) && n < 4000; i++)
            if (head[i] == '\n') { if (n % 4 == 1) maxLen = std::max(maxLen, (int)(i - ls)); n++; ls = i + 1; }
        maxLen += 64;
    }
    if (chunkBytes == 0) chunkBytes = (size_t)packSize * (size_t)(2 * maxLen + 64);
    GpuChainWorker worker(&opt, maxLen, 0, packSize);
    if (!worker.ok()) { fprintf(stderr, "fastp_gpu_cli: %s\n", worker.error().c_str()); return 1; }
    std::ofstream o1, o2;
    if (!out1.empty()) o1.open(out1);
    if (!out2.empty()) o2.open(out2);
    if (deviceFastq) {
        if (evalDup || dedup) {                            /* accuracy level: 3 with --dedup, else 1, unless given (src/main.cpp:203-209) */
            if (!worker.setDedup(dupLevel ? dupLevel : (dedup ? 3 : 1), dedup)) { fprintf(stderr, "fastp_gpu_cli: duplicate filter: %s\n", fp_last_error()); return 1; }
        }
        /* text path: raw file chunks go to the device, which parses, filters and re-encodes them (fp_fastq_process_host);
           whatever a chunk's last, incomplete record (or the longer side of a pair) leaves over is carried into the next chunk */
        /* .gz / BGZF inputs are inflated on the host (fp_gz_open: zlib streaming reader, plain files pass through), .gz outputs are written as
           one gzip member per round, compressed by `zthreads` host threads (what the reference's writer threads do per pack,
           src/writerthread.cpp:118-168) */
        std::string buf1, buf2;
        bool eof1 = false, eof2 = !opt.paired;
        void* g1 = fp_gz_open(in1.c_str()); void* g2 = opt.paired ? fp_gz_open(in2.c_str()) : nullptr;
        if (!g1 || (opt.paired && !g2)) { fprintf(stderr, "cannot open input\n"); return 1; }
        auto fill = [&](void* g, std::string& buf, bool& eof) {
            if (eof) return;
            const size_t old = buf.size();
            buf.resize(old + chunkBytes);
            const int64_t got = fp_gz_read(g, reinterpret_cast<uint8_t*>(&buf[old]), (int64_t)chunkBytes);
            if (got < 0) { fprintf(stderr, "fastp_gpu_cli: corrupt compressed input\n"); exit(1); }
            buf.resize(old + (size_t)got);
            if ((size_t)got < chunkBytes) eof = true;
        };
        auto gz_name = [](const std::string& n) { return n.size() > 3 && n.compare(n.size() - 3, 3, ".gz") == 0; };
        const bool zout1 = gz_name(out1), zout2 = gz_name(out2);
        std::vector<uint8_t> zbuf;
        auto emit = [&](std::ofstream& o, const std::string& t, bool z) {
            if (!o.is_open() || t.empty()) return;
            if (!z) { o << t; return; }
            int64_t nz = 0;
            zbuf.resize((size_t)fp_gz_deflate_bound((int64_t)t.size(), 1 << 20));
            if (fp_gz_deflate(reinterpret_cast<const uint8_t*>(t.data()), (int64_t)t.size(), zbuf.data(), (int64_t)zbuf.size(), &nz, 1 << 20, zlevel, zthreads) != FP_OK) {
                fprintf(stderr, "fastp_gpu_cli: gzip output failed\n"); exit(1);
            }
            o.write(reinterpret_cast<const char*>(zbuf.data()), (std::streamsize)nz);
        };
        for (;;) {
            fill(g1, buf1, eof1);
            if (opt.paired) fill(g2, buf2, eof2);
            const bool final = eof1 && eof2;
            std::string s1, s2; size_t c1 = 0, c2 = 0; long units = 0;
            if (!worker.processFastqText(buf1.data(), buf1.size(), buf2.data(), buf2.size(), final, phred64, &s1, &s2, &c1, &c2, &units)) {
                fprintf(stderr, "fastp_gpu_cli: %s\n", worker.error().c_str()); return 1;
            }
            emit(o1, s1, zout1);
            emit(o2, s2, zout2);
            buf1.erase(0, c1); if (opt.paired) buf2.erase(0, c2);
            if (worker.inputEnded()) break;                    /* a reader gave up on a record: the reference stops reading there */
            if (final && (units == 0 || (buf1.empty() && buf2.empty()))) break;
            if (final && c1 == 0 && c2 == 0) break;
        }
        fp_gz_close(g1); fp_gz_close(g2);
    } else
    for (;;) {
        ReadPack* lp = new ReadPack{new Read*[packSize], 0};
        ReadPack* rp = opt.paired ? new ReadPack{new Read*[packSize], 0} : nullptr;
        while (lp->count < packSize) {
            Read* a = readRecord(f1);
            if (!a) break;
            if (opt.paired) { Read* b = readRecord(f2); if (!b) { delete a; break; } rp->data[rp->count++] = b; }
            lp->data[lp->count++] = a;
        }
        const bool last = lp->count < packSize;
        std::string s1, s2;
        if (lp->count == 0) { delete[] lp->data; delete lp; if (rp) { delete[] rp->data; delete rp; } break; }
        if (opt.paired) worker.processPairEnd(lp, rp, &s1, &s2); else worker.processSingleEnd(lp, &s1);
        if (!worker.error().empty()) { fprintf(stderr, "fastp_gpu_cli: %s\n", worker.error().c_str()); return 1; }
        if (o1.is_open()) o1 << s1;
        if (o2.is_open()) o2 << s2;
        if (last) break;
    }
    Stats pre1, post1, pre2, post2; FilterResult fr; std::vector<long> isize;
    if (!worker.finish(&pre1, &post1, &pre2, &post2, &fr, &isize)) { fprintf(stderr, "fastp_gpu_cli: %s\n", worker.error().c_str()); return 1; }
    long dupTotal = 0, dupCount = 0;
    if (deviceFastq && (evalDup || dedup)) worker.dupTotals(&dupTotal, &dupCount);
    if (!json.empty()) {
        std::ofstream js(json);
        auto tot = [&](long Stats::*m) { return pre1.*m + (opt.paired ? pre2.*m : 0); };
        auto tota = [&](long Stats::*m) { return post1.*m + (opt.paired ? post2.*m : 0); };
        js << "{\n \"before_filtering\": {\"total_reads\": " << tot(&Stats::mReads) << ", \"total_bases\": " << tot(&Stats::mBases)
           << ", \"q20_bases\": " << tot(&Stats::mQ20Total) << ", \"q30_bases\": " << tot(&Stats::mQ30Total) << "},\n"
           << " \"after_filtering\": {\"total_reads\": " << tota(&Stats::mReads) << ", \"total_bases\": " << tota(&Stats::mBases)
           << ", \"q20_bases\": " << tota(&Stats::mQ20Total) << ", \"q30_bases\": " << tota(&Stats::mQ30Total) << "},\n"
           << " \"filtering_result\": {\"passed_filter_reads\": " << fr.mFilterReadStats[FP_PASS_FILTER] << ", \"low_quality_reads\": " << fr.mFilterReadStats[FP_FAIL_QUALITY]
           << ", \"too_many_N_reads\": " << fr.mFilterReadStats[FP_FAIL_N_BASE] << ", \"too_short_reads\": " << fr.mFilterReadStats[FP_FAIL_LENGTH]
           << ", \"too_long_reads\": " << fr.mFilterReadStats[FP_FAIL_TOO_LONG] << ", \"low_complexity_reads\": " << fr.mFilterReadStats[FP_FAIL_COMPLEXITY]
           << ", \"adapter_dimer_reads\": " << fr.mFilterReadStats[FP_FAIL_ADAPTER_DIMER] << "},\n"
           << " \"adapter_cutting\": {\"adapter_trimmed_reads\": " << fr.mTrimmedAdapterRead << ", \"adapter_trimmed_bases\": " << fr.mTrimmedAdapterBases << "},\n"
           << " \"duplication\": {\"total\": " << dupTotal << ", \"duplicates\": " << dupCount << "},\n"
           << " \"corrected_reads\": " << fr.mCorrectedReads << ",\n \"insert_size_unknown\": " << (isize.empty() ? 0 : isize.back()) << "\n}\n";
    }
    return 0;
}
