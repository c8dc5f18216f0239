"""Known-answer tests that pin the CPU oracle (oracle/fastp_oracle.c) to the reference's OWN unit-test
vectors (src/filter.cpp:245-264, src/polyx.cpp:118-130, src/adaptertrimmer.cpp:159-185,
src/basecorrector.cpp:85-106, src/overlapanalysis.cpp:181-210) and to the extra vectors probed from the
reference build."""
import ctypes as C

import numpy as np

import fp_testlib as T
from fastp_b200 import capi

S40 = "ACGTTGCAACGTTGCAACGTTGCAACGTTGCAACGTTGCA"
TRUSEQ_R1, TRUSEQ_R2 = T.TRUSEQ_R1, T.TRUSEQ_R2


def trim_and_cut(seq, qual, front, tail, **kw):
    p = capi.default_params(0, lib=T.oracle(), **kw)
    s = C.create_string_buffer(seq.encode(), len(seq) + 1)
    q = C.create_string_buffer(qual.encode(), len(qual) + 1)
    f, ln = C.c_int(), C.c_int()
    ok = T.oracle().fp_oracle_trim_and_cut(C.byref(p), s, q, len(seq), front, tail, C.byref(f), C.byref(ln))
    return (f.value, ln.value) if ok else None


def rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def test_filter_test_vector():
    # Filter::test: cut_front + cut_tail, W=4, Q20, tail=1 -> exact strings
    seq = "TTTTAACCCCCCCCCCCCCCCCCCCCCCCCCCCCAATTTT"
    qual = "/////CCCCCCCCCCCC////CCCCCCCCCCCCCC////E"
    f, ln = trim_and_cut(seq, qual, 0, 1, cut_front=1, cut_tail=1)
    assert seq[f:f + ln] == "CCCCCCCCCCCCCCCCCCCCCCCCCCCC"
    assert qual[f:f + ln] == "CCCCCCCCCCC////CCCCCCCCCCCCC"


def test_polyx_test_vector():
    seq = "ATTTTAAAAAAAAAATAAAAAAAAAAAAACAAAAAAAAAAAAAAAAAAAAAAAAAT"
    s = C.create_string_buffer(seq.encode(), len(seq) + 1)
    poly, plen = C.c_int(), C.c_int()
    ln = T.oracle().fp_oracle_trim_polyx(s, len(seq), 10, C.byref(poly), C.byref(plen))
    assert seq[:ln] == "ATTTT" and plen.value == 51 and poly.value == 0


def test_adaptertrimmer_test_vector():
    seq = "TTTTAACCCCCCCCCCCCCCCCCCCCCCCCCCCCAATTTTAAAATTTTCCCCGGGG"
    s = C.create_string_buffer(seq.encode(), len(seq) + 1)
    tr = C.c_int()
    ln = T.oracle().fp_oracle_trim_by_sequence(s, len(seq), b"TTTTCCACGGGGATACTACTG", C.byref(tr))
    assert tr.value == 1 and seq[:ln] == "TTTTAACCCCCCCCCCCCCCCCCCCCCCCCCCCCAATTTTAAAA"
    # trimByMultiSequences: 4 adapters applied in turn
    seq = "TTTTAACCCCCCCCCCCCCCCCCCCCCCCCCCCCAATTTTAAAATTTTCCCCGGGGAAATTTCCCGGGAAATTTCCCGGGATCGATCGATCGATCGAATTCC"
    s = C.create_string_buffer(seq.encode(), len(seq) + 1)
    ln = len(seq)
    for ad in (b"GCTAGCTAGCTAGCTA", b"AAATTTCCCGGGAAATTTCCCGGG", b"ATCGATCGATCGATCG", b"AATTCCGGAATTCCGG"):
        ln = T.oracle().fp_oracle_trim_by_sequence(s, ln, ad, C.byref(tr))
    assert seq[:ln] == "TTTTAACCCCCCCCCCCCCCCCCCCCCCCCCCCCAATTTTAAAATTTTCCCCGGGG"


def analyze(r1, r2, dl, req, pct):
    a = C.create_string_buffer(r1.encode(), len(r1) + 1)
    b = C.create_string_buffer(r2.encode(), len(r2) + 1)
    return T.oracle().fp_oracle_analyze(a, len(r1), b, len(r2), dl, req, pct)


def test_overlapanalysis_test_vectors():
    r1 = "CAGCGCCTACGGGCCCCTTTTTCTGCGCGACCGCGTGGCTGTGGGCGCGGATGCCTTTGAGCGCGGTGACTTCTCACTGCGTATCGAGC"
    r2 = "ACCTCCAGCGGCTCGATACGCAGTGAGAAGTCACCGCGCTCAAAGGCATCCGCGCCCACAGCCACGCGGTCGCGCAGAAAAAGGGGTCC"
    ov = analyze(r1, r2, 2, 30, 0.2)
    assert (ov.overlapped, ov.offset, ov.overlap_len, ov.diff) == (1, 10, 79, 1)
    # late mismatch: 50 matching + 30 mismatching bases -> still overlapped, diff 30 (protected-prefix rule)
    late1 = "A" * 50 + "C" * 30
    late2 = rc("A" * 50 + "G" * 30)
    ov = analyze(late1, late2, 0, 30, 0.0)
    assert (ov.overlapped, ov.offset, ov.overlap_len, ov.diff) == (1, 0, 80, 30)


def test_basecorrector_test_vector():
    r1 = "TTTTAACCCCCCCCCCCCCCCCCCCCCCCCCCCCAATTTTAAAATTTTCCACGGGG"
    q1 = "EEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEE/EEEEE"
    r2 = "AAAAAAAAAACCCCGGGGAAAATTTTAAAATTGGGGGGGGGGTGGGGGGGGGGGGG"
    q2 = "EEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEE/EEEEEEEEEEEEE"
    p = capi.default_params(1, lib=T.oracle(), correction_enabled=1, adapter_enabled=0, qual_filter_enabled=0,
                            length_filter_enabled=0)
    b, arrs = capi.batch_from_strings([(r1, q1)], [(r2, q2)])
    res = T.run_cpu("oracle", p, arrs, arrs["seq1"].shape[1])
    a = res["arrs"]
    n = len(r1)
    assert bytes(a["seq1"][0, :n]).decode() == "TTTTAACCCCCCCCCCCCCCCCCCCCCCCCCCCCAATTTTAAAATTTTCCCCGGGG"
    assert bytes(a["seq2"][0, :n]).decode() == "AAAAAAAAAACCCCGGGGAAAATTTTAAAATTGGGGGGGGGGGGGGGGGGGGGGGG"
    assert bytes(a["qual1"][0, :n]).decode() == "E" * n and bytes(a["qual2"][0, :n]).decode() == "E" * n
    fr = res["counters"].filter
    # only the diagonal of the correction matrix is ever incremented: C->C:1, G->G:1, 2 reads
    assert fr[capi.FR_CORRECTION + 3 * 8 + 3] == 1 and fr[capi.FR_CORRECTION + 7 * 8 + 7] == 1
    assert fr[capi.FR_CORRECTION: capi.FR_CORRECTION + 64].sum() == 2 and fr[capi.FR_CORRECTED_READS] == 2


def test_appendix_d_trim_and_cut():
    I, H = "I", "#"
    assert trim_and_cut(S40, H * 6 + I * 34, 0, 0, cut_front=1) == (7, 33)                                  # V1
    assert trim_and_cut(S40, I * 20 + "####IIII" + H * 12, 0, 0, cut_right=1) == (0, 20)                    # V2
    assert trim_and_cut(S40, I * 32 + H * 8, 0, 0, cut_tail=1) == (0, 31)                                   # V3
    assert trim_and_cut(S40, "####" + I * 24 + "####IIII####", 2, 3, cut_front=1, cut_right=1) == (5, 23)   # V4
    assert trim_and_cut(S40, H * 40, 0, 0, cut_right=1) is None                                             # V5
    assert trim_and_cut(S40, I * 40, 3, 5) == (3, 32)                                                       # V6


def test_appendix_d_poly_and_adapter():
    o = T.oracle()
    seq = "ACGT" * 5 + "TTTTTTTTTTTATTTTTTTT"
    s = C.create_string_buffer(seq.encode(), len(seq) + 1)
    poly, plen = C.c_int(), C.c_int()
    assert o.fp_oracle_trim_polyx(s, len(seq), 10, C.byref(poly), C.byref(plen)) == 19 and plen.value == 21 and poly.value == 1   # V7
    seq = "ACGT" * 5 + "AC" + "GGGGGGGGGAGGGGGGGG"
    s = C.create_string_buffer(seq.encode(), len(seq) + 1)
    assert o.fp_oracle_trim_polyg(s, len(seq), 10) == 22                                                                        # V8
    seq = "ACGT" * 5 + "G" * 12
    s = C.create_string_buffer(seq.encode(), len(seq) + 1)
    assert o.fp_oracle_trim_polyg(s, len(seq), 10) == 18                                                                        # 8(c) probe
    tr = C.c_int()
    seq = "ACGTTGCAACGTTGCAACGTTGCAACGT" + "AGATCGGTAGAGCACACG"
    s = C.create_string_buffer(seq.encode(), len(seq) + 1)
    assert o.fp_oracle_trim_by_sequence(s, len(seq), TRUSEQ_R1.encode(), C.byref(tr)) == 28 and tr.value == 1                   # V9
    seq = "GATCGGAAGAGCACACGTCTGAACTCCAGTCACGTT"
    s = C.create_string_buffer(seq.encode(), len(seq) + 1)
    assert o.fp_oracle_trim_by_sequence(s, len(seq), TRUSEQ_R1.encode(), C.byref(tr)) == 0 and tr.value == 1                    # V10
    seq = "ACGTTGCAACGTTGCAACGTTGCAACGTTGCAACGTTGCAAC"
    s = C.create_string_buffer(seq.encode(), len(seq) + 1)
    assert o.fp_oracle_trim_by_sequence(s, len(seq), TRUSEQ_R1.encode(), C.byref(tr)) == 42 and tr.value == 0                   # V11


def lcg_fragment(n=100):
    x, out = 12345, []
    for _ in range(n):
        x = (x * 1103515245 + 12345) & 0xFFFFFFFF
        out.append("ACGT"[(x >> 16) & 3])
    return "".join(out)


def test_appendix_d_overlap():
    frag = "ACGGTCATTGCAGTCCATGAAGCTTGACCTGAATCGGTAC"
    r1, r2 = frag + TRUSEQ_R1[:20], rc(frag) + TRUSEQ_R2[:20]
    ov = analyze(r1, r2, 5, 30, 0.2)
    assert (ov.overlapped, ov.offset, ov.overlap_len, ov.diff) == (1, -20, 40, 0)                            # V12
    p = capi.default_params(1, lib=T.oracle(), qual_filter_enabled=0, length_filter_enabled=0)
    _, arrs = capi.batch_from_strings([(r1, "I" * 60)], [(r2, "I" * 60)])
    res = T.run_cpu("oracle", p, arrs, 64)
    assert res["out1"]["len"][0] == 40 and res["out2"]["len"][0] == 40
    f = lcg_fragment()
    ov = analyze(f[:60], rc(f)[:60], 5, 30, 0.2)
    assert ov.overlapped == 0 and ov.offset == 0 and ov.overlap_len == 0                                    # V13
    ov = analyze(f[:80], rc(f)[:80], 5, 30, 0.2)
    assert (ov.overlapped, ov.offset, ov.overlap_len, ov.diff) == (1, 20, 60, 0)                            # V14


def test_appendix_d_pass_filter_and_limits():
    p = capi.default_params(0, lib=T.oracle())
    pf = lambda s, q: T.oracle().fp_oracle_pass_filter(C.byref(p), s.encode(), q.encode(), len(s))  # noqa: E731
    assert pf("ACGTACGTACGTAC" + "NNNNNN", "I" * 20) == capi.FAIL_N_BASE                                   # V15a
    assert pf("ACGTACGTACGTACGTACGT", "#" * 9 + "I" * 11) == capi.FAIL_QUALITY                             # V15b
    assert pf("ACGTACGTACGTAC", "I" * 14) == capi.FAIL_LENGTH                                              # V15c
    assert pf("ACGTACGTACGTACGTACGT", "#" * 8 + "I" * 12) == capi.PASS_FILTER                              # V15d
    # (int)(ol*0.2) for ol = 5,15,35,45,55,95,149
    assert [int(ol * (20 / 100.0)) for ol in (5, 15, 35, 45, 55, 95, 149)] == [1, 3, 7, 9, 11, 19, 29]


def test_appendix_d_stat_read():
    p = capi.default_params(0, lib=T.oracle(), qual_filter_enabled=0, length_filter_enabled=0, adapter_enabled=0, seq_len1=20)
    _, arrs = capi.batch_from_strings([("ACGTNACGTACGTTTTTTTT", "IIIII55555?????#####")])
    res = T.run_cpu("oracle", p, arrs, 32)
    s = res["counters"].summary(capi.STATS_PRE1)
    assert s == {"reads": 1, "bases": 20, "q20": 15, "q30": 10, "cycles": 20}                               # V16
    st = res["counters"].stats(capi.STATS_PRE1)
    assert st["qualhist"][ord("I")] == 5
    gc = st["cycle"][16 + (ord("G") & 7)].sum() + st["cycle"][16 + (ord("C") & 7)].sum()
    assert gc == 6
