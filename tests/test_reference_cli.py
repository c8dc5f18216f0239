"""Cross-checks the ORDER of operator calls restated in oracle/ref_build/ref_harness.cpp (and therefore in the
port and the CUDA path) against the UNMODIFIED reference CLI (oracle/_ref/fastp_ref, --thread 1): the same
synthetic batch is dumped as FASTQ, run through the CLI, and its JSON report + output reads are compared
with the harness counters / per-read records.  Runs only where oracle/_ref exists."""
import json
import os
import subprocess

import numpy as np
import pytest

import fp_testlib as T
from fastp_b200 import capi

pytestmark = pytest.mark.reference
needs_ref = pytest.mark.skipif(not (T.have_ref() and os.path.exists(T.REF_CLI)), reason="oracle/_ref not built")


def write_fastq(path, seq, qual, lens, tag):
    with open(path, "w") as f:
        for i in range(seq.shape[0]):
            n = int(lens[i])
            f.write(f"@SIM:1:{i} {tag}\n{bytes(seq[i, :n]).decode()}\n+\n{bytes(qual[i, :n]).decode()}\n")


def read_fastq(path):
    lines = open(path).read().split("\n")
    return [(lines[i + 1], lines[i + 3]) for i in range(0, len(lines) - 3, 4)]


CASES = {
    "default": ([], dict()),
    "full": (["--cut_right", "-g", "-x", "-c", "-a", T.TRUSEQ_R1, "--adapter_sequence_r2", T.TRUSEQ_R2],
             dict(cut_right=1, polyg_enabled=1, polyx_enabled=1, correction_enabled=1, adapter_seq_r1=T.TRUSEQ_R1, adapter_seq_r2=T.TRUSEQ_R2)),
    "cuts": (["--cut_front", "--cut_tail", "-f", "2", "-T", "3", "-y", "-l", "30"],
             dict(cut_front=1, cut_tail=1, trim_front1=2, trim_front2=2, trim_tail2=3, complexity_filter_enabled=1, length_required=30)),
}


@needs_ref
@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("paired", [1, 0])
def test_harness_order_matches_cli(tmp_path, case, paired):
    flags, kw = CASES[case]
    if not paired:
        flags = [f for f in flags]
        if "--adapter_sequence_r2" in flags:
            i = flags.index("--adapter_sequence_r2"); del flags[i:i + 2]
        if "-c" in flags:
            flags.remove("-c")
        if "-T" in flags:
            i = flags.index("-T"); del flags[i:i + 2]
        kw = {k: v for k, v in kw.items() if k not in ("adapter_seq_r2", "correction_enabled", "trim_front2", "trim_tail2")}
    n, L = 3000, 150
    _, arrs = T.synth_host(n, 160, paired, 0, 31, 1, L)
    # the CLI's reader stops at an empty record's... keep zero-length reads: fastp handles empty seq lines
    p = capi.default_params(paired, lib=T.oracle(), seq_len1=L, seq_len2=L, **kw)
    res = T.run_cpu("ref", p, arrs, 160)
    write_fastq(tmp_path / "r1.fq", arrs["seq1"], arrs["qual1"], arrs["len1"], "1:N:0")
    cmd = [T.REF_CLI, "-i", str(tmp_path / "r1.fq"), "-o", str(tmp_path / "o1.fq"), "-w", "1", "--dont_eval_duplication",
           "-j", str(tmp_path / "t.json"), "-h", str(tmp_path / "t.html")] + flags
    if paired:
        write_fastq(tmp_path / "r2.fq", arrs["seq2"], arrs["qual2"], arrs["len2"], "2:N:0")
        cmd += ["-I", str(tmp_path / "r2.fq"), "-O", str(tmp_path / "o2.fq")]
    subprocess.run(cmd, check=True, capture_output=True, cwd=tmp_path)
    js = json.load(open(tmp_path / "t.json"))
    c = res["counters"]
    sides = (capi.STATS_PRE1, capi.STATS_PRE2) if paired else (capi.STATS_PRE1,)
    posts = (capi.STATS_POST1, capi.STATS_POST2) if paired else (capi.STATS_POST1,)
    bf, af = js["summary"]["before_filtering"], js["summary"]["after_filtering"]
    for key, field in (("reads", "total_reads"), ("bases", "total_bases"), ("q20", "q20_bases"), ("q30", "q30_bases")):
        assert sum(c.summary(s)[key] for s in sides) == bf[field], (key, "before")
        assert sum(c.summary(s)[key] for s in posts) == af[field], (key, "after")
    fr = js["filtering_result"]
    assert c.filter[capi.PASS_FILTER] == fr["passed_filter_reads"]
    assert c.filter[capi.FAIL_QUALITY] == fr["low_quality_reads"]
    assert c.filter[capi.FAIL_N_BASE] == fr["too_many_N_reads"]
    assert c.filter[capi.FAIL_LENGTH] == fr["too_short_reads"]
    assert c.filter[capi.FAIL_COMPLEXITY] == fr.get("low_complexity_reads", 0)
    if "adapter_cutting" in js:
        assert c.filter[capi.FR_ADAPTER_READS] == js["adapter_cutting"]["adapter_trimmed_reads"]
        assert c.filter[capi.FR_ADAPTER_BASES] == js["adapter_cutting"]["adapter_trimmed_bases"]
    if paired:
        hist = js["insert_size"]["histogram"]
        assert list(c.isize[:len(hist)]) == hist and c.isize[-1] == js["insert_size"]["unknown"]
    # per-cycle content / quality curves of read1 before filtering
    st = c.stats(capi.STATS_PRE1)
    cyc = c.summary(capi.STATS_PRE1)["cycles"]
    mean_q = st["cycle"][33][:cyc] / np.maximum(st["cycle"][32][:cyc], 1)
    np.testing.assert_allclose(mean_q, js["read1_before_filtering"]["quality_curves"]["mean"], rtol=1e-5)
    assert int(st["kmer"].sum()) == sum(js["read1_before_filtering"]["kmer_count"].values())
    # byte-identical output reads (the reference's own parity definition, scripts/bench_e2e.sh:183-225)
    keep = np.nonzero(res["out1"]["pair_verdict"] == 0)[0]
    for side, key, fn in (("1", "out1", "o1.fq"), ("2", "out2", "o2.fq"))[: 2 if paired else 1]:
        want = read_fastq(tmp_path / fn)
        assert len(want) == len(keep)
        for j, i in enumerate(keep):
            r = res[key][i]
            s = bytes(res["arrs"]["seq" + side][i, r["front"]: r["front"] + r["len"]]).decode()
            q = bytes(res["arrs"]["qual" + side][i, r["front"]: r["front"] + r["len"]]).decode()
            assert (s, q) == want[j], (side, i)


_COMP = bytes.maketrans(b"ACGTacgt", b"TGCATGCA")


def _revcomp(s):
    return "".join(chr(c) if chr(c) in "ACGT" else "N" for c in s.encode().translate(_COMP)[::-1])


@needs_ref
@pytest.mark.parametrize("include_unmerged", [0, 1])
def test_merge_mode_matches_cli(tmp_path, include_unmerged):
    """--merge [--include_unmerged] of the unmodified CLI against the harness / records: merged reads rebuilt from the per-read records and
    fp_ov_result (fp_merged_lens) are byte-identical to --merged_out, the unmerged passing pairs to --out1/--out2, and the JSON's
    after-filtering totals are the post-filter Stats of read 1 alone (src/peprocessor.cpp:519-591)."""
    n, L = 3000, 150
    _, arrs = T.synth_host(n, 160, 1, 0, 33, 1, L)
    p = capi.default_params(1, lib=T.oracle(), seq_len1=L, seq_len2=L, merge_enabled=1, correction_enabled=1, merge_include_unmerged=include_unmerged)
    res = T.run_cpu("ref", p, arrs, 320)
    write_fastq(tmp_path / "r1.fq", arrs["seq1"], arrs["qual1"], arrs["len1"], "1:N:0")
    write_fastq(tmp_path / "r2.fq", arrs["seq2"], arrs["qual2"], arrs["len2"], "2:N:0")
    cmd = [T.REF_CLI, "-i", str(tmp_path / "r1.fq"), "-I", str(tmp_path / "r2.fq"), "-m", "--merged_out", str(tmp_path / "m.fq"), "-w", "1",
           "--dont_eval_duplication", "-j", str(tmp_path / "t.json"), "-h", str(tmp_path / "t.html")]
    cmd += ["--include_unmerged"] if include_unmerged else ["-o", str(tmp_path / "o1.fq"), "-O", str(tmp_path / "o2.fq")]
    subprocess.run(cmd, check=True, capture_output=True, cwd=tmp_path)
    js = json.load(open(tmp_path / "t.json"))
    c = res["counters"]
    af = js["summary"]["after_filtering"]
    for key, field in (("reads", "total_reads"), ("bases", "total_bases"), ("q20", "q20_bases"), ("q30", "q30_bases")):
        assert c.summary(capi.STATS_POST1)[key] == af[field], key
        assert c.summary(capi.STATS_POST2)[key] == 0
    fr = js["filtering_result"]
    assert c.filter[capi.PASS_FILTER] == fr["passed_filter_reads"]
    assert c.filter[capi.FAIL_QUALITY] == fr["low_quality_reads"]
    assert c.filter[capi.FAIL_LENGTH] == fr["too_short_reads"]
    o1, o2, ov, a = res["out1"], res["out2"], res["ov"], res["arrs"]

    def window(side, i, lo, ln):
        r = (o1 if side == "1" else o2)[i]
        return (bytes(a["seq" + side][i, r["front"] + lo: r["front"] + lo + ln]).decode(), bytes(a["qual" + side][i, r["front"] + lo: r["front"] + lo + ln]).decode())
    want_m, want_1, want_2 = [], [], []
    for i in range(n):
        if o1["flags"][i] & 0x80:                                   # FP_F_MERGED
            if o1["verdict"][i] != 0:
                continue
            len1 = int(ov["overlap_len"][i]) + max(0, int(ov["offset"][i]))
            len2 = int(o2["len"][i]) - int(ov["overlap_len"][i]) if ov["offset"][i] > 0 else 0
            s1, q1 = window("1", i, 0, len1)
            s2, q2 = window("2", i, 0, len2)
            want_m.append((s1 + _revcomp(s2), q1 + q2[::-1]))
        elif include_unmerged:
            if o1["verdict"][i] == 0 and not (o1["flags"][i] & 1):
                want_m.append(window("1", i, 0, int(o1["len"][i])))
            if o2["verdict"][i] == 0 and not (o2["flags"][i] & 1):
                want_m.append(window("2", i, 0, int(o2["len"][i])))
        elif o1["pair_verdict"][i] == 0:
            want_1.append(window("1", i, 0, int(o1["len"][i]))); want_2.append(window("2", i, 0, int(o2["len"][i])))
    assert len(want_m) > 100
    assert read_fastq(tmp_path / "m.fq") == want_m
    assert int(c.filter[capi.FR_MERGED_PAIRS] if hasattr(capi, "FR_MERGED_PAIRS") else c.filter[107]) == sum(1 for i in range(n) if (o1["flags"][i] & 0x80) and o1["verdict"][i] == 0)
    if not include_unmerged:
        assert read_fastq(tmp_path / "o1.fq") == want_1 and read_fastq(tmp_path / "o2.fq") == want_2
