"""-m gpu: the REAL reference CLI with its two worker bodies bound to libfastp_b200.so (fastp_b200/host/reference_binding.cpp compiled
against the reference's own headers, oracle/Makefile target _ref/fastp_gpu) against the UNMODIFIED reference CLI (oracle/_ref/fastp_ref):
same FASTQ files, `--thread 1`; the output FASTQ files are byte-identical and the JSON reports are equal (every number: summary,
filtering_result, adapter_cutting incl. the adapter-string histograms, duplication, insert_size, per-cycle curves, k-mer counts)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import fp_testlib as T

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_CLI = os.path.join(ROOT, "oracle", "_ref", "fastp_gpu")
sys.path.insert(0, ROOT)


def _fastq_np(seq, qual, lens, tag):
    import bench
    return bench.fastq_text_np(np, seq, qual, lens, tag)


CASES = {
    "default": [],
    "full": ["--cut_right", "-g", "-x", "-c", "-a", T.TRUSEQ_R1, "--adapter_sequence_r2", T.TRUSEQ_R2],
    "cuts_failed_out": ["--cut_front", "--cut_tail", "-f", "2", "-T", "3", "-y", "-l", "30", "--failed_out", "failed.fq"],
    # merging mode (PE only): merged reads built by the reference's own OverlapAnalysis::merge from the device's records
    "merge": ["-m", "--merged_out", "m.fq", "--cut_right", "-g", "-a", T.TRUSEQ_R1, "--adapter_sequence_r2", T.TRUSEQ_R2],
    "merge_include_unmerged": ["-m", "--merged_out", "m.fq", "--include_unmerged"],
}


@pytest.mark.skipif(not (os.path.exists(GPU_CLI) and os.path.exists(T.REF_CLI)), reason="oracle/_ref/fastp_gpu not built (needs the reference sources at build time)")
@pytest.mark.parametrize("n", [3000, 400000])
@pytest.mark.parametrize("paired", [1, 0])
@pytest.mark.parametrize("case", list(CASES))
def test_bound_cli_equals_unmodified_cli(tmp_path, case, paired, n):
    if n > 3000 and case in ("cuts_failed_out", "merge_include_unmerged"):
        pytest.skip("large input covered by the other option sets")
    if case.startswith("merge") and not paired:
        pytest.skip("merging mode is paired-end only")
    flags = list(CASES[case])
    if not paired:
        for f2 in ("--adapter_sequence_r2", "-T"):
            if f2 in flags:
                i = flags.index(f2); del flags[i:i + 2]
        if "-c" in flags:
            flags.remove("-c")
    _, arrs = T.synth_host(n, 160, paired, 0, 31, 1, 150)
    _fastq_np(arrs["seq1"], arrs["qual1"], arrs["len1"], "1:N:0").tofile(tmp_path / "r1.fq")
    if paired:
        _fastq_np(arrs["seq2"], arrs["qual2"], arrs["len2"], "2:N:0").tofile(tmp_path / "r2.fq")
    outs = {}
    for tag, cli in (("ref", T.REF_CLI), ("gpu", GPU_CLI)):
        d = tmp_path / tag
        os.makedirs(d)
        cmd = [cli, "-i", str(tmp_path / "r1.fq"), "-w", "1", "-j", "t.json", "-h", "t.html"] + flags
        if case != "merge_include_unmerged":          # (--include_unmerged ignores --out1 / --out2, options.cpp:127-134)
            cmd += ["-o", "o1.fq"]
        if paired:
            cmd += ["-I", str(tmp_path / "r2.fq")] + (["-O", "o2.fq"] if case != "merge_include_unmerged" else [])
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=d, env=dict(os.environ, FASTP_B200_TRACE="1"), timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        if tag == "gpu":
            assert "units on the device" in r.stderr, "the bound CLI did not take the device path"
        js = json.load(open(d / "t.json"))
        js.pop("command", None)
        outs[tag] = (js, {f: open(d / f, "rb").read() for f in sorted(os.listdir(d)) if f.endswith(".fq")})
    assert outs["gpu"][1].keys() == outs["ref"][1].keys()
    for f in outs["ref"][1]:
        assert outs["gpu"][1][f] == outs["ref"][1][f], f"{f} differs"
    jr, jg = outs["ref"][0], outs["gpu"][0]
    assert jg.keys() == jr.keys()
    for k in jr:
        assert jg[k] == jr[k], f"JSON section {k} differs"
