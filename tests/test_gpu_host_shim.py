"""-m gpu: the drop-in boundary end to end.  Plain FASTQ -> fastp_gpu_cli (C++ GpuChainWorker::processPairEnd /
processSingleEnd: stage -> C-ABI -> unstage) -> output FASTQ + summary, compared with
  * the UNMODIFIED reference CLI's outputs on the reference's own testdata (committed golden, config 1), and
  * the oracle on a synthetic batch with the full option set (byte-identical output reads: the reference's own
    definition of parity, scripts/bench_e2e.sh:183-225)."""
import json
import os
import subprocess

import numpy as np
import pytest

import fp_testlib as T
from fastp_b200 import capi
from test_golden import load_fixture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "fastp_b200", "host", "fastp_gpu_cli")


def write_fastq(path, seq, qual, lens):
    with open(path, "w") as f:
        for i in range(seq.shape[0]):
            n = int(lens[i])
            f.write(f"@r{i}\n{bytes(seq[i, :n]).decode()}\n+\n{bytes(qual[i, :n]).decode()}\n")


def read_fastq(path):
    lines = open(path).read().split("\n")
    return [[lines[i + 1], lines[i + 3]] for i in range(0, len(lines) - 3, 4)]


@pytest.fixture(scope="module")
def cli():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required")
    assert os.path.exists(CLI), "build with __graft_entry__.build()"
    return CLI


def test_testdata_equals_reference_cli(cli, tmp_path):
    meta, p, arrs, want = load_fixture(os.path.join(ROOT, "tests", "golden", "testdata_pe.npz"))
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "testdata_cli.json")))
    write_fastq(tmp_path / "r1.fq", arrs["seq1"], arrs["qual1"], arrs["len1"])
    write_fastq(tmp_path / "r2.fq", arrs["seq2"], arrs["qual2"], arrs["len2"])
    subprocess.run([cli, "-i", str(tmp_path / "r1.fq"), "-I", str(tmp_path / "r2.fq"), "-o", str(tmp_path / "o1.fq"), "-O", str(tmp_path / "o2.fq"),
                    "-j", str(tmp_path / "s.json"), "-g"], check=True, timeout=300)
    assert read_fastq(tmp_path / "o1.fq") == ref["out1"]
    assert read_fastq(tmp_path / "o2.fq") == ref["out2"]
    js = json.load(open(tmp_path / "s.json"))
    for part in ("before_filtering", "after_filtering"):
        for k in ("total_reads", "total_bases", "q20_bases", "q30_bases"):
            assert js[part][k] == ref["summary"][part][k], (part, k)
    for k in ("passed_filter_reads", "low_quality_reads", "too_many_N_reads", "too_short_reads"):
        assert js["filtering_result"][k] == ref["filtering_result"][k]


@pytest.mark.parametrize("paired", [1, 0])
def test_synthetic_full_chain_output_reads(cli, tmp_path, paired):
    n = 20000
    p = T.config_params("cfg4_full", paired)
    _, arrs = T.synth_host(n, 160, paired, 0, 77, 1, 150)
    want = T.run_cpu("oracle", p, arrs, 160)
    write_fastq(tmp_path / "r1.fq", arrs["seq1"], arrs["qual1"], arrs["len1"])
    cmd = [cli, "-i", str(tmp_path / "r1.fq"), "-o", str(tmp_path / "o1.fq"), "-j", str(tmp_path / "s.json"),
           "--cut_right", "-g", "-x", "-a", T.TRUSEQ_R1, "--pack_size", "4096", "--max_read_len", "150"]
    if paired:
        write_fastq(tmp_path / "r2.fq", arrs["seq2"], arrs["qual2"], arrs["len2"])
        cmd += ["-I", str(tmp_path / "r2.fq"), "-O", str(tmp_path / "o2.fq"), "-c", "--adapter_sequence_r2", T.TRUSEQ_R2]
    subprocess.run(cmd, check=True, timeout=600)
    keep = np.nonzero(want["out1"]["pair_verdict"] == 0)[0]
    for side, key, fn in (("1", "out1", "o1.fq"), ("2", "out2", "o2.fq"))[: 2 if paired else 1]:
        got = read_fastq(tmp_path / fn)
        assert len(got) == len(keep)
        for j, i in enumerate(keep):
            r = want[key][i]
            s = bytes(want["arrs"]["seq" + side][i, r["front"]: r["front"] + r["len"]]).decode()
            q = bytes(want["arrs"]["qual" + side][i, r["front"]: r["front"] + r["len"]]).decode()
            assert got[j] == [s, q], (side, i)
    js = json.load(open(tmp_path / "s.json"))
    c = want["counters"]
    assert js["filtering_result"]["passed_filter_reads"] == c.filter[0]
    assert js["adapter_cutting"]["adapter_trimmed_reads"] == c.filter[capi.FR_ADAPTER_READS]
    assert js["adapter_cutting"]["adapter_trimmed_bases"] == c.filter[capi.FR_ADAPTER_BASES]
    posts = (capi.STATS_POST1, capi.STATS_POST2) if paired else (capi.STATS_POST1,)
    assert js["after_filtering"]["total_bases"] == sum(c.summary(s)["bases"] for s in posts)


@pytest.mark.parametrize("chunk", [0, 70001])
@pytest.mark.parametrize("paired", [1, 0])
def test_device_fastq_text_path_equals_object_path(cli, tmp_path, paired, chunk):
    """--device_fastq (raw file chunks -> device parse / filter / encode) writes the same files and summary as the object path;
    a small odd chunk size puts chunk borders inside records, names and between the sides of a pair."""
    n = 20000
    _, arrs = T.synth_host(n, 160, paired, 0, 78, 1, 150)
    write_fastq(tmp_path / "r1.fq", arrs["seq1"], arrs["qual1"], arrs["len1"])
    base = ["-i", str(tmp_path / "r1.fq"), "--cut_right", "-g", "-x", "-a", T.TRUSEQ_R1, "--pack_size", "4096", "--max_read_len", "150"]
    if paired:
        write_fastq(tmp_path / "r2.fq", arrs["seq2"], arrs["qual2"], arrs["len2"])
        base += ["-I", str(tmp_path / "r2.fq"), "-c", "--adapter_sequence_r2", T.TRUSEQ_R2]
    outs = {}
    for mode in ("obj", "text"):
        cmd = [cli] + base + ["-o", str(tmp_path / f"{mode}1.fq"), "-j", str(tmp_path / f"{mode}.json")]
        if paired:
            cmd += ["-O", str(tmp_path / f"{mode}2.fq")]
        if mode == "text":
            cmd += ["--device_fastq"] + (["--chunk_bytes", str(chunk)] if chunk else [])
        subprocess.run(cmd, check=True, timeout=600)
        outs[mode] = [(tmp_path / f"{mode}{k}.fq").read_bytes() for k in ((1, 2) if paired else (1,))]
    assert outs["text"] == outs["obj"]
    assert len(outs["text"][0]) > 1000000
    jt, jo = json.load(open(tmp_path / "text.json")), json.load(open(tmp_path / "obj.json"))
    jt.pop("duplication", None); jo.pop("duplication", None)      # only the text path runs the duplicate filter on the device
    assert jt == jo


@pytest.mark.skipif(not (T.have_ref() and os.path.exists(T.REF_CLI)), reason="oracle/_ref not built")
@pytest.mark.parametrize("paired", [1, 0])
def test_dedup_on_the_text_path_equals_reference_cli(cli, tmp_path, paired):
    """-D / --dedup: the duplicate filter runs on the device on the decoded rows (fp_dup_check, accuracy level 3 as the reference
    picks with --dedup), duplicates are dropped from the output and left out of the post-filter stats (src/peprocessor.cpp:397-401,575);
    output files and the filter / duplication numbers equal the unmodified reference CLI at --thread 1.  Also through .gz in and out."""
    import gzip
    from test_duplicate_oracle import planted
    arrs = planted(paired, n=30000, seed=91 + paired)
    write_fastq(tmp_path / "r1.fq", arrs["seq1"], arrs["qual1"], arrs["len1"])
    ref = [T.REF_CLI, "-i", str(tmp_path / "r1.fq"), "-o", str(tmp_path / "ref1.fq"), "-w", "1", "-D", "-j", str(tmp_path / "ref.json"), "-h", str(tmp_path / "ref.html")]
    gz_in = tmp_path / "r1.fq.gz"
    gz_in.write_bytes(gzip.compress((tmp_path / "r1.fq").read_bytes()[:1500000]) + gzip.compress((tmp_path / "r1.fq").read_bytes()[1500000:]))
    mine = [cli, "-i", str(gz_in), "-o", str(tmp_path / "gpu1.fq.gz"), "--device_fastq", "-D", "-j", str(tmp_path / "gpu.json"), "--pack_size", "8192", "--max_read_len", "150"]
    if paired:
        write_fastq(tmp_path / "r2.fq", arrs["seq2"], arrs["qual2"], arrs["len2"])
        ref += ["-I", str(tmp_path / "r2.fq"), "-O", str(tmp_path / "ref2.fq")]
        mine += ["-I", str(tmp_path / "r2.fq"), "-O", str(tmp_path / "gpu2.fq")]
    else:                       # single-end: the reference would auto-detect an adapter (Evaluator, host control plane) -- keep it out of the comparison
        ref += ["-A"]; mine += ["-A"]
    subprocess.run(ref, check=True, capture_output=True, cwd=tmp_path, timeout=600)
    subprocess.run(mine, check=True, timeout=600)
    assert gzip.decompress((tmp_path / "gpu1.fq.gz").read_bytes()) == (tmp_path / "ref1.fq").read_bytes()
    if paired:
        assert (tmp_path / "gpu2.fq").read_bytes() == (tmp_path / "ref2.fq").read_bytes()
    jr, jg = json.load(open(tmp_path / "ref.json")), json.load(open(tmp_path / "gpu.json"))
    for k in ("passed_filter_reads", "low_quality_reads", "too_many_N_reads", "too_short_reads"):
        assert jg["filtering_result"][k] == jr["filtering_result"][k]
    assert jg["after_filtering"]["total_reads"] == jr["summary"]["after_filtering"]["total_reads"]
    assert jg["duplication"]["duplicates"] > 100
    assert abs(jg["duplication"]["duplicates"] / jg["duplication"]["total"] - jr["duplication"]["rate"]) < 1e-6
