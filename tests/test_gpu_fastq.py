"""-m gpu: the device FASTQ codec (fp_fastq_decode / fp_fastq_encode / fp_fastq_process_host, through the C-ABI) against the CPU
oracle of FastqReader / Read::appendToString, and the whole text-in / text-out path against the UNMODIFIED reference CLI's
output files (oracle/_ref/fastp_ref travels to the GPU box) or, where that binary is absent, committed digests."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import fp_testlib as T
from fastp_b200 import capi

pytestmark = pytest.mark.gpu
CASES = T.fastq_edge_cases()
DIGESTS = os.path.join(os.path.dirname(__file__), "golden", "fastq_cli_digests.json")


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (no CPU fallback exists)")
    import fp_gpu
    return fp_gpu


@pytest.fixture(scope="module")
def ctx512(gpu):
    c = gpu.GpuCtx(T.config_params("default", 0), 4096, 512, 512)
    yield c
    c.close()


def same_decode(a, b, what):
    assert a["info"] == b["info"], (what, a["info"], b["info"])
    for k in ("seq", "qual", "len"):
        assert np.array_equal(a[k], b[k]), (what, k)
    for f in ("name_off", "strand_off", "strand_len"):
        assert np.array_equal(a["recs"][f], b["recs"][f]), (what, f)
    assert np.array_equal(a["recs"]["name_len"] & 0x0FFFFFFF, b["recs"]["name_len"]), (what, "name_len")


@pytest.mark.parametrize("final", [1, 0])
@pytest.mark.parametrize("name", list(CASES))
def test_decode_edge_cases(gpu, ctx512, name, final):
    text = CASES[name]
    want = T.oracle_fastq_decode(text, final=final, stride=512)
    got = gpu.gpu_fastq_decode(ctx512, text, final=final)
    same_decode(got, want, f"{name}/final={final}")


@pytest.mark.parametrize("cut", [1, 3, 77, 4096])
def test_decode_every_prefix_length(gpu, ctx512, cut):
    """Chunk boundaries anywhere (inside names, between '\\r' and '\\n', ...): n_records / consumed equal the oracle's."""
    text = CASES["mixed_eol"] + CASES["crlf"] + CASES["blank_lines_between"]
    for end in range(0, len(text) + 1, cut):
        chunk = text[:end]
        for final in (0, 1):
            same_decode(gpu.gpu_fastq_decode(ctx512, chunk, final=final), T.oracle_fastq_decode(chunk, final=final, stride=512), f"prefix {end} final {final}")


def test_decode_fuzz(gpu, ctx512):
    """Random texts (mixed line ends, junk, broken records), whole and as non-final prefixes: device == oracle."""
    rng = np.random.default_rng(77)
    for k in range(150):
        text = T.fastq_fuzz_text(rng)
        cut = int(rng.integers(0, len(text) + 1))
        for chunk, final in ((text, 1), (text[:cut], 0)):
            same_decode(gpu.gpu_fastq_decode(ctx512, chunk, final=final), T.oracle_fastq_decode(chunk, final=final, stride=512), f"fuzz {k} final {final}")


def test_decode_capacity_and_phred64(gpu, ctx512):
    text = CASES["plain"]
    same_decode(gpu.gpu_fastq_decode(ctx512, text, capacity=7), T.oracle_fastq_decode(text, stride=512, capacity=7), "capacity")
    rng = np.random.default_rng(5)
    recs = []
    for i in range(300):
        n = int(rng.integers(0, 200))
        recs.append("@q%d\n%s\n+\n%s\n" % (i, "".join(rng.choice(list("ACGTN"), n)), bytes(rng.integers(59, 127, n).astype(np.uint8)).decode("latin1")))
    text = "".join(recs).encode("latin1")
    same_decode(gpu.gpu_fastq_decode(ctx512, text, phred64=1), T.oracle_fastq_decode(text, phred64=1, stride=512), "phred64")


@pytest.mark.parametrize("paired", [1, 0])
def test_large_text_decode_chain_encode(gpu, paired):
    """60 K units of enriched synthetic reads as text: device decode == oracle decode, chain, device encode == oracle encode."""
    p = T.config_params("cfg4_full", paired)
    n = 60000
    _, arrs = T.synth_host(n, 160, paired, 0, 99, 1, 150)
    ctx = gpu.GpuCtx(p, n, 160, 160)
    for side in ("1", "2")[: 2 if paired else 1]:
        text = T.fastq_text(arrs["seq" + side], arrs["qual" + side], arrs["len" + side], side + ":N:0")
        want = T.oracle_fastq_decode(text, stride=160, capacity=n)
        got = gpu.gpu_fastq_decode(ctx, text, capacity=n)
        same_decode(got, want, "large side " + side)
        assert np.array_equal(got["seq"], arrs["seq" + side]) and np.array_equal(got["len"], arrs["len" + side])
        res = T.run_cpu("oracle", p, arrs, 160)["out" + side]
        assert gpu.gpu_fastq_encode(ctx, got["dev"], res, n) == T.oracle_fastq_encode(text, want["recs"], res, want["seq"], want["qual"], 160)
    ctx.close()


CLI_CASES = {
    "default": ([], dict()),
    "full": (["--cut_right", "-g", "-x", "-c", "-a", T.TRUSEQ_R1, "--adapter_sequence_r2", T.TRUSEQ_R2],
             dict(cut_right=1, polyg_enabled=1, polyx_enabled=1, correction_enabled=1, adapter_seq_r1=T.TRUSEQ_R1, adapter_seq_r2=T.TRUSEQ_R2)),
}


def cli_inputs(case, paired, n=3000, L=150):
    flags, kw = CLI_CASES[case]
    if not paired:
        flags = list(flags)
        if "--adapter_sequence_r2" in flags:
            i = flags.index("--adapter_sequence_r2"); del flags[i:i + 2]
        if "-c" in flags:
            flags.remove("-c")
        kw = {k: v for k, v in kw.items() if k not in ("adapter_seq_r2", "correction_enabled")}
    _, arrs = T.synth_host(n, 160, paired, 0, 31, 1, L)
    p = capi.default_params(paired, lib=T.oracle(), seq_len1=L, seq_len2=L, **kw)
    t1 = T.fastq_text(arrs["seq1"], arrs["qual1"], arrs["len1"], "1:N:0")
    t2 = T.fastq_text(arrs["seq2"], arrs["qual2"], arrs["len2"], "2:N:0") if paired else None
    return flags, p, t1, t2


def run_cli(tmp_path, flags, t1, t2):
    (tmp_path / "r1.fq").write_bytes(t1)
    cmd = [T.REF_CLI, "-i", str(tmp_path / "r1.fq"), "-o", str(tmp_path / "o1.fq"), "-w", "1", "--dont_eval_duplication",
           "-j", str(tmp_path / "t.json"), "-h", str(tmp_path / "t.html")] + flags
    if t2 is not None:
        (tmp_path / "r2.fq").write_bytes(t2)
        cmd += ["-I", str(tmp_path / "r2.fq"), "-O", str(tmp_path / "o2.fq")]
    subprocess.run(cmd, check=True, capture_output=True, cwd=tmp_path)
    outs = [(tmp_path / "o1.fq").read_bytes()]
    if t2 is not None:
        outs.append((tmp_path / "o2.fq").read_bytes())
    return outs


@pytest.mark.parametrize("paired", [1, 0])
@pytest.mark.parametrize("case", list(CLI_CASES))
def test_text_in_text_out_equals_reference_cli(gpu, tmp_path, case, paired):
    """The reference's own parity definition (scripts/bench_e2e.sh:183-225): byte-identical output FASTQ files."""
    flags, p, t1, t2 = cli_inputs(case, paired)
    ctx = gpu.GpuCtx(p, 4096, 160, 160)
    got = gpu.gpu_fastq_process_host(ctx, t1, t2)
    ctx.close()
    outs = [got["out1"]] + ([got["out2"]] if paired else [])
    key = f"{case}/{'pe' if paired else 'se'}"
    if os.path.exists(T.REF_CLI):
        want = run_cli(tmp_path, flags, t1, t2)
        assert [len(x) for x in outs] == [len(x) for x in want]
        assert outs == want, key
    digests = json.load(open(DIGESTS))
    assert [hashlib.md5(x).hexdigest() for x in outs] == digests[key], key
    assert got["n"] == 3000 and got["consumed"][0] == len(t1)
