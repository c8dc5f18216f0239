"""-m gpu: the CUDA hot path against the committed golden fixtures generated from the REFERENCE build
(tests/golden/, no oracle in the loop), plus the edge cases the reference tests (empty / ragged reads,
the 9-record testdata with an empty record) and size-independent properties at larger sizes."""
import os

import numpy as np
import pytest

import fp_testlib as T
from fastp_b200 import capi
from test_golden import GOLDEN, load_fixture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (no CPU fallback exists)")
    import fp_gpu
    return fp_gpu


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_reproduces_reference_golden(gpu, path):
    meta, p, arrs, want = load_fixture(path)
    got = gpu.run_gpu(p, arrs, meta["cycles"], mode="device")
    T.assert_results_equal(got, want, meta["paired"], skip=("adapter_pos",), what=os.path.basename(path))


def test_empty_batch_and_edge_reads(gpu):
    p = T.config_params("cfg4_full", 1)
    reads1 = [("", ""), ("A", "I"), ("ACGT", "IIII"), ("N" * 40, "#" * 40), ("G" * 150, "I" * 150), ("ACGT" * 37 + "AC", "5" * 150)]
    reads2 = [("ACGT", "IIII"), ("", ""), ("T", "#"), ("N" * 40, "I" * 40), ("C" * 150, "I" * 150), ("GT" + "ACGT"[::-1] * 37, "?" * 150)]
    _, arrs = capi.batch_from_strings(reads1, reads2, stride=160)
    want = T.run_cpu("oracle", p, arrs, 160)
    got = gpu.run_gpu(p, arrs, 160, mode="device")
    T.assert_results_equal(got, want, 1, what="edge")
    got = gpu.run_gpu(p, arrs, 160, mode="host")
    T.assert_results_equal(got, want, 1, what="edge-host")
    # n = 0
    _, arrs0 = capi.host_batch(0, 160, 1)
    got = gpu.run_gpu(p, arrs0, 160, mode="host")
    assert got["counters"].data.sum() == 0


def test_non_acgtn_bytes_take_the_exact_path(gpu):
    """lower-case / IUPAC bases are outside the fast path's alphabet; the device must still match the reference port."""
    p = T.config_params("cfg4_full", 1)
    _, arrs = T.synth_host(2000, 160, 1, 0, 3, 1, 150)
    rng = np.random.default_rng(0)
    for k in ("seq1", "seq2"):
        rows = rng.integers(0, 2000, 300)
        cols = rng.integers(0, 150, 300)
        arrs[k][rows, cols] = rng.choice(np.frombuffer(b"acgtnRYKM.", np.uint8), 300)
        for r in range(2000):                       # keep padding zero
            arrs[k][r, arrs["len" + k[-1]][r]:] = 0
    want = T.run_cpu("oracle", p, arrs, 160)
    got = gpu.run_gpu(p, arrs, 160, mode="device")
    T.assert_results_equal(got, want, 1, what="non-ACGTN")


@pytest.mark.parametrize("paired", [1, 0])
def test_large_batch_properties(gpu, paired):
    """Size-independent properties at a size the scalar oracle would take minutes for (2M units): shard additivity
    (counters of the whole == sum of counters of two halves), verdict/read conservation, idempotence of a re-run."""
    import ctypes as C
    import torch
    n = 2_000_000
    p = T.config_params("cfg4_full", paired)
    ctx = gpu.GpuCtx(p, n, 160, 160)
    lib = ctx.lib

    def alloc(nbytes):
        return torch.empty(nbytes, dtype=torch.uint8, device="cuda:0")
    t = {"seq1": alloc(n * 160), "qual1": alloc(n * 160), "len1": alloc(n * 2)}
    if paired:
        t.update(seq2=alloc(n * 160), qual2=alloc(n * 160), len2=alloc(n * 2))
    o1, o2, ov = alloc(n * 16), alloc(n * 16), alloc(n * 8)

    def batch(lo, hi):
        b = capi.Batch()
        b.n, b.stride = hi - lo, 160
        for k, v in t.items():
            setattr(b, k, v.data_ptr() + lo * (2 if k.startswith("len") else 160))
        return b

    def run(lo, hi):
        ctx.reset()
        b = batch(lo, hi)
        if paired:
            capi.check(lib.fp_process_pe(ctx.h, C.byref(b), o1.data_ptr() + lo * 16, o2.data_ptr() + lo * 16, ov.data_ptr() + lo * 8,
                                         None, 0, None, None), lib)
        else:
            capi.check(lib.fp_process_se(ctx.h, C.byref(b), o1.data_ptr() + lo * 16, None), lib)
        return ctx.counters().data.copy()
    full = batch(0, n)
    capi.check(lib.fp_synth_fill(ctx.h, C.byref(full), 0, 42, 1, 150, None), lib)
    torch.cuda.synchronize()
    first = run(0, n)          # applies base corrections in place
    whole = run(0, n)
    for _ in range(6):         # base correction rewrites rows in place: iterate to its fixed point (idempotence)
        again = run(0, n)
        if (whole == again).all():
            break
        whole = again
    assert (whole == again).all()
    halves = run(0, n // 2) + run(n // 2, n)
    assert (halves == whole).all()
    cv = capi.CounterView(ctx.L, whole)
    assert cv.stats(capi.STATS_PRE1)["reads"] == n
    assert int(cv.filter[:32].sum()) == n * (2 if paired else 1)
    assert int(cv.stats(capi.STATS_PRE1)["qualhist"].sum()) == cv.stats(capi.STATS_PRE1)["length_sum"]
    assert int(cv.stats(capi.STATS_POST1)["cycle"][32].sum()) == cv.stats(capi.STATS_POST1)["length_sum"]
    assert cv.stats(capi.STATS_POST1)["reads"] == int(cv.filter[0]) // (2 if paired else 1)
    # records: pair verdict histogram == counter block
    rec = o1.cpu().numpy().view(capi.READ_RESULT_DTYPE)
    hist = np.bincount(rec["pair_verdict"], minlength=32)[:32] * (2 if paired else 1)
    assert (hist == cv.filter[:32]).all()
    # first pass differs from the later ones only through base correction (PE) -- verdict totals are conserved
    assert int(capi.CounterView(ctx.L, first).filter[:32].sum()) == n * (2 if paired else 1)
    ctx.close()
