"""-m gpu parity tests proper: the CUDA hot path (through the C-ABI) against the CPU oracle on the same
seeded inputs -- bit-exact per-read records, overlap records, corrected bases and every counter."""
import numpy as np
import pytest

import fp_testlib as T
from fastp_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (no CPU fallback exists)")
    import fp_gpu
    return fp_gpu


@pytest.mark.parametrize("paired", [1, 0])
@pytest.mark.parametrize("name", T.CONFIG_NAMES)
def test_enriched_device_mode(gpu, name, paired):
    p = T.config_params(name, paired)
    _, arrs = T.synth_host(6000, 160, paired, 1000, 42, 1, 150)
    want = T.run_cpu("oracle", p, arrs, 160)
    got = gpu.run_gpu(p, arrs, 160, mode="device")
    T.assert_results_equal(got, want, paired, what=f"{name}/{'PE' if paired else 'SE'}")


@pytest.mark.parametrize("paired", [1, 0])
@pytest.mark.parametrize("name", ["cfg4_full", "cfg3_overlap_correction", "default"])
def test_enriched_host_mode(gpu, name, paired):
    p = T.config_params(name, paired)
    _, arrs = T.synth_host(5000, 160, paired, 77, 43, 1, 150)
    want = T.run_cpu("oracle", p, arrs, 160)
    got = gpu.run_gpu(p, arrs, 160, mode="host")
    T.assert_results_equal(got, want, paired, what=f"host {name}")


@pytest.mark.parametrize("mode", ["device", "host"])
@pytest.mark.parametrize("L,stride", [(150, 160), (100, 112), (250, 256)])
@pytest.mark.parametrize("name", T.GAP_CONFIG_NAMES)
def test_gap_overlap_passes(gpu, name, L, stride, mode):
    """--allow_gap_overlap_trimming (analyze allowGap passes + diffWithOneInsertion) on reads with single-base indels."""
    if mode == "host" and L != 150:
        pytest.skip("host mode covered at L=150")
    p = T.config_params(name, 1)
    _, arrs = T.synth_host(6000, stride, 1, 300, 31, 2, L)
    want = T.run_cpu("oracle", p, arrs, stride)
    got = gpu.run_gpu(p, arrs, stride, mode=mode)
    T.assert_results_equal(got, want, 1, what=f"{name}/L{L}/{mode}")


@pytest.mark.parametrize("paired", [1, 0])
def test_ref_style_profile(gpu, paired):
    p = T.config_params("cfg4_full", paired)
    _, arrs = T.synth_host(4000, 160, paired, 0, 42, 0, 150)
    want = T.run_cpu("oracle", p, arrs, 160)
    got = gpu.run_gpu(p, arrs, 160, mode="device")
    T.assert_results_equal(got, want, paired, what="ref-style")


@pytest.mark.parametrize("paired", [1, 0])
def test_len250_stride256(gpu, paired):
    p = T.config_params("cfg4_full", paired)
    _, arrs = T.synth_host(3000, 256, paired, 5, 9, 1, 250)
    want = T.run_cpu("oracle", p, arrs, 256)
    got = gpu.run_gpu(p, arrs, 256, mode="device")
    T.assert_results_equal(got, want, paired, what="L250")


def test_patch_list_matches_corrected_rows(gpu):
    p = T.config_params("cfg3_overlap_correction", 1)
    _, arrs = T.synth_host(8000, 160, 1, 0, 5, 1, 150)
    got = gpu.run_gpu(p, arrs, 160, mode="device")
    rebuilt = {k: v.copy() for k, v in arrs.items()}
    assert got["n_patches"] == len(got["patches"]) > 0
    for pt in got["patches"]:
        side = "2" if pt["which"] else "1"
        rebuilt["seq" + side][pt["pair"], pt["pos"]] = pt["base"]
        rebuilt["qual" + side][pt["pair"], pt["pos"]] = pt["qual"]
    for k in ("seq1", "qual1", "seq2", "qual2"):
        assert (rebuilt[k] == got["arrs"][k]).all()


@pytest.mark.parametrize("paired", [1, 0])
@pytest.mark.parametrize("L,stride,sampling", [(150, 160, 20), (100, 112, 7), (250, 256, 20)])
def test_overrepresentation_analysis(gpu, paired, L, stride, sampling):
    """config 5's scan (stats.cpp:270-288): candidate substrings counted on 1 of every `sampling` reads, pre-filter by
    read index, post-filter by rank among the counted reads; one batch, and the same stream cut into 3 launches."""
    _, arrs = T.synth_host(7000, stride, paired, 0, 5, 1, L)
    p = T.overrep_params("cfg2_cut_right_polyg" if not paired else "all_cuts", paired, arrs, L, sampling)
    p.correction_enabled = 0
    want = T.run_cpu("oracle", p, arrs, stride)
    assert sum(int(want["counters"].overrep(s)[0].sum()) for s in range(4 if paired else 2)) > 10
    got = gpu.run_gpu(p, arrs, stride, mode="device")
    T.assert_results_equal(got, want, paired, what="overrep")
    got = gpu.run_gpu(p, arrs, stride, mode="device", splits=3)
    T.assert_results_equal(got, want, paired, what="overrep-3-launches")


def test_overrepresentation_with_correction_host_mode(gpu):
    _, arrs = T.synth_host(600000, 160, 1, 0, 9, 1, 150)      # > one host chunk (262144): sampling state crosses chunks
    p = T.overrep_params("cfg3_overlap_correction", 1, arrs, 150)
    want = T.run_cpu("oracle", p, arrs, 160)
    got = gpu.run_gpu(p, arrs, 160, mode="host")
    T.assert_results_equal(got, want, 1, what="overrep-host")


@pytest.mark.parametrize("paired", [1, 0])
def test_random_option_sets(gpu, paired):
    """30 random option sets (the same generator the port is pinned to the reference with): CUDA == oracle."""
    rng = np.random.default_rng(4321 + paired)
    for k in range(30):
        p, kw = T.random_params(rng, paired)
        _, arrs = T.synth_host(2500, 160, paired, 100 * k, 700 + k, 2 if paired else 1, 150)
        want = T.run_cpu("oracle", p, arrs, 160)
        got = gpu.run_gpu(p, arrs, 160, mode="device")
        T.assert_results_equal(got, want, paired, what=f"random set {k}: {kw}")


def _dev_tensors(torch, n, paired, with_patches):
    t = {"out1": torch.zeros(max(n, 1) * 16, dtype=torch.uint8, device="cuda:0")}
    if paired:
        t["out2"] = torch.zeros(max(n, 1) * 16, dtype=torch.uint8, device="cuda:0")
        t["ov"] = torch.zeros(max(n, 1) * 8, dtype=torch.uint8, device="cuda:0")
        if with_patches:
            t["patches"] = torch.zeros((4 * n + 16) * 12, dtype=torch.uint8, device="cuda:0")
            t["np"] = torch.zeros(1, dtype=torch.int32, device="cuda:0")
    return t


def test_patches_undo_restores_the_rows(gpu):
    """fp_patches_undo plays a pass's patch list backwards: the resident rows are pristine again, a second pass gives the same
    records, counters and patch list (bench.py relies on it so that every timed step of configs[2] corrects bases)."""
    import ctypes as C
    import torch
    p = T.config_params("cfg3_overlap_correction", 1)
    n = 9000
    _, arrs = T.synth_host(n, 160, 1, 0, 17, 1, 150)
    want = T.run_cpu("oracle", p, arrs, 160)
    ctx = gpu.GpuCtx(p, n, 160, 160)
    lib = ctx.lib
    b, t = gpu.device_batch({k: v.copy() for k, v in arrs.items()})
    d = _dev_tensors(torch, n, 1, True)
    cap = 4 * n + 16
    for rep in range(2):
        ctx.reset()
        d["np"].zero_()
        capi.check(lib.fp_process_pe(ctx.h, C.byref(b), d["out1"].data_ptr(), d["out2"].data_ptr(), d["ov"].data_ptr(), d["patches"].data_ptr(), cap,
                                     d["np"].data_ptr(), None), lib)
        torch.cuda.synchronize()
        npatch = int(d["np"].item())
        assert 0 < npatch <= cap
        got = {"out1": d["out1"].cpu().numpy().view(capi.READ_RESULT_DTYPE)[:n], "out2": d["out2"].cpu().numpy().view(capi.READ_RESULT_DTYPE)[:n],
               "ov": d["ov"].cpu().numpy().view(capi.OV_RESULT_DTYPE)[:n], "counters": ctx.counters(), "arrs": {k: t[k].cpu().numpy() for k in arrs}, "layout": ctx.L}
        T.assert_results_equal(got, want, 1, what=f"pass {rep}")
        pts = d["patches"].cpu().numpy().view(capi.PATCH_DTYPE)[:npatch]
        for pt in pts[:200]:                                   # old values are the ORIGINAL bytes
            side = "2" if pt["which"] else "1"
            assert arrs["seq" + side][pt["pair"], pt["pos"]] == pt["old_base"] and arrs["qual" + side][pt["pair"], pt["pos"]] == pt["old_qual"]
        capi.check(lib.fp_patches_undo(ctx.h, C.byref(b), d["patches"].data_ptr(), d["np"].data_ptr(), cap, None), lib)
        torch.cuda.synchronize()
        for k in ("seq1", "qual1", "seq2", "qual2"):
            assert (t[k].cpu().numpy() == arrs[k]).all(), k
    ctx.close()


@pytest.mark.parametrize("nshards", [2, 3])
def test_sharded_overrepresentation_equals_single_stream(gpu, nshards):
    """shards of ONE stream processed by separate contexts (as separate ranks would), pre-filter sampling by the
    global read index (fp_batch.first_read_index), post-filter sampling by the exclusive scan of the shards' pass counts
    (fp_overrep_defer_post / fp_pass_count / fp_overrep_post); the SUM of the shards' counter blocks == the block of the whole
    stream processed as `--thread 1` would (the oracle)."""
    import ctypes as C
    import torch
    L, S, n = 250, 256, 9000
    _, arrs = T.synth_host(n, S, 1, 0, 11, 3, L)
    p = T.overrep_params("cfg3_overlap_correction", 1, arrs, L, 20)
    want = T.run_cpu("oracle", p, arrs, S)
    assert int(want["counters"].overrep(1)[0].sum()) > 5 and int(want["counters"].overrep(3)[0].sum()) > 5
    bounds = [n * i // nshards + (3 if 0 < i < nshards else 0) for i in range(nshards + 1)]      # uneven cuts, not multiples of the sampling step
    ctxs, outs, batches, counts = [], [], [], []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        ctx = gpu.GpuCtx(p, hi - lo, S, S)
        capi.check(ctx.lib.fp_overrep_defer_post(ctx.h, 1), ctx.lib)
        b, t = gpu.device_batch({k: np.ascontiguousarray(v[lo:hi]) for k, v in arrs.items()})
        b.flags, b.first_read_index = 1, lo
        d = _dev_tensors(torch, hi - lo, 1, False)
        capi.check(ctx.lib.fp_process_pe(ctx.h, C.byref(b), d["out1"].data_ptr(), d["out2"].data_ptr(), d["ov"].data_ptr(), None, 0, None, None), ctx.lib)
        c = C.c_int64()
        capi.check(ctx.lib.fp_pass_count(ctx.h, d["out1"].data_ptr(), hi - lo, C.byref(c), None), ctx.lib)
        ctxs.append(ctx); outs.append(d); batches.append((b, t)); counts.append(c.value)
    total = np.zeros(ctxs[0].L.total, np.int64)
    base = 0
    for ctx, d, (b, t), c in zip(ctxs, outs, batches, counts):
        capi.check(ctx.lib.fp_overrep_post(ctx.h, C.byref(b), d["out1"].data_ptr(), d["out2"].data_ptr(), base, None), ctx.lib)
        base += c
        total += ctx.counters().data
    assert base == int(want["counters"].stats(capi.STATS_POST1)["reads"])
    got = capi.CounterView(ctxs[0].L, total)
    # per-cycle totals (kinds 32/33) are derived per fetch, so they add up as well
    T.assert_counters_equal(got, want["counters"], what=f"{nshards} shards")
    for ctx in ctxs:
        ctx.close()


@pytest.mark.skipif(not T.have_ref(), reason="reference build (oracle/_ref) not present")
@pytest.mark.parametrize("mode", ["device", "host"])
@pytest.mark.parametrize("name,paired", [("default", 1), ("cfg4_full", 1), ("fasta_adapters", 1), ("cfg4_full", 0), ("fasta_adapters", 0), ("short_adapter", 1)])
def test_adapter_string_histograms_equal_reference(gpu, name, paired, mode):
    """the device records every FilterResult::addAdapterTrimmed call as an fp_adapter_event; replayed on the host in
    input order they give the reference's mAdapter1 / mAdapter2 maps (`adapter_cutting.read*_adapter_counts` of the JSON report)."""
    import ctypes as C
    import torch
    n, S = 20000, 160
    p = T.config_params(name, paired)
    _, arrs = T.synth_host(n, S, paired, 0, 77, 1, 150)
    want, wcnt = T.ref_adapter_maps(p, arrs, S)
    assert len(want[0]) > 3
    ctx = gpu.GpuCtx(p, n, S, S)
    lib = ctx.lib
    if mode == "device":
        cap = 8 * n
        d_ev = torch.zeros(cap * 16, dtype=torch.uint8, device="cuda:0"); d_n = torch.zeros(1, dtype=torch.int32, device="cuda:0")
        capi.check(lib.fp_set_event_sink(ctx.h, d_ev.data_ptr(), cap, d_n.data_ptr()), lib)
        got = gpu.run_gpu(p, arrs, S, mode="device", ctx=ctx)
        ne = int(d_n.item())
        ev = d_ev.cpu().numpy().view(capi.EVENT_DTYPE)[:ne].copy()
    else:
        cap = 8 * n
        h_ev = np.zeros(cap, capi.EVENT_DTYPE); h_n = C.c_uint64()
        capi.check(lib.fp_set_host_event_sink(ctx.h, h_ev.ctypes.data, cap, C.byref(h_n)), lib)
        got = gpu.run_gpu(p, arrs, S, mode="host", ctx=ctx)
        ne = h_n.value
        ev = h_ev[:ne].copy()
    assert 0 < ne <= cap
    adapters = [(p.adapter_seq_r1 or b"").decode(), (p.adapter_seq_r2 or b"").decode()] + [p.fasta_adapters[i].decode() for i in range(p.n_fasta_adapters)]
    maps = T.rebuild_adapter_maps(ev, got["arrs"], adapters)
    assert maps[0] == want[0]
    assert maps[1] == want[1]
    T.assert_counters_equal(got["counters"], wcnt, what="counters")
    ctx.close()


@pytest.mark.parametrize("name,paired,L,S", [("cfg3_overlap_correction", 1, 150, 160), ("cfg4_full", 1, 150, 160), ("cfg2_cut_right_polyg", 0, 150, 160),
                                             ("cfg4_full", 1, 250, 256), ("default", 1, 100, 112)])
def test_packed_host_rows(gpu, name, paired, L, S):
    """the packed end-to-end entry points (2-bit bases, N exception list, unpadded qualities; H2D of ~60 % of the bytes, rows restored
    on the device) give the same records, counters and -- through the patch list -- corrected rows as the oracle"""
    p = T.config_params(name, paired)
    _, arrs = T.synth_host(300000 if L == 150 and paired else 20000, S, paired, 5, 44, 1, L)      # the large one crosses a host chunk
    want = T.run_cpu("oracle", p, arrs, S)
    got = gpu.run_gpu(p, arrs, S, mode="packed")
    T.assert_results_equal(got, want, paired, what=f"packed {name}")


@pytest.mark.parametrize("name,paired", [("cfg3_overlap_correction", 1), ("cfg2_cut_right_polyg", 0)])
def test_host_rows_at_a_tight_pitch(gpu, name, paired):
    """fp_process_*_host with fp_batch.stride = the longest read (150) instead of the device stride (160): same results, corrected rows
    patched at the caller's pitch"""
    p = T.config_params(name, paired)
    _, arrs = T.synth_host(300000, 160, paired, 5, 46, 1, 150)
    want = T.run_cpu("oracle", p, arrs, 160)
    got = gpu.run_gpu(p, arrs, 160, mode="host_tight")
    T.assert_results_equal(got, want, paired, what=f"tight pitch {name}")


@pytest.mark.parametrize("name,paired,mode", [("cfg3_overlap_correction", 1, "host_tight_pack2bit"), ("cfg4_full", 1, "host_pack2bit"),
                                              ("cfg2_cut_right_polyg", 0, "host_tight_pack2bit")])
def test_host_rows_packed_on_the_fly(gpu, name, paired, mode):
    """FP_B_PACK2BIT: the library's host threads pack the bases to 2 bits chunk by chunk while the previous chunks are copied (1.2 M units =
    five chunks, so the staging slots of the packing team are re-used); results, counters and the corrected rows in the caller's buffers
    are the same as without the flag"""
    p = T.config_params(name, paired)
    _, arrs = T.synth_host(1200000, 160, paired, 5, 47, 1, 150)
    want = T.run_cpu("oracle", p, arrs, 160)
    got = gpu.run_gpu(p, arrs, 160, mode=mode)
    T.assert_results_equal(got, want, paired, what=f"{mode} {name}")


def test_pack2bit_rejects_other_bytes(gpu):
    """a base outside {A,C,G,T,N} cannot be packed: FP_E_UNSUPPORTED, not a wrong answer"""
    p = T.config_params("default", 1)
    _, arrs = T.synth_host(5000, 160, 1, 5, 48, 1, 150)
    arrs["seq1"][1234, 7] = ord("R")
    with pytest.raises(Exception):
        gpu.run_gpu(p, arrs, 160, mode="host_pack2bit")


@pytest.mark.parametrize("L,S", [(150, 160), (100, 112)])
@pytest.mark.parametrize("name", T.MERGE_CONFIG_NAMES)
def test_merge_mode(gpu, name, L, S):
    """--merge / --include_unmerged on the device (src/peprocessor.cpp:519-560, OverlapAnalysis::merge src/overlapanalysis.cpp:148-179):
    records (FP_F_MERGED, the merged read's verdict), the second overlap analysis, mMergedPairs and the post-filter Stats over merged reads
    up to two rows long (counter block sized 2 x stride) equal the oracle's"""
    p = T.config_params(name, 1)
    _, arrs = T.synth_host(20000, S, 1, 700, 78, 1, L)
    want = T.run_cpu("oracle", p, arrs, 2 * S)
    got = gpu.run_gpu(p, arrs, 2 * S)
    T.assert_results_equal(got, want, 1, what=name)
    assert int(got["counters"].filter[107]) > 0


def test_merge_mode_with_other_bytes(gpu):
    """merge mode on rows holding bytes outside {A,C,G,T,N} (byte paths of the analysis, complement of an unknown base = N)"""
    p = T.config_params("merge_cfg4_full", 1)
    _, arrs = T.synth_host(6000, 160, 1, 0, 79, 1, 150)
    rng = np.random.default_rng(5)
    for k in ("seq1", "seq2"):
        rows = rng.integers(0, 6000, 600); cols = rng.integers(0, 150, 600)
        arrs[k][rows, cols] = rng.choice(np.frombuffer(b"acgtnRYK.", np.uint8), 600)
    want = T.run_cpu("oracle", p, arrs, 320)
    got = gpu.run_gpu(p, arrs, 320)
    T.assert_results_equal(got, want, 1, what="merge other bytes")


def test_two_contexts_with_different_params_on_one_device(gpu):
    """The operator parameters sit in one __constant__ block per device: contexts with DIFFERENT fp_params that take turns on the device
    (each on its own stream, no host synchronisation in between) must each see their own parameters in every launch."""
    import ctypes as C
    torch = gpu._torch()
    pa, pb = T.config_params("cfg3_overlap_correction", 1), T.config_params("all_cuts", 1)
    n, S = 60000, 160
    _, arrs = T.synth_host(n, S, 1, 5, 91, 1, 150)
    want = {k: T.run_cpu("oracle", p, arrs, S) for k, p in (("a", pa), ("b", pb))}
    ctx = {"a": gpu.GpuCtx(pa, n, S, S), "b": gpu.GpuCtx(pb, n, S, S)}
    st = {k: torch.cuda.Stream() for k in ctx}
    dev, outs = {}, {}
    for k in ctx:
        _, t = gpu.device_batch({kk: v.copy() for kk, v in arrs.items()})
        dev[k] = t
        outs[k] = [torch.zeros(n * 16, dtype=torch.uint8, device="cuda:0") for _ in range(2)] + [torch.zeros(n * 8, dtype=torch.uint8, device="cuda:0")]
        outs[k] += [torch.zeros((4 * n + 16) * 12, dtype=torch.uint8, device="cuda:0"), torch.zeros(1, dtype=torch.int32, device="cuda:0")]
    torch.cuda.synchronize()
    parts = 12
    bounds = [n * i // parts for i in range(parts + 1)]
    lib = ctx["a"].lib
    for lo, hi in zip(bounds[:-1], bounds[1:]):              # a, b, a, b ... slices of the same rows, no sync between the launches
        for k in ("a", "b"):
            sb = capi.Batch(); sb.n, sb.stride = hi - lo, S
            for kk, v in dev[k].items():
                setattr(sb, kk, v.data_ptr() + lo * (2 if kk.startswith("len") else S))
            o = outs[k]
            # patch indices are batch-local: give every slice its own region of the patch buffer (only the counters / records are compared)
            capi.check(lib.fp_process_pe(ctx[k].h, C.byref(sb), o[0].data_ptr() + lo * 16, o[1].data_ptr() + lo * 16, o[2].data_ptr() + lo * 8,
                                         None, 0, None, C.c_void_p(st[k].cuda_stream)), lib)
    torch.cuda.synchronize()
    for k in ctx:
        got = {"out1": outs[k][0].cpu().numpy().view(capi.READ_RESULT_DTYPE)[:n], "out2": outs[k][1].cpu().numpy().view(capi.READ_RESULT_DTYPE)[:n],
               "ov": outs[k][2].cpu().numpy().view(capi.OV_RESULT_DTYPE)[:n], "counters": ctx[k].counters(),
               "arrs": {kk: v.cpu().numpy() for kk, v in dev[k].items()}, "layout": ctx[k].L}
        T.assert_results_equal(got, want[k], 1, what=f"context {k}")
        ctx[k].close()


def test_failed_ctx_create_leaves_nothing_behind(gpu):
    """a context that fails half-way through its construction is torn down (fp_ctx_create unwinds through fp_ctx_destroy): many failures
    in a row neither exhaust the device nor disturb a following context"""
    import ctypes as C
    torch = gpu._torch()
    lib = capi.load()
    free0 = torch.cuda.mem_get_info()[0]
    bad = T.config_params("default", 1)
    capi.set_params(bad, adapter_seq_r1="A" * 300)                      # longer than FP_MAX_ADAPTER_LEN: rejected inside the construction
    for _ in range(50):
        h = C.c_void_p()
        rc = lib.fp_ctx_create(C.byref(bad), 0, 1 << 20, 160, 160, C.byref(h))
        assert rc != 0 and not h.value
        assert b"adapter" in lib.fp_last_error()
    assert free0 - torch.cuda.mem_get_info()[0] < 64 << 20
    p = T.config_params("cfg3_overlap_correction", 1)
    _, arrs = T.synth_host(3000, 160, 1, 5, 3, 1, 150)
    T.assert_results_equal(gpu.run_gpu(p, arrs, 160), T.run_cpu("oracle", p, arrs, 160), 1, what="after failed creates")
