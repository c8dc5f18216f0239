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
