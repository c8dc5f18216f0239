"""CPU-side checks of the drop-in boundary: libfastp_b200.so loads, exports every symbol that
include/fastp_b200.h declares, the ctypes mirrors have the C sizes, and the product does not depend on oracle/.
No compute call is made (no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from fastp_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return capi.load()


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "fastp_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    names = set(re.findall(r"^\s*(?:int|void|void\*|size_t|int64_t|const char\*)\s+\*?\s*(fp_\w+)\s*\(", h, flags=re.M))
    return names


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    assert names == set(capi.SYMBOLS), names ^ set(capi.SYMBOLS)
    for n in names:
        assert getattr(lib, n) is not None


def test_struct_sizes_match(lib):
    assert lib.fp_abi_sizeof(0) == C.sizeof(capi.Params)
    assert lib.fp_abi_sizeof(1) == C.sizeof(capi.Batch)
    assert lib.fp_abi_sizeof(2) == capi.READ_RESULT_DTYPE.itemsize == 16
    assert lib.fp_abi_sizeof(3) == capi.OV_RESULT_DTYPE.itemsize == 8
    assert lib.fp_abi_sizeof(4) == capi.PATCH_DTYPE.itemsize == 12
    assert lib.fp_abi_sizeof(5) == C.sizeof(capi.CounterLayout)
    assert lib.fp_abi_sizeof(8) == capi.EVENT_DTYPE.itemsize == 16
    assert lib.fp_abi_sizeof(9) == C.sizeof(capi.PackedBatch)
    assert lib.fp_abi_sizeof(10) == 8


def test_defaults_mirror_reference_options(lib):
    p = capi.default_params(1, lib=lib)
    # Options::Options() options.cpp:9-32 and nested ctors options.h
    assert (p.overlap_require, p.overlap_diff_limit, p.overlap_diff_percent_limit, p.insert_size_max) == (30, 5, 20, 512)
    assert (p.qual_filter_enabled, p.qualified_qual, p.unqualified_percent_limit, p.n_base_limit) == (1, ord("0"), 40, 5)
    assert (p.length_filter_enabled, p.length_required, p.adapter_enabled, p.dimer_max_len) == (1, 15, 1, 2)
    assert (p.cut_front_window, p.cut_right_quality, p.polyg_min_len, p.polyx_min_len) == (4, 20, 10, 10)
    L = capi.make_layout(lib, 1, 152, 512)
    assert L.total == 4 * (34 * 152 + 1024 + 128 + 2) + 108 + 513


def test_no_gpu_means_loud_failure_not_fallback(lib):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    p = capi.default_params(1, lib=lib)
    h = C.c_void_p()
    rc = lib.fp_ctx_create(C.byref(p), 0, 1024, 160, 160, C.byref(h))
    assert rc == -2 and b"no CPU fallback" in lib.fp_last_error()


def test_product_does_not_link_the_oracle(lib):
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "fastp_ref" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "fp_oracle" not in syms and "fp_ref_" not in syms
    for root, _, files in os.walk(os.path.join(ROOT, "fastp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cpp")):
                txt = open(os.path.join(root, f)).read()
                assert "fp_oracle_process" not in txt and "libfastp_oracle" not in txt and "libfastp_ref" not in txt, f
