"""Duplication bloom filter (src/duplicate.cpp): the C port pinned against the reference's own Duplicate object.
Oracle first -- the device path is next round's work ("What comes next in 8f")."""
import ctypes as C

import numpy as np
import pytest

import fp_testlib as T
from fastp_b200 import capi

pytestmark = pytest.mark.reference
needs_ref = pytest.mark.skipif(not T.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def planted(paired, n=4000, seed=9):
    """n units, then 40 % more drawn from them again (exact duplicates) plus a few one-base variants, shuffled."""
    rng = np.random.default_rng(seed)
    _, arrs = T.synth_host(n, 160, paired, 0, seed, 1, 150)
    pick = rng.integers(0, n, int(n * 0.4))
    out = {k: np.concatenate([v, v[pick]]) for k, v in arrs.items()}
    for i in rng.integers(n, len(out["len1"]), 60):              # near-duplicates: one substituted base
        if out["len1"][i] > 10:
            out["seq1"][i, 5] = ord("A") if out["seq1"][i, 5] != ord("A") else ord("C")
    perm = rng.permutation(len(out["len1"]))
    return {k: np.ascontiguousarray(v[perm]) for k, v in out.items()}


@needs_ref
@pytest.mark.parametrize("paired", [1, 0])
def test_port_equals_reference_duplicate(paired):
    olib, rlib = T.oracle(), T.ref()
    olib.fp_oracle_dup_create.restype = C.c_void_p; olib.fp_oracle_dup_create.argtypes = [C.c_int]
    olib.fp_oracle_dup_check.argtypes = [C.c_void_p, C.POINTER(capi.Batch), C.c_int, C.c_void_p]
    olib.fp_oracle_dup_totals.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    olib.fp_oracle_dup_destroy.argtypes = [C.c_void_p]
    rlib.fp_ref_dup_create.restype = C.c_void_p; rlib.fp_ref_dup_create.argtypes = [C.c_int]
    rlib.fp_ref_dup_check.argtypes = [C.c_void_p, C.POINTER(capi.Batch), C.c_int, C.c_void_p]
    rlib.fp_ref_dup_totals.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    rlib.fp_ref_dup_destroy.argtypes = [C.c_void_p]
    arrs = planted(paired)
    n = len(arrs["len1"])
    od, rd = olib.fp_oracle_dup_create(1), rlib.fp_ref_dup_create(1)
    assert od and rd
    try:
        flags_o, flags_r = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        for lo, hi in ((0, n // 3), (n // 3, n // 2), (n // 2, n)):          # state carries over batches
            sub = {k: np.ascontiguousarray(v[lo:hi]) for k, v in arrs.items()}
            b = capi.batch_from_arrays(sub)
            fo, fr = np.zeros(hi - lo, np.uint8), np.zeros(hi - lo, np.uint8)
            olib.fp_oracle_dup_check(od, C.byref(b), paired, fo.ctypes.data)
            rlib.fp_ref_dup_check(rd, C.byref(b), paired, fr.ctypes.data)
            flags_o[lo:hi], flags_r[lo:hi] = fo, fr
        assert np.array_equal(flags_o, flags_r)
        to, do, tr, dr, rate = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_double()
        olib.fp_oracle_dup_totals(od, C.byref(to), C.byref(do))
        rlib.fp_ref_dup_totals(rd, C.byref(tr), C.byref(dr), C.byref(rate))
        assert (to.value, do.value) == (tr.value, dr.value) == (n, int(flags_r.sum()))
        assert abs(rate.value - do.value / n) < 1e-12
        # every second copy of a planted exact duplicate is flagged (the filter has no false negatives)
        assert do.value >= int(n / 1.4 * 0.4) - 80
    finally:
        olib.fp_oracle_dup_destroy(od); rlib.fp_ref_dup_destroy(rd)
