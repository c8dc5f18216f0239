"""-m gpu: the duplication bloom filter on the device (fp_dup_check) against the sequential C port, which is pinned
to the reference's Duplicate object (tests/test_duplicate_oracle.py).

Status: the per-thread bodies (fastp_b200/csrc/fp_dup.h) are validated on the host -- run one thread at a time in shuffled order they
reproduce the oracle -- but round 1 ran out of GPU minutes before the CUDA wrappers saw hardware, so this first on-device comparison is
marked xfail(strict=False): a pass shows up as XPASS, a difference as xfail, neither hides behind a green tick."""
import ctypes as C

import numpy as np
import pytest

import fp_testlib as T
from fastp_b200 import capi
from test_duplicate_oracle import planted

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="first hardware run of the fp_dup kernels (host-emulated so far)")]


@pytest.mark.parametrize("paired", [1, 0])
def test_device_duplicate_filter_equals_oracle(paired):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required")
    import fp_gpu
    olib = T.oracle()
    olib.fp_oracle_dup_create.restype = C.c_void_p; olib.fp_oracle_dup_create.argtypes = [C.c_int]
    olib.fp_oracle_dup_check.argtypes = [C.c_void_p, C.POINTER(capi.Batch), C.c_int, C.c_void_p]
    olib.fp_oracle_dup_totals.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    olib.fp_oracle_dup_destroy.argtypes = [C.c_void_p]
    arrs = planted(paired, n=20000, seed=55 + paired)
    n = len(arrs["len1"])
    ctx = fp_gpu.GpuCtx(T.config_params("default", paired), n, 160, 160)
    od = olib.fp_oracle_dup_create(1)
    try:
        for lo, hi in ((0, n // 3), (n // 3, n // 2), (n // 2, n)):
            sub = {k: np.ascontiguousarray(v[lo:hi]) for k, v in arrs.items()}
            hb = capi.batch_from_arrays(sub)
            want = np.zeros(hi - lo, np.uint8)
            olib.fp_oracle_dup_check(od, C.byref(hb), paired, want.ctypes.data)
            db, _t = fp_gpu.device_batch(sub)
            d_flags = torch.zeros(hi - lo, dtype=torch.uint8, device="cuda:0")
            capi.check(ctx.lib.fp_dup_check(ctx.h, C.byref(db), 1, d_flags.data_ptr(), None), ctx.lib)
            torch.cuda.synchronize()
            assert np.array_equal(d_flags.cpu().numpy(), want), (paired, lo)
        to, do, tg, dg = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        olib.fp_oracle_dup_totals(od, C.byref(to), C.byref(do))
        capi.check(ctx.lib.fp_dup_totals(ctx.h, C.byref(tg), C.byref(dg)), ctx.lib)
        assert (tg.value, dg.value) == (to.value, do.value)
        capi.check(ctx.lib.fp_dup_reset(ctx.h), ctx.lib)
        capi.check(ctx.lib.fp_dup_totals(ctx.h, C.byref(tg), C.byref(dg)), ctx.lib)
        assert (tg.value, dg.value) == (0, 0)
    finally:
        olib.fp_oracle_dup_destroy(od)
        ctx.close()
