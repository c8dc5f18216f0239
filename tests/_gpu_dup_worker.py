"""Child process of tests/test_gpu_duplicate.py: device duplicate filter vs the sequential C port.  Runs in its own process so that a
fault in kernels that have not seen hardware yet cannot poison the CUDA context of the rest of the -m gpu suite.  Exit code 0 = equal."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import fp_testlib as T  # noqa: E402
from fastp_b200 import capi  # noqa: E402
from test_duplicate_oracle import planted  # noqa: E402


def main(paired):
    import torch
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
    for lo, hi in ((0, n // 3), (n // 3, n // 2), (n // 2, n)):
        sub = {k: np.ascontiguousarray(v[lo:hi]) for k, v in arrs.items()}
        hb = capi.batch_from_arrays(sub)
        want = np.zeros(hi - lo, np.uint8)
        olib.fp_oracle_dup_check(od, C.byref(hb), paired, want.ctypes.data)
        db, _t = fp_gpu.device_batch(sub)
        d_flags = torch.zeros(hi - lo, dtype=torch.uint8, device="cuda:0")
        capi.check(ctx.lib.fp_dup_check(ctx.h, C.byref(db), 1, d_flags.data_ptr(), None), ctx.lib)
        torch.cuda.synchronize()
        got = d_flags.cpu().numpy()
        if not np.array_equal(got, want):
            print("flags differ in batch", lo, hi, int((got != want).sum()), "of", hi - lo); return 1
    to, do, tg, dg = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    olib.fp_oracle_dup_totals(od, C.byref(to), C.byref(do))
    capi.check(ctx.lib.fp_dup_totals(ctx.h, C.byref(tg), C.byref(dg)), ctx.lib)
    if (tg.value, dg.value) != (to.value, do.value):
        print("totals differ", tg.value, dg.value, to.value, do.value); return 1
    capi.check(ctx.lib.fp_dup_reset(ctx.h), ctx.lib)
    capi.check(ctx.lib.fp_dup_totals(ctx.h, C.byref(tg), C.byref(dg)), ctx.lib)
    if (tg.value, dg.value) != (0, 0):
        print("reset failed"); return 1
    print("device duplicate filter == oracle:", to.value, "units,", do.value, "duplicates")
    return 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1])))
