#!/usr/bin/env python
"""One-off wider fuzz on the GPU box: random option sets x read lengths, CUDA (device and host mode) vs the C port."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fp_testlib as T  # noqa: E402
import fp_gpu  # noqa: E402

nfail = 0
for paired in (1, 0):
    rng = np.random.default_rng(99 + paired)
    for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
        L, stride = [(100, 112), (150, 160), (250, 256)][k % 3]
        p, kw = T.random_params(rng, paired)
        _, arrs = T.synth_host(1200, stride, paired, 37 * k, 5000 + k, 2 if paired else 1, L)
        want = T.run_cpu("oracle", p, arrs, stride)
        for mode in (("device",) if k % 4 else ("device", "host")):
            try:
                got = fp_gpu.run_gpu(p, arrs, stride, mode=mode)
                T.assert_results_equal(got, want, paired, what=f"{'PE' if paired else 'SE'} set {k} L{L} {mode}")
            except AssertionError as e:
                nfail += 1
                print("FAIL", str(e)[:300], kw)
print("fuzz done, failures:", nfail)
