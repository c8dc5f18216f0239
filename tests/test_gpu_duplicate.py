"""-m gpu: the duplication bloom filter on the device (fp_dup_check) against the sequential C port, which is pinned
to the reference's Duplicate object (tests/test_duplicate_oracle.py).

Status: the per-thread bodies (fastp_b200/csrc/fp_dup.h) are validated on the host -- run one thread at a time in shuffled order they
reproduce the oracle -- but round 1 ran out of GPU minutes before the CUDA wrappers saw hardware.  So this first on-device comparison
(a) runs in a CHILD PROCESS (tests/_gpu_dup_worker.py): a fault cannot poison the CUDA context of the rest of the suite, and
(b) is marked xfail(strict=False): a pass shows up as XPASS, a difference as xfail, neither hides behind a green tick."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="first hardware run of the fp_dup kernels (host-emulated so far)")]
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("paired", [1, 0])
def test_device_duplicate_filter_equals_oracle(paired):
    r = subprocess.run([sys.executable, os.path.join(HERE, "_gpu_dup_worker.py"), str(paired)], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-1500:])
