"""-m gpu: the duplication bloom filter on the device (fp_dup_check) against the sequential C port, which is pinned
to the reference's Duplicate object (tests/test_duplicate_oracle.py).  First ran on a B200 at the end of round 1 (equal); a plain test
since.  It still runs in a child process (tests/_gpu_dup_worker.py) so that its multi-GiB bit arrays are released at once."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("paired", [1, 0])
def test_device_duplicate_filter_equals_oracle(paired):
    r = subprocess.run([sys.executable, os.path.join(HERE, "_gpu_dup_worker.py"), str(paired)], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-1500:])
