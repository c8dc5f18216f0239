"""-m gpu, needs >= 2 GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`): N ranks over NCCL, each the CUDA hot path on its shard,
counter block all-reduced through the C-ABI collective on a raw ncclComm_t, over-representation sampling through the pass-count scan;
the result equals the single-process `--thread 1` block of the whole stream."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4])
def test_allreduced_block_equals_single_process(tmp_path, world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    out = os.path.join(tmp_path, "blk.npy")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(HERE, "_nccl_worker.py"), "12001", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert os.path.exists(out)
