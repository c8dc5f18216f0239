"""Worker for tests/test_multirank_gloo.py: one rank of a world_size-N gloo job.  Each rank runs the hot path
(CPU oracle standing in for the device on this GPU-less box) over ITS shard and the counter blocks are
all-reduced exactly as bench.py does with NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fp_testlib as T  # noqa: E402
from fastp_b200 import sharding  # noqa: E402


def main():
    total, out_path, mode = int(sys.argv[1]), sys.argv[2], sys.argv[3]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    p = T.config_params("cfg4_full", 1)
    if mode == "strong":
        lo, hi = sharding.shard_range(rank, world, total)
    else:
        lo = sharding.weak_first_index(rank, total // world); hi = lo + total // world
    _, arrs = T.synth_host(hi - lo, 160, 1, lo, 42, 1, 150)
    res = T.run_cpu("oracle", p, arrs, 160)
    t = torch.from_numpy(res["counters"].data.copy())
    sharding.allreduce_counters(t)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
