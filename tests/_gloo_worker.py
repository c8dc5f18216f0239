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
    if mode == "overrep":
        return overrep(total, out_path, rank, world)
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


def overrep(total, out_path, rank, world):
    """SURVEY 8(e)'s second exchange on CPU: pre-filter over-representation sampling by the global read index, post-filter sampling
    by the exclusive scan of the shards' pass counts (sharding.exclusive_pass_base over gloo).  The oracle samples on the running
    `reads` counters of the block it adds into (as Stats::statRead does on mReads), so a shard is run with those counters seeded
    with its position in the stream and the seeds are taken out again before the all-reduce."""
    import ctypes as C
    from fastp_b200 import capi
    L_, S = 150, 160
    _, arrs = T.synth_host(total, S, 1, 0, 42, 3, L_)
    p = T.overrep_params("cfg3_overlap_correction", 1, arrs, L_, 20)
    lo, hi = sharding.shard_range(rank, world, total)
    lay = capi.make_layout(T.oracle(), 1, S, p.insert_size_max, params=p)

    def run(pass_base):
        a = {k: np.ascontiguousarray(v[lo:hi]).copy() for k, v in arrs.items()}
        b = capi.batch_from_arrays(a)
        c = np.zeros(lay.total, np.int64)
        seeds = {capi.STATS_PRE1: lo, capi.STATS_PRE2: lo, capi.STATS_POST1: pass_base, capi.STATS_POST2: pass_base}
        for s, v in seeds.items():
            c[s * lay.stats_stride + lay.off_reads] = v
        o1 = np.zeros(hi - lo, capi.READ_RESULT_DTYPE); o2 = np.zeros(hi - lo, capi.READ_RESULT_DTYPE)
        assert T.oracle().fp_oracle_process(C.byref(p), C.byref(lay), C.byref(b), o1.ctypes.data, o2.ctypes.data, None, c.ctypes.data) == 0
        for s, v in seeds.items():
            c[s * lay.stats_stride + lay.off_reads] -= v
        return c, int((o1["pair_verdict"] == 0).sum())
    _, npass = run(0)
    base, _tot = sharding.exclusive_pass_base(npass)
    c, _ = run(base)
    t = torch.from_numpy(c)
    sharding.allreduce_counters(t)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
