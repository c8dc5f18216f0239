"""N>1 path on CPU: world_size-2 and -3 gloo jobs shard the units by contiguous index range and all-reduce the
counter block; the result must equal the single-process block over the union (sums commute)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import fp_testlib as T
from fastp_b200 import sharding

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_range_partitions():
    for total in (0, 1, 7, 1000, 12345):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(r, world, total) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("world,mode", [(2, "weak"), (2, "strong"), (3, "strong")])
def test_sharded_allreduce_equals_single_process(tmp_path, world, mode):
    total = 3001 if mode == "strong" else 3000
    out = str(tmp_path / "c.npy")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world * 7 + (1 if mode == "weak" else 0)), os.path.join(HERE, "_gloo_worker.py"), str(total), out, mode]
    subprocess.run(cmd, check=True, env=env, timeout=300, capture_output=True)
    got = np.load(out)
    p = T.config_params("cfg4_full", 1)
    n = total if mode == "strong" else (total // world) * world
    _, arrs = T.synth_host(n, 160, 1, 0, 42, 1, 150)
    want = T.run_cpu("oracle", p, arrs, 160)["counters"].data
    assert (got == want).all()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_overrepresentation_sampling_equals_single_process(tmp_path, world):
    """the pass-count exclusive scan (sharding.exclusive_pass_base) over gloo: all-reduced over-representation counts of N shards ==
    the single-process `--thread 1` block (src/stats.cpp:270-290)"""
    from fastp_b200 import capi
    total = 4003
    out = str(tmp_path / "o.npy")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29560 + world), os.path.join(HERE, "_gloo_worker.py"), str(total), out, "overrep"]
    subprocess.run(cmd, check=True, env=env, timeout=600, capture_output=True)
    got = np.load(out)
    _, arrs = T.synth_host(total, 160, 1, 0, 42, 3, 150)
    p = T.overrep_params("cfg3_overlap_correction", 1, arrs, 150, 20)
    want = T.run_cpu("oracle", p, arrs, 160)["counters"]
    assert int(want.overrep(capi.STATS_POST1)[0].sum()) > 3
    # the finalised per-cycle totals (kinds 32/33) are part of the oracle's block; the raw sums carry them too
    assert (got == want.data).all()
