"""Worker for tests/test_gpu_multirank.py: one rank (= one GPU) of a torchrun job.  Every rank runs the CUDA hot path over ITS shard of
one stream, the two cross-rank exchanges of SURVEY 8(e) go through the C-ABI (fp_counters_allreduce on a raw ncclComm_t) and the
pass-count scan, and rank 0 compares the all-reduced block with the CPU oracle's block of the WHOLE stream (= `--thread 1`)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fp_testlib as T  # noqa: E402
import fp_gpu  # noqa: E402
from fastp_b200 import capi, sharding  # noqa: E402


def main():
    total, out_path = int(sys.argv[1]), sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    comm = sharding.NcclComm(world, rank)
    L, S = 250, 256
    _, arrs = T.synth_host(total, S, 1, 0, 23, 3, L)                 # every rank regenerates the stream (candidates need its start)
    p = T.overrep_params("cfg3_overlap_correction", 1, arrs, L, 20)
    lo, hi = sharding.shard_range(rank, world, total)
    ctx = fp_gpu.GpuCtx(p, hi - lo, S, S, device=local)
    lib = ctx.lib
    capi.check(lib.fp_overrep_defer_post(ctx.h, 1), lib)
    t = {k: torch.from_numpy(np.ascontiguousarray(v[lo:hi])).to(f"cuda:{local}") for k, v in arrs.items()}
    b = capi.Batch(); b.n, b.stride, b.flags, b.first_read_index = hi - lo, S, 1, lo
    for k, v in t.items():
        setattr(b, k, v.data_ptr())
    o1 = torch.zeros((hi - lo) * 16, dtype=torch.uint8, device=f"cuda:{local}"); o2 = torch.zeros_like(o1)
    st = torch.cuda.Stream()
    sp = C.c_void_p(st.cuda_stream)
    capi.check(lib.fp_process_pe(ctx.h, C.byref(b), o1.data_ptr(), o2.data_ptr(), None, None, 0, None, sp), lib)
    cnt = C.c_int64()
    capi.check(lib.fp_pass_count(ctx.h, o1.data_ptr(), hi - lo, C.byref(cnt), sp), lib)
    base, tot = sharding.exclusive_pass_base(cnt.value, device=f"cuda:{local}")
    capi.check(lib.fp_overrep_post(ctx.h, C.byref(b), o1.data_ptr(), o2.data_ptr(), base, sp), lib)
    capi.check(lib.fp_counters_allreduce(ctx.h, comm.handle, sp), lib)      # the C-ABI collective on a raw ncclComm_t
    got = ctx.counters()
    ok = True
    if rank == 0:
        want = T.run_cpu("oracle", p, arrs, S)
        try:
            T.assert_counters_equal(got, want["counters"], what=f"{world} ranks all-reduced")
            assert tot == int(want["counters"].stats(capi.STATS_POST1)["reads"])
            assert int(want["counters"].overrep(1)[0].sum()) > 5
            np.save(out_path, got.data)
        except AssertionError as e:
            print("MISMATCH", e); ok = False
    dist.barrier()
    ctx.close()
    comm.destroy()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
