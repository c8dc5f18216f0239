"""Multi-GPU plumbing of the hot path (one process per GPU).

Reads/pairs are independent units: rank r of W owns the contiguous global index range
[r*units_per_rank, (r+1)*units_per_rank) (weak scaling) or an even split of a fixed total (strong).
There is no data-path collective.  Two tiny exchanges exist:

  * ONE sum all-reduce of the packed int64 counter block at the end of a pass -- what Stats::merge
    (src/stats.cpp:877-955) and FilterResult::merge (src/filterresult.cpp:38-89) do across the reference's worker
    threads.  `NcclComm` bootstraps a raw ncclComm_t (unique id from rank 0, carried by whatever process group is
    up) so that the C-ABI collective `fp_counters_allreduce(ctx, comm, stream)` can be used as a C++ caller would.
  * with over-representation analysis, an exclusive scan of the per-shard counts of PASSING units, because the
    post-filter Stats sample every `sampling`-th read that passed (src/stats.cpp:272,290): `exclusive_pass_base`.

torch is plumbing only (rendezvous + the small all_gather).
"""
import ctypes as C
import glob
import os


def shard_range(rank, world, total):
    """Even contiguous split of [0,total): first (total % world) ranks get one extra unit."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def weak_first_index(rank, units_per_rank):
    return rank * units_per_rank


def allreduce_counters(t):
    """In-place SUM all-reduce of an int64 counter tensor (CPU/gloo or CUDA/NCCL); no-op when not distributed."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def exclusive_pass_base(count, device=None):
    """Number of passing units on the ranks before this one (exclusive scan of one int64 per rank) and the job's total.
    `count` is this rank's fp_pass_count; works on gloo (CPU tensor) and NCCL (pass the rank's cuda device)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0, int(count)
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([int(count)], dtype=torch.int64, device=device or "cpu")
    allc = torch.zeros(world, dtype=torch.int64, device=device or "cpu")
    dist.all_gather_into_tensor(allc, mine)
    allc = allc.cpu().tolist()
    return int(sum(allc[:rank])), int(sum(allc))


class _NcclUniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]          # nccl.h: NCCL_UNIQUE_ID_BYTES


def _find_nccl():
    """The NCCL already in (or loadable into) this process: torch bundles libnccl.so.2."""
    for name in ("libnccl.so.2", "libnccl.so"):
        try:
            return C.CDLL(name, mode=C.RTLD_GLOBAL)
        except OSError:
            pass
    try:
        import torch
        cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "nccl", "lib", "libnccl.so*"))
    except Exception:
        cands = []
    for c in sorted(cands):
        try:
            return C.CDLL(c, mode=C.RTLD_GLOBAL)
        except OSError:
            pass
    raise RuntimeError("libnccl not found")


class NcclComm:
    """A raw ncclComm_t for the C-ABI collective.  Call with the CUDA device of this rank current."""

    def __init__(self, world, rank):
        import torch
        import torch.distributed as dist
        self.lib = _find_nccl()
        self.lib.ncclGetUniqueId.argtypes = [C.POINTER(_NcclUniqueId)]
        self.lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _NcclUniqueId, C.c_int]
        self.lib.ncclCommDestroy.argtypes = [C.c_void_p]
        uid = _NcclUniqueId()
        if rank == 0:
            rc = self.lib.ncclGetUniqueId(C.byref(uid))
            if rc != 0:
                raise RuntimeError(f"ncclGetUniqueId failed ({rc})")
        # the 128 id bytes travel over the process group that is already up
        buf = torch.frombuffer(bytearray(C.string_at(C.byref(uid), 128) if rank == 0 else bytes(128)), dtype=torch.uint8).clone()
        backend = dist.get_backend()
        if backend == "nccl":
            buf = buf.cuda()
        dist.broadcast(buf, src=0)
        raw = bytes(buf.cpu().numpy().tobytes())
        C.memmove(C.byref(uid), raw, 128)
        self.comm = C.c_void_p()
        rc = self.lib.ncclCommInitRank(C.byref(self.comm), world, uid, rank)
        if rc != 0:
            raise RuntimeError(f"ncclCommInitRank failed ({rc})")

    @property
    def handle(self):
        return self.comm

    def destroy(self):
        if self.comm:
            self.lib.ncclCommDestroy(self.comm)
            self.comm = C.c_void_p()
