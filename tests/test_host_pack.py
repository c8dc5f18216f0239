"""fp_host_pack_rows (host side of the packed end-to-end path): 2-bit bases + 'N' exception list + qualities at the caller's pitch.
Unpacked again in numpy it must give back the rows; bases outside {A,C,G,T,N} are refused."""
import ctypes as C

import numpy as np
import pytest

import fp_testlib as T
from fastp_b200 import capi
import fp_gpu


def unpack(pb, keep, n, S, side):
    L = keep["len" + side][:n]
    bases = keep["bases" + side][: n * pb.pitch_b].reshape(n, pb.pitch_b)
    qual = keep["qual" + side][: n * pb.pitch_q].reshape(n, pb.pitch_q)
    seq = np.zeros((n, S), np.uint8); q = np.zeros((n, S), np.uint8)
    lut = np.frombuffer(b"ACTG", np.uint8)
    for k in range(int(L.max(initial=0))):
        code = (bases[:, k >> 2] >> (2 * (k & 3))) & 3
        m = L > k
        seq[m, k] = lut[code[m]]; q[m, k] = qual[m, k]
    npos = keep["npos"][: pb.n_npos * 8].view(np.dtype([("unit", "<u4"), ("pos", "<u2"), ("which", "u1"), ("_pad", "u1")]))
    sel = npos[npos["which"] == (1 if side == "2" else 0)]
    seq[sel["unit"], sel["pos"]] = ord("N")
    assert (np.diff(npos["unit"].astype(np.int64)) >= 0).all()          # sorted by unit
    return seq, q, L


@pytest.mark.parametrize("paired,L,S,threads", [(1, 150, 160, 4), (0, 150, 160, 1), (1, 250, 256, 3), (1, 37, 48, 2)])
def test_pack_roundtrip(paired, L, S, threads):
    n = 20000
    _, arrs = T.synth_host(n, S, paired, 0, 3, 1, L)
    lib = capi.load()
    b = capi.batch_from_arrays(arrs)
    pb, keep = fp_gpu.pack_rows(lib, b, arrs, paired, threads=threads)
    assert pb.n == n and pb.n_npos > 0
    for side in ("1", "2")[: 2 if paired else 1]:
        seq, q, ln = unpack(pb, keep, n, S, side)
        assert (ln == arrs["len" + side]).all()
        cols = np.arange(S)[None, :] < ln[:, None]
        assert (np.where(cols, seq, 0) == np.where(cols, arrs["seq" + side], 0)).all()
        assert (np.where(cols, q, 0) == np.where(cols, arrs["qual" + side], 0)).all()
    assert keep["bytes"] < 0.65 * (2 if paired else 1) * n * (2 * S + 2)          # the point of it


def test_pack_refuses_other_bytes():
    _, arrs = T.synth_host(5000, 160, 0, 0, 3, 1, 150)
    arrs["seq1"][4321, 17] = ord("R")
    lib = capi.load()
    b = capi.batch_from_arrays(arrs)
    with pytest.raises(Exception):
        fp_gpu.pack_rows(lib, b, arrs, 0)


_ISA_WORKER = r'''
import ctypes as C, hashlib, sys
import numpy as np
sys.path.insert(0, "tests")
import fp_testlib as T
from fastp_b200 import capi
import fp_gpu
S = 208
rng = np.random.default_rng(11)
n = 6000
seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, (n, S))].copy()
qual = rng.integers(33, 75, (n, S)).astype(np.uint8)
ln = rng.integers(0, S + 1, n).astype(np.uint16)
ln[:S + 1] = np.arange(S + 1)                          # every row length once, incl. 0 and the block edges 63 / 64 / 65 / 127 / 128 / 129
for r in range(0, n, 9):                               # 'N's: at the block edges, at the row's last base, in runs
    for p in (0, 31, 32, 63, 64, 127, 128, 191, 192, int(ln[r]) - 1):
        if 0 <= p < ln[r]: seq[r, p] = ord("N")
seq[7::197, 60:70] = ord("N")
arrs = {"seq1": seq, "qual1": qual, "len1": ln, "seq2": seq[::-1].copy(), "qual2": qual[::-1].copy(), "len2": ln[::-1].copy()}
lib = capi.load()
b = capi.batch_from_arrays(arrs)
pb, keep = fp_gpu.pack_rows(lib, b, arrs, 1, threads=3)
h = hashlib.sha256()
for k in ("bases1", "bases2", "len1", "len2"): h.update(keep[k].tobytes())
h.update(keep["npos"][: pb.n_npos * 8].tobytes())
bad = []
for pos in (0, 63, 64, 100, 127, 128, 149):           # a byte outside {A,C,G,T,N} anywhere in a row is refused
    a2 = {k: v.copy() for k, v in arrs.items()}
    a2["len1"][:] = 150; a2["seq1"][1234, pos] = ord("R")
    b2 = capi.batch_from_arrays(a2)
    try:
        fp_gpu.pack_rows(lib, b2, a2, 1, threads=2); bad.append(0)
    except Exception:
        bad.append(1)
print("PACKHASH", h.hexdigest(), pb.n_npos, "".join(map(str, bad)))
'''


def test_every_isa_path_packs_the_same_bytes():
    """AVX-512, AVX2 and SWAR packers (FP_HOSTPACK_ISA) on rows of every length with 'N's at the block edges: same packed bytes,
    same 'N' list, same refusals."""
    import os
    import subprocess
    import sys
    outs = {}
    for isa in ("native", "avx2", "swar"):
        env = dict(os.environ)
        env.pop("FP_HOSTPACK_ISA", None)
        if isa != "native":
            env["FP_HOSTPACK_ISA"] = isa
        r = subprocess.run([sys.executable, "-c", _ISA_WORKER], cwd=T.ROOT, env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[isa] = [l for l in r.stdout.splitlines() if l.startswith("PACKHASH")][0]
    assert outs["native"] == outs["avx2"] == outs["swar"], outs
    assert outs["swar"].split()[-1] == "1111111"
