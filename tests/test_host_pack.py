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
