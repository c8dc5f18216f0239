"""gzip / BGZF either side of the text path: fp_gz_inflate / fp_gz_deflate / the streaming reader against python's zlib.
Parity = the DECOMPRESSED bytes (the reference's own .gz output differs from run to run with --thread: one member per pack)."""
import ctypes as C
import gzip
import struct
import zlib

import numpy as np
import pytest

import fp_testlib as T
from fastp_b200 import capi


def bgzf(data, block=60000):
    """BGZF writer (SAM spec 4.1): gzip members with the BC extra subfield, terminated by the empty EOF block."""
    out = bytearray()
    for lo in list(range(0, len(data), block)) + [None]:
        chunk = b"" if lo is None else data[lo:lo + block]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(body) + 8
        out += b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        out += body + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))
    return bytes(out)


def text(n=30000):
    _, arrs = T.synth_host(n, 160, 0, 0, 8, 1, 150)
    return T.fastq_text(arrs["seq1"], arrs["qual1"], arrs["len1"], "1")


def inflate(lib, comp, cap, threads):
    src = np.frombuffer(comp, np.uint8)
    out = np.zeros(cap, np.uint8); n = C.c_int64()
    rc = lib.fp_gz_inflate(src.ctypes.data, src.size, out.ctypes.data, cap, C.byref(n), threads)
    return rc, out[:min(n.value, cap)].tobytes(), n.value


@pytest.mark.parametrize("threads", [1, 4])
def test_inflate_bgzf_multimember_and_plain_gzip(threads):
    lib = capi.load()
    t = text()
    for comp, isb in ((bgzf(t), 1), (gzip.compress(t[:100000]) + gzip.compress(t[100000:]), 0), (gzip.compress(t), 0)):
        assert lib.fp_gz_is_bgzf(np.frombuffer(comp, np.uint8).ctypes.data, len(comp)) == isb
        rc, got, n = inflate(lib, comp, len(t) + 16, threads)
        assert rc == 0 and n == len(t) and got == t
    rc, _, n = inflate(lib, bgzf(t), 1000, threads)
    assert rc == -4 and n == len(t)                       # FP_E_TOOLARGE, BGZF knows its size upfront
    rc, _, _ = inflate(lib, b"\x1f\x8b\x08\x00garbage-not-deflate" * 4, 1 << 16, threads)
    assert rc != 0


@pytest.mark.parametrize("threads,member", [(1, 1 << 20), (4, 200000)])
def test_deflate_members_roundtrip(threads, member):
    lib = capi.load()
    t = text()
    src = np.frombuffer(t, np.uint8)
    cap = lib.fp_gz_deflate_bound(len(t), member)
    out = np.zeros(cap, np.uint8); n = C.c_int64()
    assert lib.fp_gz_deflate(src.ctypes.data, src.size, out.ctypes.data, cap, C.byref(n), member, 4, threads) == 0
    comp = out[:n.value].tobytes()
    assert gzip.decompress(comp) == t                     # python reads the concatenated members as one stream
    assert comp.count(b"\x1f\x8b\x08") >= (len(t) + member - 1) // member
    assert n.value < len(t) // 2


def test_streaming_reader(tmp_path):
    lib = capi.load()
    t = text(8000)
    for name, payload in (("a.fq.gz", gzip.compress(t[:70000]) + gzip.compress(t[70000:])), ("b.fq", t), ("c.fq.gz", bgzf(t))):
        fn = tmp_path / name
        fn.write_bytes(payload)
        h = lib.fp_gz_open(str(fn).encode())
        assert h
        got = b""
        buf = np.zeros(50000, np.uint8)
        while True:
            k = lib.fp_gz_read(h, buf.ctypes.data, buf.size)
            assert k >= 0
            if k == 0:
                break
            got += buf[:k].tobytes()
        lib.fp_gz_close(h)
        assert got == t
