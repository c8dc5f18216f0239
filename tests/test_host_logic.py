"""CPU-only checks of the C++ host side above the C-ABI (fastp_b200/host): the Options -> fp_params mapping and the unpacking of the
counter block into the reference-shaped Stats objects.  No device call is made; the block comes from the CPU oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import fp_testlib as T
from fastp_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "fastp_b200", "host")
LIBDIR = os.path.join(ROOT, "fastp_b200")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if not os.path.exists(os.path.join(LIBDIR, "libfastp_b200.so")):
        pytest.skip("libfastp_b200.so not built")
    out = tmp_path_factory.mktemp("hostlogic") / "host_logic_check"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", HOST, os.path.join(ROOT, "tests", "host", "host_logic_check.cpp"),
                    os.path.join(HOST, "gpu_worker.cpp"), "-o", str(out), "-L", LIBDIR, "-lfastp_b200", "-Wl,-rpath," + LIBDIR], check=True)
    return str(out)


def test_options_to_params_mapping(exe):
    kv = dict(l.split("=", 1) for l in subprocess.run([exe, "params"], check=True, capture_output=True, text=True).stdout.strip().split("\n"))
    want = dict(paired=1, thread0_semantics=1, trim_front1=1, trim_tail1=2, trim_front2=3, trim_tail2=4, max_len1=101, max_len2=102,
                cut_front=1, cut_tail=0, cut_right=1, cut_front_window=5, cut_front_quality=21, cut_tail_window=6, cut_tail_quality=22,
                cut_right_window=7, cut_right_quality=23, polyg_enabled=1, polyg_min_len=11, polyx_enabled=1, polyx_min_len=12,
                adapter_enabled=1, has_seq_r1=1, has_seq_r2=1, n_fasta_adapters=2, allow_gap_overlap_trimming=1, dimer_max_len=3,
                correction_enabled=1, overlap_require=31, overlap_diff_limit=4, overlap_diff_percent_limit=19,
                qual_filter_enabled=1, qualified_qual=ord("5"), unqualified_percent_limit=41, n_base_limit=6, avg_qual_req=13,
                length_filter_enabled=1, length_required=16, length_limit=140, complexity_filter_enabled=1,
                insert_size_max=600, seq_len1=151, seq_len2=149, overrep_enabled=1, overrep_sampling=7, n_overrep1=2, n_overrep2=1)
    for k, v in want.items():
        assert int(kv[k]) == v, k
    assert abs(float(kv["complexity_threshold"]) - 0.25) < 1e-9
    assert kv["adapter_seq_r1"] == "AGATCGGAAGAGC" and kv["adapter_seq_r2"] == "CTGTCTCTTATA"
    assert [kv["fasta0"], kv["fasta1"]] == ["AAAACCCC", "GGGGTTTTAA"]
    assert sorted([kv["ovr1_0"], kv["ovr1_1"]]) == ["ACGTACGTAC", "TTTTTTTTTTGG"] and kv["ovr2_0"] == "CCCCCCCCCC"


@pytest.mark.parametrize("paired", [1, 0])
def test_stats_fill_from_counter_block(exe, tmp_path, paired):
    p = T.config_params("cfg4_full", paired)
    _, arrs = T.synth_host(3000, 160, paired, 0, 5, 1, 150)
    res = T.run_cpu("oracle", p, arrs, 160)
    cv = res["counters"]
    path = tmp_path / "block.bin"
    with open(path, "wb") as f:
        f.write(bytes(cv.L)); f.write(np.ascontiguousarray(cv.data, np.int64).tobytes())
    out = subprocess.run([exe, "stats", str(path)], check=True, capture_output=True, text=True).stdout.strip().split("\n")
    assert len(out) == (4 if paired else 2) + 1
    fl = dict(x.split("=") for x in out[-1].split()[1:])
    F = cv.filter
    assert int(fl["pass"]) == F[capi.PASS_FILTER] and int(fl["lowq"]) == F[capi.FAIL_QUALITY] and int(fl["nbase"]) == F[capi.FAIL_N_BASE]
    assert int(fl["tooshort"]) == F[capi.FAIL_LENGTH]
    assert int(fl["adapter_reads"]) == F[capi.FR_ADAPTER_READS] and int(fl["adapter_bases"]) == F[capi.FR_ADAPTER_BASES]
    assert int(fl["corrected_reads"]) == F[capi.FR_CORRECTED_READS] and int(fl["corrections"]) == int(F[capi.FR_CORRECTION:capi.FR_CORRECTION + 64].sum())
    assert int(fl["polyx_reads"]) == int(F[capi.FR_POLYX_READS:capi.FR_POLYX_READS + 4].sum()) and int(fl["polyx_bases"]) == int(F[capi.FR_POLYX_BASES:capi.FR_POLYX_BASES + 4].sum())
    assert int(fl["adapter_reads"]) > 0 and int(fl["polyx_reads"]) > 0
    for line in out[:-1]:
        f = dict(x.split("=") for x in line.split()[1:])
        which = int(line.split()[0][5:])
        s, st = cv.summary(which), cv.stats(which)
        assert int(f["reads"]) == s["reads"] and int(f["bases"]) == s["bases"] and int(f["q20"]) == s["q20"] and int(f["q30"]) == s["q30"]
        assert int(f["cycles"]) == s["cycles"]
        assert int(f["kmer"]) == int(st["kmer"].sum()) and int(f["qualhist"]) == int(st["qualhist"].sum())
        assert int(f["tq0"]) == int(st["cycle"][33][0]) and int(f["tb0"]) == int(st["cycle"][32][0])
        assert int(f["A5"]) == int(st["cycle"][2 * 8 + (ord("A") & 7)][5])
