"""Golden fixtures generated from the REFERENCE build (tests/golden/make_golden.py): the CPU oracle must
reproduce every per-read record, overlap record, corrected base and counter.  (-m gpu twin: test_gpu_golden.py)"""
import glob
import json
import os

import numpy as np
import pytest

import fp_testlib as T
from fastp_b200 import capi

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def load_fixture(path):
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    if meta["config"] == "testdata_cli_defaults":
        p = capi.default_params(1, lib=T.oracle(), polyg_enabled=1, seq_len1=151, seq_len2=151)
        arrs = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    elif "overrep_sampling" in meta:
        _, arrs = T.synth_host(meta["n"], meta["stride"], meta["paired"], 0, meta["seed"], meta["profile"], meta["read_len"])
        p = T.overrep_params(meta["base_config"], meta["paired"], arrs, meta["read_len"], meta["overrep_sampling"])
    else:
        p = T.config_params(meta["config"], meta["paired"])
        _, arrs = T.synth_host(meta["n"], meta["stride"], meta["paired"], 0, meta["seed"], meta["profile"], meta["read_len"])
    cnt = np.zeros(int(z["counters_total"]), np.int64)
    cnt[z["counters_idx"]] = z["counters_val"]
    after = {k: v.copy() for k, v in arrs.items()}
    for k in ("seq1", "qual1", "seq2", "qual2"):
        if "patch_" + k in z.files:
            for r, c, v in z["patch_" + k]:
                after[k][r, c] = v
    L = capi.make_layout(T.oracle(), meta["paired"], meta["cycles"], p.insert_size_max, p)
    want = {"out1": z["out1"], "out2": z["out2"], "ov": z["ov"], "counters": capi.CounterView(L, cnt), "arrs": after}
    return meta, p, arrs, want


def test_fixture_count():
    assert len(GOLDEN) >= 36


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_reference_golden(path):
    meta, p, arrs, want = load_fixture(path)
    got = T.run_cpu("oracle", p, arrs, meta["cycles"])
    # adapter_pos is not observable through the reference's API; everything else must match bit for bit
    T.assert_results_equal(got, want, meta["paired"], skip=("adapter_pos",), what=os.path.basename(path))


def test_testdata_matches_reference_cli_report():
    """config 1 (testdata/R1.fq + R2.fq, defaults, --thread 1): numbers of the unmodified reference CLI's JSON."""
    meta, p, arrs, want = load_fixture(os.path.join(os.path.dirname(__file__), "golden", "testdata_pe.npz"))
    cli = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "testdata_cli.json")))
    got = T.run_cpu("oracle", p, arrs, meta["cycles"])
    c = got["counters"]
    pre = [c.summary(capi.STATS_PRE1), c.summary(capi.STATS_PRE2)]
    post = [c.summary(capi.STATS_POST1), c.summary(capi.STATS_POST2)]
    bf, af = cli["summary"]["before_filtering"], cli["summary"]["after_filtering"]
    assert sum(s["reads"] for s in pre) == bf["total_reads"] == 18
    assert sum(s["bases"] for s in pre) == bf["total_bases"] == 2567
    assert sum(s["q20"] for s in pre) == bf["q20_bases"] and sum(s["q30"] for s in pre) == bf["q30_bases"]
    assert sum(s["reads"] for s in post) == af["total_reads"] == 16
    assert sum(s["bases"] for s in post) == af["total_bases"] and sum(s["q20"] for s in post) == af["q20_bases"]
    fr = cli["filtering_result"]
    assert c.filter[capi.PASS_FILTER] == fr["passed_filter_reads"] and c.filter[capi.FAIL_LENGTH] == fr["too_short_reads"] == 2
    assert c.filter[capi.FAIL_QUALITY] == fr["low_quality_reads"] and c.filter[capi.FAIL_N_BASE] == fr["too_many_N_reads"]
    assert int(np.argmax(c.isize[:-1])) == cli["insert_size"]["peak"] == 187
    assert c.isize[-1] == cli["insert_size"]["unknown"]
    assert list(c.isize[:len(cli["insert_size"]["histogram"])]) [:512] == cli["insert_size"]["histogram"][:512]
    # the reference's own definition of parity: byte-identical output reads (scripts/bench_e2e.sh:183-225)
    keep = [i for i in range(meta["n"]) if got["out1"]["pair_verdict"][i] == 0]
    for side, key in (("1", "out1"), ("2", "out2")):
        outs = []
        for i in keep:
            r = got[key][i]
            s = bytes(got["arrs"]["seq" + side][i, r["front"]: r["front"] + r["len"]]).decode()
            q = bytes(got["arrs"]["qual" + side][i, r["front"]: r["front"] + r["len"]]).decode()
            outs.append([s, q])
        assert outs == cli[key]
