"""The device evaluates Matcher::matchWithOneInsertion (src/matcher.cpp:10-54) in closed form:
     exists i in [1,c-1]:  P1[i] - P2[i] + P2[c] <= limit,   D1[j] = ins[j]!=norm[j], D2[j] = ins[j+1]!=norm[j]
(fp_device.cuh dev_gap_scan).  Check it against the literal restatement in the oracle on random and
adversarial strings, for every (cmplen, limit) the adapter scans can produce."""
import numpy as np

import fp_testlib as T


def closed_form(ins, norm, c, limit):
    d1 = [int(ins[j] != norm[j]) for j in range(c)]
    d2 = [int(ins[j + 1] != norm[j]) for j in range(c)]
    p1 = np.concatenate([[0], np.cumsum(d1)])
    p2 = np.concatenate([[0], np.cumsum(d2)])
    return int(any(p1[i] - p2[i] + p2[c] <= limit for i in range(1, c)))


def test_closed_form_equals_literal_matcher():
    rng = np.random.default_rng(1)
    o = T.oracle()
    checked = hits = 0
    for trial in range(3000):
        n = int(rng.integers(6, 70))
        base = rng.integers(0, 4, n + 2)
        alpha = np.frombuffer(b"ACGT", np.uint8)
        ins = alpha[base].copy()
        norm = ins.copy()
        mode = trial % 4
        if mode == 0:      # unrelated strings
            norm = alpha[rng.integers(0, 4, n + 2)]
        elif mode == 1:    # one insertion in `ins`
            k = int(rng.integers(1, n))
            norm = np.concatenate([ins[:k], ins[k + 1:], alpha[rng.integers(0, 4, 1)]])
        # sprinkle substitutions
        for _ in range(int(rng.integers(0, 4))):
            norm[int(rng.integers(0, n))] = alpha[int(rng.integers(0, 4))]
        for c in range(3, n + 1):
            for limit in (c // 8 - 1, c // 8, -1, 0, 1, 2):
                want = o.fp_oracle_match_with_one_insertion(ins.ctypes.data, norm.ctypes.data, c, limit)
                got = closed_form(ins, norm, c, limit)
                assert want == got, (trial, c, limit, bytes(ins), bytes(norm))
                checked += 1
                hits += want
    assert checked > 100000 and hits > 1000
