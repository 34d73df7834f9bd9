"""Collects concrete utterances in which the reference ACCEPTS the re-score of an evicted member
(decoder.h:167-199) -- found by the search in tools/anomaly_search.py -- checks the oracle against the
compiled reference on them and writes them to tests/golden/rescore_cases.npz (logits + shape + the
reference's outputs).   python tools/anomaly_cases.py [n_cases]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np

import ctcx_testlib as L
from anomaly_search import make

want = int(sys.argv[1]) if len(sys.argv) > 1 else 24
L.build_oracles()
rng = np.random.default_rng(4242)
cases = []
tried = 0
while len(cases) < want and tried < 200000:
    tried += 1
    C = int(rng.integers(2, 9)); W = int(rng.integers(1, 9)); T = int(rng.integers(3, 60)); B = 64
    blank = int(rng.integers(0, C)); merge = bool(rng.integers(0, 2))
    x = make(rng, T, B, C)
    sl = np.full(B, T, np.int32)
    try:
        _, st = L.oracle_decode(x, sl, W, 1, merge, blank, -1, want_stats=True)
    except L.OracleError:
        continue
    if not st.revisit_accepts:
        continue
    for b in range(B):  # which utterance(s)?
        xb = np.ascontiguousarray(x[:, b:b + 1])
        P = min(W, 3)
        try:
            r, mg, s1 = L.oracle_decode(xb, sl[:1], W, P, merge, blank, -1, want_margin=True, want_stats=True)
        except L.OracleError:
            continue
        if s1.revisit_accepts:
            cases.append((xb[:, 0, :].copy(), W, P, merge, blank, r, mg[0]))
            if len(cases) >= want:
                break
print("found %d cases in %d batches" % (len(cases), tried))
out = {}
agree = 0
for k, (x, W, P, merge, blank, r, mg) in enumerate(cases):
    out["c%d/x" % k] = x
    out["c%d/attrs" % k] = np.asarray([W, P, int(merge), blank], np.int32)
    if L.have_ref():
        ref = L.ref_decode(x[:, None, :], np.asarray([x.shape[0]], np.int32), W, P, merge, blank, -1)
        same = not L.same_result(ref, r)
        agree += int(same)
        out["c%d/ref_logp" % k] = ref.logp
        out["c%d/ref_dec_len" % k] = ref.dec_len
        out["c%d/ref_dec" % k] = ref.dec
        out["c%d/ref_ali_len" % k] = ref.ali_len
        out["c%d/ref_ali" % k] = ref.ali
        out["c%d/oracle_equals_ref" % k] = np.asarray([int(same)])
        out["c%d/tie_free" % k] = np.asarray([int(mg[[1, 2, 4]].min() > 0)])
out["n"] = np.asarray([len(cases)])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rescore_cases.npz"), **out)
print("oracle == compiled reference on %d of %d cases" % (agree, len(cases)))
print("tie-free cases: %d" % sum(int(out["c%d/tie_free" % k][0]) for k in range(len(cases))))
