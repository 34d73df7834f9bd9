"""CPU search for the one case the beam kernels flag instead of modelling (DESIGN.md "Known deviation"):
a member that was evicted during the grow phase, is re-scored by its parent (decoder.h:167-187) and
ACCEPTED again because rounding made the re-score exceed the beam bottom (decoder.h:189-199). The oracle
counts these events (revisit_accepts) and the necessary condition the kernels test (revisit_above_former).
  python tools/anomaly_search.py [seconds] [processes]     -- test infrastructure, runs on the host only"""
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import ctcx_testlib as L


def make(rng, T, B, C):
    kind = rng.integers(0, 9)
    if kind >= 6:   # label-symmetric inputs: all labels share one logit per frame, so label-permuted prefixes tie
        lab = rng.standard_normal((T, B, 1)) * rng.choice([1, 5, 15, 30])
        blk = rng.standard_normal((T, B, 1)) * rng.choice([1, 5, 15, 30])
        x = np.repeat(lab, C, axis=2)
        if kind == 7:   # two groups of labels
            x[..., : C // 2] += rng.choice([1e-6, 0.5, 3.0])
        if kind == 8:   # quantised
            x = np.round(x * 2) / 2
            blk = np.round(blk * 2) / 2
        x[..., 0:1] = blk   # (class 0 plays the blank when blank == 0; any class otherwise)
        return x.astype(np.float32)
    if kind == 0:   # flat Gaussian, assorted scales
        return (rng.standard_normal((T, B, C)) * rng.choice([0.3, 1, 3, 8])).astype(np.float32)
    if kind == 1:   # few quantised levels: many exact ties and near ties
        return (rng.integers(0, rng.integers(2, 5), (T, B, C)) * rng.choice([1e-6, 1e-3, 0.5, 2.0])).astype(np.float32)
    if kind == 2:   # large offsets: scores of magnitude 1e3..1e5, coarse ULPs
        return (rng.standard_normal((T, B, C)) + rng.choice([1e3, 1e4, 1e5])).astype(np.float32)
    if kind == 3:   # spikes: one class dominates by 10..40 nats (absorbed LogSumExp terms)
        x = rng.standard_normal((T, B, C)).astype(np.float32)
        idx = rng.integers(0, C, (T, B))
        np.put_along_axis(x, idx[..., None], x.max() + rng.choice([10, 20, 40]), axis=2)
        return x
    if kind == 4:   # micro-spaced classes
        return (np.arange(C)[None, None, :] * rng.choice([1e-7, 1e-6, 1e-5]) +
                rng.integers(0, 2, (T, B, 1)) * 1.0).astype(np.float32)
    return (rng.standard_normal((T, B, C)) * 20).astype(np.float32)  # extreme spread


def work(args):
    seed, seconds = args
    rng = np.random.default_rng(seed)
    t_end = time.time() + seconds
    frames = accepts = above = wipes = 0
    found = []
    while time.time() < t_end:
        C = int(rng.integers(2, 9)); W = int(rng.integers(1, 9)); T = int(rng.integers(3, 60)); B = 64
        blank = int(rng.integers(0, C))
        x = make(rng, T, B, C)
        sl = np.full(B, T, np.int32)
        try:
            _, st = L.oracle_decode(x, sl, W, 1, bool(rng.integers(0, 2)), blank, -1, want_stats=True)
        except L.OracleError:
            continue
        frames += st.frames; accepts += st.revisit_accepts; above += st.revisit_above_former; wipes += st.wipes
        if st.revisit_accepts or st.revisit_above_former:
            found.append((seed, T, C, W, blank, int(st.revisit_accepts), int(st.revisit_above_former)))
    return frames, wipes, above, accepts, found[:5]


if __name__ == "__main__":
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    L.build_oracles()
    with ProcessPoolExecutor(procs) as ex:
        res = list(ex.map(work, [(1000 + i, seconds) for i in range(procs)]))
    frames = sum(r[0] for r in res); wipes = sum(r[1] for r in res); above = sum(r[2] for r in res); acc = sum(r[3] for r in res)
    print("frames %d, revisit-wipes %d, re-score above the former total %d, re-score ACCEPTED %d" % (frames, wipes, above, acc))
    for r in res:
        for f in r[4]:
            print("  case seed=%d T=%d C=%d W=%d blank=%d accepts=%d above=%d" % f)
