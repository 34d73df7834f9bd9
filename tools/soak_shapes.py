"""Random-shape parity soak: many small decodes with random (T, B, C, beam_width, top_paths, blank,
merge) -- narrow, wide and very wide vocabularies, all beam-width tiers, Gaussian / peaky / quantised
(tie-heavy) / micro-spaced / masked logits, float32 and float64, with and without a scorer table,
inputs as numpy arrays, page-locked host tensors (overlapped / slab-wise host feed) or CUDA tensors --
each compared bit for bit with the CPU oracle.   python tools/soak_shapes.py [n_cases] [seed]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import ctcx_testlib as L
import ctc_beam_search_op_b200 as op

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
L.build_oracles()
bad_cases, n_err, frames = [], 0, 0
kinds = ["gauss", "peaky", "quant", "micro", "masked"]
for case in range(N):
    C = int(rng.choice([2, 3, 5, 12, 29, 32, 33, 40, 64, 65, 100, 257, 700, 1024, 2048, 3000]))
    W = int(rng.choice([1, 2, 3, 7, 16, 31, 32, 33, 64, 100, 128, 129, 200, 256, 257, 400]))
    if W * C > 300000:
        W = max(1, 300000 // C)
    T = int(rng.integers(1, 70))
    B = int(rng.integers(1, 5))
    P = int(rng.integers(1, min(W, 4) + 1))
    blank = int(rng.integers(0, C))
    merge = bool(rng.integers(0, 2))
    kind = kinds[int(rng.integers(0, len(kinds)))]
    f64 = bool(rng.random() < 0.2) and W <= 512
    scorer = (not f64) and bool(rng.random() < 0.2)
    seed = int(rng.integers(0, 1 << 30))
    r2 = np.random.default_rng(seed)
    if kind in ("gauss", "peaky"):
        x = L.make_logits(kind, T, B, C, blank, seed, float(rng.choice([0.5, 1, 3])))
    elif kind == "quant":
        x = r2.integers(0, int(rng.choice([2, 3, 5])), (T, B, C)).astype(np.float32) * float(rng.choice([0.5, 1.0, 2.0]))
    elif kind == "micro":
        x = (r2.integers(0, 6, (T, B, C)) * float(rng.choice([1e-6, 3e-7, 1e-5]))).astype(np.float32)
        if rng.random() < 0.5:
            x = x + (np.arange(C, dtype=np.float64) * 1e-6).astype(np.float32)
    else:
        x = L.make_logits("gauss", T, B, C, blank, seed)
        m = r2.random((T, B, C)) < 0.4
        m[..., blank] = False
        x = np.where(m, -np.inf, x).astype(np.float32)
    if f64:
        x = x.astype(np.float64)
        if kind in ("gauss", "peaky"):
            x = x + r2.standard_normal(x.shape) * 1e-9
    sl = r2.integers(0 if P == 1 else 1, T + 1, B).astype(np.int32)
    lm = (-np.abs(r2.standard_normal((C + 1, C))).astype(np.float32)) if scorer else None
    place = ["numpy", "pinned", "cuda"][int(rng.integers(0, 3))]
    tag = "case %d: %s T=%d B=%d C=%d W=%d P=%d blank=%d merge=%d f64=%d scorer=%d %s seed=%d" % (
        case, kind, T, B, C, W, P, blank, merge, f64, scorer, place, seed)
    if place == "numpy":
        xin, sin = x, sl
    else:
        import torch
        xin = torch.from_numpy(x).pin_memory() if place == "pinned" else torch.from_numpy(x).cuda()
        sin = sl if place == "pinned" else torch.from_numpy(sl).cuda()
    try:
        want = L.oracle_decode(x, sl, W, P, merge, blank, -1, lm=lm)
    except L.OracleError as e:
        try:
            op.ctc_ext_beam_search_decoder_raw(xin, sin, beam_width=W, top_paths=P, merge_repeated=merge,
                                               blank_index=blank, expansion_scores=lm)
            bad_cases.append(tag + " -> oracle error '%s' but the GPU path succeeded" % e)
        except op.CtcxError:
            n_err += 1
        continue
    try:
        raw = op.ctc_ext_beam_search_decoder_raw(xin, sin, beam_width=W, top_paths=P, merge_repeated=merge,
                                                 blank_index=blank, blank_label=-1, expansion_scores=lm)
    except op.CtcxError as e:
        bad_cases.append(tag + " -> GPU error %s" % e)
        continue
    packed = L.pack_sparse(want)
    view = np.uint64 if f64 else np.uint32
    def npy(a):
        return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    ok = all(np.array_equal(npy(raw[g][p]), packed[g][p]) for g in range(6) for p in range(P))
    ok = ok and np.array_equal(npy(raw[6]).view(view), np.asarray(packed[6]).view(view))
    frames += int(sl.sum())
    if not ok:
        bad_cases.append(tag + " -> MISMATCH")
    if (case + 1) % 100 == 0:  # progress survives a time limit
        print("... %d cases, %d frames, %d failures so far" % (case + 1, frames, len(bad_cases)), flush=True)
print("%d cases, %d frames, %d expected errors (both sides), %d failures" % (N, frames, n_err, len(bad_cases)))
for b in bad_cases[:20]:
    print(b)
sys.exit(1 if bad_cases else 0)
