"""Full-batch parity soak: every utterance of the BASELINE shapes (cfg1-cfg4, Gaussian and peaky
logits, float32; cfg2 / cfg3 also in float64, cfg2 with a random scorer table) decoded on the GPU and by the
CPU oracle (oracle/, sharded over host threads), compared bit for bit -- labels, alignments, the
IEEE bits of log_probability.   python tools/soak_parity.py [threads] [n_seeds] > profiles/<round>_soak.txt
The oracle is the checker here, never the thing measured."""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import ctcx_testlib as L
import ctc_beam_search_op_b200 as op

NT = int(sys.argv[1]) if len(sys.argv) > 1 else min(64, os.cpu_count() or 1)
NSEEDS = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L.build_oracles()
CFGS = [("cfg1", 50, 8, 29, 10, 3, False, 28), ("cfg2", 500, 256, 29, 100, 1, True, 28),
        ("cfg3", 1500, 64, 32, 64, 4, False, 31), ("cfg4", 400, 128, 1024, 16, 1, False, 1023)]


def oracle_sharded(x, sl, W, P, merge, blank, lm=None):
    B = x.shape[1]
    n = min(NT, B)
    bounds = [(B * i // n, B * (i + 1) // n) for i in range(n)]

    def one(bb):
        b0, b1 = bb
        return L.oracle_decode(np.ascontiguousarray(x[:, b0:b1]), sl[b0:b1], W, P, merge, blank, -1, lm=lm)
    with ThreadPoolExecutor(n) as ex:
        return list(zip(bounds, ex.map(one, bounds)))


def dense_from_raw(raw, B, P):
    dec = [[[] for _ in range(P)] for _ in range(B)]
    ali = [[[] for _ in range(P)] for _ in range(B)]
    for p in range(P):
        for (b, _), v in zip(np.asarray(raw[0][p]).tolist(), np.asarray(raw[1][p]).tolist()):
            dec[b][p].append(v)
        for (b, _), v in zip(np.asarray(raw[3][p]).tolist(), np.asarray(raw[4][p]).tolist()):
            ali[b][p].append(v)
    return dec, ali


total_utt = total_frames = total_bad = 0
print("threads for the oracle: %d" % NT)
for seed, kind in [(21 + 100 * k, kd) for k in range(NSEEDS) for kd in ("gauss", "peaky")]:
    for name, T, B, C, W, P, merge, blank in CFGS:
        variants = [("f32", np.float32, None)]
        if name == "cfg3":
            variants += [("f64", np.float64, None)]
        if name == "cfg2":
            variants += [("f64", np.float64, None),
                         ("f32+scorer", np.float32,
                          -np.abs(np.random.default_rng(seed + 2).standard_normal((C + 1, C))).astype(np.float32))]
        for vname, dt, lm in variants:
            x = L.make_logits(kind, T, B, C, blank, seed).astype(dt)
            if dt == np.float64:  # genuine double values
                x = x + np.random.default_rng(seed + 1).standard_normal(x.shape) * 1e-9
            sl = L.ragged_lengths(T, B, seed)
            t0 = time.time()
            raw = op.ctc_ext_beam_search_decoder_raw(x, sl, beam_width=W, top_paths=P, merge_repeated=merge,
                                                     blank_index=blank, blank_label=-1, expansion_scores=lm)
            t_gpu = time.time() - t0
            t0 = time.time()
            parts = oracle_sharded(x, sl, W, P, merge, blank, lm)
            t_cpu = time.time() - t0
            dec, ali = dense_from_raw(raw, B, P)
            lp = np.asarray(raw[6])
            view = np.uint64 if dt == np.float64 else np.uint32
            bad = 0
            for (b0, b1), r in parts:
                for b in range(b0, b1):
                    for p in range(P):
                        ok = (dec[b][p] == r.decoded(b - b0, p) and ali[b][p] == r.alignment(b - b0, p) and
                              lp[b, p].view(view) == np.asarray(r.logp[b - b0, p]).view(view))
                        bad += 0 if ok else 1
            total_utt += B
            total_frames += int(sl.sum())
            total_bad += bad
            print("seed %4d %-5s %-5s %-10s T=%4d B=%3d C=%4d W=%3d P=%d: %6d frames, mismatching (utterance,path) pairs: %d"
                  "   [gpu call %.3f s, oracle %.1f s, flags %d]"
                  % (seed, kind, name, vname, T, B, C, W, P, int(sl.sum()), bad, t_gpu, t_cpu, raw.flags))
print("TOTAL %d utterances, %d frames, %d mismatches" % (total_utt, total_frames, total_bad))
sys.exit(1 if total_bad else 0)
