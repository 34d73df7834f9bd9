"""Generates tests/golden/ref_cases.npz from the REFERENCE ITSELF: oracle/_ref/libctcx_ref.so is the
reference's unmodified decoder headers compiled by oracle/Makefile (only possible where
/root/reference exists). The fixture lets the oracle, the model and the CUDA path be checked against
the reference on machines that have neither the reference tree nor the compiled library.

    python tests/golden/make_golden.py

Inputs are not stored: each case is regenerated from (kind, shape, seed) by ctcx_testlib.make_logits
(or is one of the literal tables below); only the reference's outputs are stored.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ctcx_testlib as L  # noqa: E402

# (name, kind, T, B, C, W, P, merge, blank_index, blank_label, seed, sigma, ragged, dtype)
RANDOM_CASES = [
    ("cfg1_gauss", "gauss", 50, 8, 29, 10, 3, False, 28, -1, 0, 1.0, False, "f32"),
    ("cfg1_peaky", "peaky", 50, 8, 29, 10, 3, False, 28, -1, 0, 1.0, False, "f32"),
    ("cfg1_peaky_ragged_merge", "peaky", 50, 8, 29, 10, 3, True, 28, -1, 5, 1.0, True, "f32"),
    ("cfg2_short_peaky", "peaky", 80, 4, 29, 100, 1, True, 28, -1, 1, 1.0, False, "f32"),
    ("cfg2_short_gauss", "gauss", 80, 4, 29, 100, 1, True, 28, -1, 1, 1.0, True, "f32"),
    ("cfg3_short_peaky", "peaky", 60, 4, 32, 64, 4, False, 31, -1, 2, 1.0, False, "f32"),
    ("cfg4_short_peaky", "peaky", 30, 2, 1024, 16, 1, False, 1023, -1, 3, 1.0, False, "f32"),
    ("blank_mid_label9", "gauss", 25, 6, 7, 5, 5, True, 3, 9, 7, 2.0, True, "f32"),
    ("beam1", "peaky", 40, 6, 12, 1, 1, False, 0, -1, 8, 1.0, False, "f32"),
    ("f64_peaky", "peaky", 40, 4, 6, 4, 2, False, 3, -1, 3, 1.0, False, "f64"),
]

# literal edge cases measured on the compiled reference
ROWS_D = [[1, 2, 3], [3, 1, 0], [0, 0, 5], [2, 2, 2]]
LITERAL_CASES = [
    # name, rows, seq_len, W, P, merge, blank_index, blank_label
    ("d_t1_w4_p3", ROWS_D[:1], 1, 4, 3, False, 0, -1),
    ("d_t4_w1", ROWS_D, 4, 1, 1, False, 0, -1),
    ("d_blank1_label9", ROWS_D, 4, 4, 4, False, 1, 9),
    ("d_blank1_label9_merge", ROWS_D, 4, 4, 4, True, 1, 9),
    ("d_seq0", ROWS_D, 0, 4, 1, False, 0, -1),
    ("d_spiky", [[5, 0, 0]] * 3 + [[0, 0, 5]], 4, 4, 2, False, 1, -1),
    ("d_seq2_of_4", ROWS_D, 2, 3, 2, False, 2, -1),
]


def case_inputs(c):
    name, kind, T, B, C, W, P, merge, blank, bl, seed, sigma, ragged, dt = c
    x = L.make_logits(kind, T, B, C, blank, seed, sigma)
    if dt == "f64":
        x = x.astype(np.float64)
    sl = L.ragged_lengths(T, B, seed) if ragged else np.full(B, T, np.int32)
    return x, sl


def literal_inputs(c):
    name, rows, sl, W, P, merge, blank, bl = c
    x = np.asarray(rows, np.float32)[:, None, :]
    return x, np.asarray([sl], np.int32)


def main():
    L.build_oracles()
    assert L.have_ref(), "needs oracle/_ref (build where /root/reference exists)"
    out, meta = {}, {"random": [], "literal": []}
    for c in RANDOM_CASES:
        x, sl = case_inputs(c)
        r = L.ref_decode(x, sl, c[5], c[6], c[7], c[8], c[9])
        _, margins = L.oracle_decode(x, sl, c[5], c[6], c[7], c[8], c[9], want_margin=True)
        for k, v in (("dec_len", r.dec_len), ("dec", r.dec), ("ali_len", r.ali_len), ("ali", r.ali),
                     ("logp", r.logp), ("tie_free", (margins[:, [1, 2, 4]].min(axis=1) > 0))):
            out["%s/%s" % (c[0], k)] = v
        meta["random"].append(list(c))
    for c in LITERAL_CASES:
        x, sl = literal_inputs(c)
        r = L.ref_decode(x, sl, c[3], c[4], c[5], c[6], c[7])
        for k, v in (("dec_len", r.dec_len), ("dec", r.dec), ("ali_len", r.ali_len), ("ali", r.ali),
                     ("logp", r.logp)):
            out["%s/%s" % (c[0], k)] = v
        meta["literal"].append([c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]])
    # the reference's own known-answer test, float64 and float32 (ops_test.py:25-64)
    for dt, nm in ((np.float64, "paper_f64"), (np.float32, "paper_f32")):
        r = L.ref_decode(L.paper_logits(dt), [8], 10, 5, False, 0, 0)
        for k, v in (("dec_len", r.dec_len), ("dec", r.dec), ("ali_len", r.ali_len), ("ali", r.ali),
                     ("logp", r.logp)):
            out["%s/%s" % (nm, k)] = v
    # per-frame beam of the W=3 trace
    tr = L.ref_trace(L.paper_logits(np.float32)[:, 0, :], 3, False, 0, 0)
    meta["trace_w3"] = [[[float(np.float32(lp)), pre, ali] for lp, pre, ali in row] for row in tr]
    np.savez_compressed(os.path.join(HERE, "ref_cases.npz"), **out)
    with open(os.path.join(HERE, "ref_cases.json"), "w") as f:
        json.dump(meta, f, indent=0)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
