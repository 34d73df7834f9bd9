"""Small decode through every kernel for compute-sanitizer runs: python tools/sanitize_case.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import ctcx_testlib as L
import ctc_beam_search_op_b200 as op

for impl in ("", "v2", "generic"):
    if impl:
        os.environ["CTCX_BEAM_IMPL"] = impl
    else:
        os.environ.pop("CTCX_BEAM_IMPL", None)
    for (kind, T, B, C, W, P, blank) in (("peaky", 24, 3, 29, 100, 2, 28), ("gauss", 16, 2, 12, 8, 3, 0)):
        x = L.make_logits(kind, T, B, C, blank, 5)
        sl = L.ragged_lengths(T, B, 5)
        raw = op.ctc_ext_beam_search_decoder_raw(x, sl, beam_width=W, top_paths=P, merge_repeated=True, blank_index=blank)
        want = L.pack_sparse(L.oracle_decode(x, sl, W, P, True, blank, -1))
        assert all(np.array_equal(np.asarray(raw[g][p]), want[g][p]) for g in range(6) for p in range(P))
os.environ.pop("CTCX_BEAM_IMPL", None)
x = L.make_logits("gauss", 10, 2, 40, 39, 6)  # generic kernel, C > 32
op.ctc_ext_beam_search_decoder_raw(x, [10, 7], beam_width=12, top_paths=1, blank_index=39)
x = L.make_logits("peaky", 20, 2, 29, 28, 7)
dec = op.CTCExtBeamSearchDecoderStream(2, 29, 50, 1, max_time=20, blank_index=28)
dec.step(x[:7]); dec.step(x[7:20]); dec.top_paths()
print("sanitize case ok")
