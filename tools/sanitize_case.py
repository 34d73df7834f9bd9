"""Small decode through every kernel for compute-sanitizer runs (one tool per gpurun call):
   compute-sanitizer --tool racecheck python tools/sanitize_case.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import ctcx_testlib as L
import ctc_beam_search_op_b200 as op

for impl in ("fast", "generic"):
    op.set_beam_impl(impl)
    for (kind, T, B, C, W, P, blank) in (("peaky", 24, 3, 29, 100, 2, 28), ("gauss", 16, 2, 12, 8, 3, 0),
                                         ("gauss", 12, 2, 40, 12, 1, 39)):  # narrow tiers, wide kernel
        x = L.make_logits(kind, T, B, C, blank, 5)
        sl = L.ragged_lengths(T, B, 5)
        for inp in (torch.from_numpy(x).cuda(), torch.from_numpy(x).pin_memory()):  # device-resident and overlapped feed
            raw = op.ctc_ext_beam_search_decoder_raw(inp, sl, beam_width=W, top_paths=P, merge_repeated=True, blank_index=blank)
            assert not L.raw_mismatches(raw, L.oracle_decode(x, sl, W, P, True, blank, -1))
op.set_beam_impl(None)
x = L.make_logits("peaky", 20, 2, 29, 28, 7)
dec = op.CTCExtBeamSearchDecoderStream(2, 29, 50, 1, max_time=20, blank_index=28)
dec.step(x[:7]); dec.step(x[7:20]); dec.top_paths()
print("sanitize case ok")
