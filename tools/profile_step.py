"""One decode of the bench workload (BASELINE configs[1]) for ncu captures:
   python tools/profile_step.py [B] [kind] [reps] [f32|f64|scorer]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import ctcx_testlib as L
import ctc_beam_search_op_b200 as op

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kind = sys.argv[2] if len(sys.argv) > 2 else "gauss"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
T, C, W = 500, 29, 100
variant = sys.argv[4] if len(sys.argv) > 4 else "f32"
x = torch.from_numpy(L.make_logits(kind, T, B, C, 28, 1)).cuda()
kw = {}
if variant == "f64":
    x = x.double()
elif variant == "scorer":  # the bench's random label-bigram table
    kw["expansion_scores"] = torch.from_numpy(
        -np.abs(np.random.default_rng(5).standard_normal((C + 1, C))).astype(np.float32)).cuda()
sl = torch.full((B,), T, dtype=torch.int32).cuda()
for _ in range(reps):
    out = op.ctc_ext_beam_search_decoder_raw(x, sl, beam_width=W, top_paths=1, merge_repeated=True,
                                             blank_index=28, blank_label=-1, **kw)
torch.cuda.synchronize()
print("ok", int(out[1][0].numel()))
