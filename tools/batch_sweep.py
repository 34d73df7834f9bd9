"""Beam-kernel time vs batch size (T=500, C=29, W=100): python tools/batch_sweep.py [kind] B1 B2 ..."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import ctcx_testlib as L
import ctc_beam_search_op_b200 as op
from ctc_beam_search_op_b200 import _lib

kind = sys.argv[1] if len(sys.argv) > 1 else "gauss"
Bs = [int(v) for v in sys.argv[2:]] or [148, 256, 296, 1024]
lib = _lib.load()
lib.ctcx_profile_enable(1)
T, C, W = 500, 29, 100
base = L.make_logits(kind, T, 256, C, 28, 1)
for B in Bs:
    reps = (B + 255) // 256
    x = torch.from_numpy(np.concatenate([base] * reps, axis=1)[:, :B].copy()).cuda()
    sl = torch.full((B,), T, dtype=torch.int32).cuda()
    ms = []
    for i in range(4):
        op.ctc_ext_beam_search_decoder_raw(x, sl, beam_width=W, top_paths=1, merge_repeated=True, blank_index=28)
        buf = (ctypes.c_float * 5)()
        lib.ctcx_profile_get(buf)
        ms.append(list(buf))
    m = np.array(ms[1:]).mean(axis=0)
    print("B=%5d beam %.3f ms trace %.3f ms lognorm %.3f total %.3f -> %.1f M frames/s (beam only %.1f M), %.2f us/frame/CTA-slot"
          % (B, m[1], m[2], m[0], m[4], B * T / m[4] / 1e3, B * T / m[1] / 1e3, m[1] * 1e3 / T))
