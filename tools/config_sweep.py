"""Device-resident timing of the BASELINE.json shapes: python tools/config_sweep.py [--f64]
(--f64: float64 logits through ctcx_decode_f64, the reference's T = double registration)"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import ctcx_testlib as L
import ctc_beam_search_op_b200 as op
from ctc_beam_search_op_b200 import _lib

F64 = "--f64" in sys.argv
lib = _lib.load()
lib.ctcx_profile_enable(1)
CFGS = [("cfg1", 50, 8, 29, 10, 3, False, 28), ("cfg2", 500, 256, 29, 100, 1, True, 28),
        ("cfg3", 1500, 64, 32, 64, 4, False, 31), ("cfg4", 400, 128, 1024, 16, 1, False, 1023)]
for kind in ("gauss", "peaky"):
    for name, T, B, C, W, P, merge, blank in CFGS:
        x = torch.from_numpy(L.make_logits(kind, T, B, C, blank, 3)).cuda()
        if F64:
            x = x.double()
        sl = torch.full((B,), T, dtype=torch.int32).cuda()
        ms = []
        for i in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            op.ctc_ext_beam_search_decoder_raw(x, sl, beam_width=W, top_paths=P, merge_repeated=merge, blank_index=blank)
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            buf = (ctypes.c_float * 5)()
            lib.ctcx_profile_get(buf)
            ms.append(list(buf) + [wall])
        m = np.array(ms[1:]).mean(axis=0)
        print(("f64 " if F64 else "") + "%s %-5s T=%4d B=%3d C=%4d W=%3d P=%d: lognorm %.3f beam %.3f trace %.3f | call %.3f ms -> %.2f M frames/s"
              % (kind, name, T, B, C, W, P, m[0], m[1], m[2], m[5], T * B / m[5] / 1e3))
