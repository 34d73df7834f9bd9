"""Half-precision logits read directly by the kernels: beam-kernel time against float32 input and against
float32 input carrying the same (rounded) values -- quantised logits tie exactly far more often, which
is a property of the data, not of the loads.   python tools/half_input_check.py [cfg2|cfg4]"""
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

lib = _lib.load()
lib.ctcx_profile_enable(1)
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
T, B, C, W, blank, merge = (500, 256, 29, 100, 28, True) if cfg == "cfg2" else (400, 128, 1024, 16, 1023, False)
x = torch.from_numpy(L.make_logits("gauss", T, B, C, blank, 1)).cuda()
sl = torch.full((B,), T, dtype=torch.int32).cuda()
xb = x.bfloat16()
print(cfg)
for name, inp in (("f32 data, f32 kernel", x), ("bf16-rounded data, f32 kernel", xb.float()), ("bf16 data, bf16 kernel", xb),
                  ("f16-rounded data, f32 kernel", x.half().float()), ("f16 data, f16 kernel", x.half())):
    ms = []
    for i in range(5):
        op.ctc_ext_beam_search_decoder_raw(inp, sl, beam_width=W, top_paths=1, merge_repeated=merge, blank_index=blank)
        buf = (ctypes.c_float * 5)()
        lib.ctcx_profile_get(buf)
        ms.append((buf[0], buf[1]))
    m = np.mean(ms[1:], axis=0)
    print("%-32s pre-pass %.3f ms  beam %.3f ms" % (name, m[0], m[1]))
