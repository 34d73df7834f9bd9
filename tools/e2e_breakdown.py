"""Where the end-to-end (host in, host out) time of the cfg2 workload goes: python tools/e2e_breakdown.py"""
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

T, B, C, W = 500, (int(sys.argv[1]) if len(sys.argv) > 1 else 256), 29, 100
NB = 10 if B <= 1024 else 2
kw = dict(beam_width=W, top_paths=1, merge_repeated=True, blank_index=28, blank_label=-1)
xs = [torch.from_numpy(L.make_logits("gauss", T, B, C, 28, s)).pin_memory() for s in range(NB)]
sl = torch.full((B,), T, dtype=torch.int32).pin_memory()
dev = torch.device("cuda", 0)


def timed(f, n=(30 if B <= 1024 else 6)):
    for i in range(3):
        f(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        f(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


xd = [x.to(dev) for x in xs]
sd = sl.to(dev)
print("h2d logits only      %.3f ms" % timed(lambda i: xs[i % NB].to(dev, non_blocking=True)))
print("decode, device in/out %.3f ms" % timed(lambda i: op.ctc_ext_beam_search_decoder_raw(xd[i % NB], sd, **kw)))
out = op.ctc_ext_beam_search_decoder_raw(xd[0], sd, **kw)
nbytes = sum(t.numel() * t.element_size() for g in out[:6] for t in g) + out[6].numel() * 4
print("d2h outputs only (%.2f MB, .cpu() per tensor) %.3f ms" % (nbytes / 1e6, timed(lambda i: [[t.cpu() for t in g] for g in out[:6]] + [out[6].cpu()])))
print("e2e (host in, host out) %.3f ms" % timed(lambda i: op.ctc_ext_beam_search_decoder_raw(xs[i % NB], sl, **kw)))

# the C-ABI host-input entry alone (no pack, no D2H of the outputs): overlapped copy vs copy-then-decode
import ctypes
from ctc_beam_search_op_b200 import _lib
lib = _lib.load()
ws_bytes = lib.ctcx_workspace_bytes(T, B, C, W, 1)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
staging = torch.empty(T * B * C * 4, dtype=torch.uint8, device=dev)
side = torch.cuda.Stream()
arr = ctypes.c_int64 * 1
n_dec, max_dec, n_ali, max_ali = arr(), arr(), arr(), arr()
sizes = _lib.CtcxSizes(n_dec, max_dec, n_ali, max_ali)
flags = ctypes.c_int32(0)
stream = torch.cuda.current_stream().cuda_stream


def hostin(i):
    rc = lib.ctcx_decode_hostin(xs[i % NB].data_ptr(), 0, 0, T, B, C, sl.data_ptr(), W, 1, 1, 28, -1,
                                staging.data_ptr(), staging.numel(), ws.data_ptr(), ws_bytes, stream, side.cuda_stream,
                                ctypes.byref(sizes), ctypes.byref(flags))
    assert rc == 0, rc


def copy_then_decode(i):
    staging.view(torch.float32).view(T, B, C).copy_(xs[i % NB], non_blocking=True)
    rc = lib.ctcx_decode_f32(staging.data_ptr(), T, B, C, sd.data_ptr(), W, 1, 1, 28, -1, ws.data_ptr(), ws_bytes, stream,
                             ctypes.byref(sizes), ctypes.byref(flags))
    assert rc == 0, rc


print("ctcx_decode_hostin (pinned logits, overlapped copy) %.3f ms" % timed(hostin))
print("copy, then ctcx_decode_f32                           %.3f ms" % timed(copy_then_decode))
xp = [x.clone() for x in xs[:min(3, NB)]]  # pageable
print("ctcx_decode_hostin (pageable logits: copy first)     %.3f ms" % timed(
    lambda i: lib.ctcx_decode_hostin(xp[i % len(xp)].data_ptr(), 0, 0, T, B, C, sl.data_ptr(), W, 1, 1, 28, -1, staging.data_ptr(),
                                     staging.numel(), ws.data_ptr(), ws_bytes, stream, side.cuda_stream, ctypes.byref(sizes),
                                     ctypes.byref(flags))))
