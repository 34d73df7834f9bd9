"""Streaming decode, small chunks: eager `step()` calls vs one captured CUDA graph replayed per chunk.
   python tools/stream_graph_bench.py [B] [chunk_frames]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import ctcx_testlib as L
import ctc_beam_search_op_b200 as op

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
Tc = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T, C, W = 2000, 29, 100
x = torch.from_numpy(L.make_logits("peaky", T, B, C, 28, 3)).cuda()
dec = op.CTCExtBeamSearchDecoderStream(batch_size=B, num_classes=C, beam_width=W, top_paths=1, max_time=T,
                                       merge_repeated=True, blank_index=28)
xs = torch.zeros((Tc, B, C), dtype=torch.float32, device="cuda")
ls = torch.full((B,), Tc, dtype=torch.int32, device="cuda")


def run(fn):
    dec.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(0, T, Tc):
        xs.copy_(x[t:t + Tc])
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / (T // Tc) * 1e6


run(lambda: dec.step_device(xs, ls))
eager_py = run(lambda: dec.step(xs))
eager = run(lambda: dec.step_device(xs, ls))
dec.reset()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    dec.step_device(xs, ls)
graph = run(g.replay)
print("B=%d chunk=%d frames: step() %.1f us/chunk, step_device() %.1f us/chunk, CUDA graph replay %.1f us/chunk "
      "(kernel floor ~%.1f us = %d frames x 4.6 us)" % (B, Tc, eager_py, eager, graph, Tc * 4.6, Tc))
