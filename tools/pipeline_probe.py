"""Per-step wall times of blocking vs pipelined (wait=False, one decode in flight) calls at the bench
workload, device-resident and page-locked host inputs:   python tools/pipeline_probe.py [steps]"""
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

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T, B, C, W = 500, 256, 29, 100
kw = dict(beam_width=W, top_paths=1, merge_repeated=True, blank_index=28)
base = L.make_logits("gauss", T, B, C, 28, 1)
rng = np.random.default_rng(3)
host = [torch.from_numpy(np.ascontiguousarray(base[:, rng.permutation(B)])).pin_memory() for _ in range(10)]
dev = [h.cuda() for h in host]
sl_h = torch.full((B,), T, dtype=torch.int32)
sl_d = sl_h.cuda()


def run(batches, seq, pipelined):
    for i in range(4):
        op.ctc_ext_beam_search_decoder_raw(batches[i], seq, **kw)
    torch.cuda.synchronize()
    marks = [time.perf_counter()]
    pend = None
    for i in range(steps):
        if pipelined:
            nxt = op.ctc_ext_beam_search_decoder_raw(batches[i % 10], seq, wait=False, **kw)
            if pend is not None:
                pend.result()
            pend = nxt
        else:
            op.ctc_ext_beam_search_decoder_raw(batches[i % 10], seq, **kw)
        marks.append(time.perf_counter())
    if pend is not None:
        pend.result()
    marks.append(time.perf_counter())
    d = np.diff(marks) * 1e3
    return d


for name, batches, seq in (("device", dev, sl_d), ("host", host, sl_h)):
    for pipelined in (False, True):
        d = run(batches, seq, pipelined)
        print("%-6s %-9s total %.3f ms/step | %s" % (name, "pipelined" if pipelined else "blocking", d.sum() / steps,
                                                     " ".join("%.2f" % v for v in d)))
