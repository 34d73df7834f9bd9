"""SM clock while the beam kernel runs back to back: python tools/clock_probe.py [B] [seconds]"""
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import ctcx_testlib as L
import ctc_beam_search_op_b200 as op

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
x = torch.from_numpy(L.make_logits("gauss", 500, B, 29, 28, 1)).cuda()
sl = torch.full((B,), 500, dtype=torch.int32).cuda()
lines = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                         "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
th = threading.Thread(target=lambda: [lines.append(l.strip()) for l in proc.stdout], daemon=True)
th.start()
time.sleep(0.5)
n0 = len(lines)
t0 = time.time()
k = 0
while time.time() - t0 < secs:
    op.ctc_ext_beam_search_decoder_raw(x, sl, beam_width=100, top_paths=1, merge_repeated=True, blank_index=28)
    k += 1
torch.cuda.synchronize()
dt = time.time() - t0
time.sleep(0.2)
proc.terminate()
print("idle:", lines[:max(n0, 1)][-3:])
print("under load:", lines[n0 + 2:n0 + 12])
print("%d decodes in %.2f s = %.3f ms each (wall, B=%d)" % (k, dt, 1e3 * dt / k, B))
