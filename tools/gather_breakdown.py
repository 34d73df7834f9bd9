"""Where the time of decode_distributed's gather goes (run under torchrun, NCCL):
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/gather_breakdown.py [B_total]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

import ctc_beam_search_op_b200 as op
from ctc_beam_search_op_b200 import sharding

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", local)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T, C, W = 500, 29, 100
g = torch.Generator(device=dev)
g.manual_seed(4)
x = torch.randn((T, B, C), generator=g, device=dev)
sl = torch.full((B,), T, dtype=torch.int32, device=dev)
b0, b1 = op.shard_bounds(B, world)[rank]
kw = dict(beam_width=W, top_paths=1, merge_repeated=True, blank_index=28, blank_label=-1)


def timed(f, n=8):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n * 1e3
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


raw = op.ctc_ext_beam_search_decoder_raw(x[:, b0:b1, :], sl[b0:b1], outputs="device", **kw)
n = raw.packed.numel()
recv = [torch.empty(n, dtype=torch.int64, device=dev) for _ in range(world)] if rank == 0 else None
hdr = torch.zeros(8, dtype=torch.int64, device=dev)
allh = torch.empty((world, 8), dtype=torch.int64, device=dev)
t_dec = timed(lambda: op.ctc_ext_beam_search_decoder_raw(x[:, b0:b1, :], sl[b0:b1], outputs="device", **kw))
t_hdr = timed(lambda: (dist.all_gather_into_tensor(allh, hdr), allh.cpu()))
t_gat = timed(lambda: dist.gather(raw.packed, recv, dst=0))
t_all = timed(lambda: op.decode_distributed(x, sl, dst=0, **kw))
if rank == 0:
    print("world %d, B=%d: shard decode %.3f ms | header all-gather + sync %.3f ms | gather of %.1f MB per rank %.3f ms | "
          "decode_distributed %.3f ms" % (world, B, t_dec, t_hdr, n * 8 / 1e6, t_gat, t_all))
dist.destroy_process_group()
