"""Helper for tests/test_gpu_distributed.py: one process per GPU under torchrun, NCCL backend.
Every rank decodes its block of the batch (decode_distributed); rank 0 compares the merged result
with a single-GPU decode of the whole batch and with the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import ctcx_testlib as L
import ctc_beam_search_op_b200 as op


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    T, B, C, W, P = 60, 11, 29, 20, 2
    x = L.make_logits("peaky", T, B, C, 28, 5)
    sl = L.ragged_lengths(T, B, 5)
    got = op.decode_distributed(x, sl, beam_width=W, top_paths=P, merge_repeated=True, blank_index=28)
    ok = 1
    if rank == 0:
        whole = op.ctc_ext_beam_search_decoder_raw(x, sl, beam_width=W, top_paths=P, merge_repeated=True,
                                                   blank_index=28)
        want = L.pack_sparse(L.oracle_decode(x, sl, W, P, True, 28, -1))
        for g in range(6):
            for p in range(P):
                ok &= int(np.array_equal(np.asarray(got[g][p]), np.asarray(whole[g][p])))
                ok &= int(np.array_equal(np.asarray(got[g][p]), want[g][p]))
        ok &= int(np.array_equal(np.asarray(got[6]).view(np.uint32), want[6].view(np.uint32)))
    else:
        assert got is None
    # an error in any block is raised where the result is gathered
    bad = sl.copy()
    bad[B - 1] = T + 3
    try:
        op.decode_distributed(x, bad, beam_width=W, top_paths=P, blank_index=28)
        raised = False
    except op.FailedPreconditionError:
        raised = True
    if rank == 0 or rank == world - 1:
        ok &= int(raised)
    t = torch.tensor([ok], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_OK" if int(t.item()) == 1 else "DIST_FAIL", "world", world)
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
