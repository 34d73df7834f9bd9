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
    # device tensors in -> the packed outputs travel GPU to GPU and the merged result stays on rank 0's GPU;
    # here every rank passes only ITS block (the usual layout of a multi-process job)
    b0, b1 = op.shard_bounds(B, world)[rank]
    xd = torch.from_numpy(x[:, b0:b1]).cuda().contiguous()
    sd = torch.from_numpy(sl[b0:b1]).cuda()
    got2 = op.decode_distributed(xd, sd, beam_width=W, top_paths=P, merge_repeated=True, blank_index=28,
                                 global_batch=B)
    if rank == 0:
        assert got2[6].is_cuda and got2[0][0].is_cuda
        for g in range(6):
            for p in range(P):
                ok &= int(np.array_equal(got2[g][p].cpu().numpy(), want[g][p]))
        ok &= int(np.array_equal(got2[6].cpu().numpy().view(np.uint32), want[6].view(np.uint32)))
    else:
        assert got2 is None
    # ... and the whole tensor on every GPU, each rank decoding a strided view of it in place
    got3 = op.decode_distributed(torch.from_numpy(x).cuda(), torch.from_numpy(sl).cuda(), beam_width=W, top_paths=P,
                                 merge_repeated=True, blank_index=28)
    if rank == 0:
        for g in range(6):
            for p in range(P):
                ok &= int(np.array_equal(got3[g][p].cpu().numpy(), want[g][p]))
    # an error in any block is raised where the result is gathered
    bad = sl.copy()
    bad[B - 1] = T + 3
    try:
        op.decode_distributed(x, bad, beam_width=W, top_paths=P, blank_index=28)
        raised = False
    except op.FailedPreconditionError as e:
        raised = ("sequence_length(%d) <= %d" % (B - 1, T)) in str(e)  # the index in the WHOLE batch
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
