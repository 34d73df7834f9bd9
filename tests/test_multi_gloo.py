"""CPU suite, part 3: the N>1 path. Two processes over gloo shard the batch, each decodes its block
(the CPU oracle stands in for the GPU decode here -- this test covers the HOST logic: sharding,
gather, index shifting, error propagation) and rank 0 must hold exactly the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import ctcx_testlib as L


def _oracle_raw(x, sl, W, P, merge, blank, bl):
    return L.pack_sparse(L.oracle_decode(x, sl, W, P, merge, blank, bl))


def _worker(rank, world, port, case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctc_beam_search_op_b200 as op
        x, sl, W, P, merge, blank, bl = case[:7]
        try:
            if len(case) > 7 and case[7] == "local":  # every rank holds only its block of the batch
                b0, b1 = op.shard_bounds(x.shape[1], world)[rank]
                out = op.decode_distributed(np.ascontiguousarray(x[:, b0:b1]), sl[b0:b1], W, P, merge, blank, bl, dst=0,
                                            decode_fn=_oracle_raw, global_batch=x.shape[1])
            else:
                out = op.decode_distributed(x, sl, W, P, merge, blank, bl, dst=0, decode_fn=_oracle_raw)
            if rank == 0:
                q.put(("ok", [[np.asarray(t) for t in g] for g in out[:6]] + [np.asarray(out[6])]))
        except Exception as e:
            if rank == 0:
                q.put(("err", str(e)))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(case, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_two_ranks_equal_single_process():
    x = L.make_logits("peaky", 30, 5, 8, 7, 11)
    sl = L.ragged_lengths(30, 5, 11)
    status, got = _run((x, sl, 6, 3, True, 7, -1))
    assert status == "ok"
    want = _oracle_raw(x, sl, 6, 3, True, 7, -1)
    for g in range(6):
        for p in range(3):
            np.testing.assert_array_equal(got[g][p], want[g][p])
    np.testing.assert_array_equal(got[6], want[6])


def test_two_ranks_each_holding_its_own_block():
    x = L.make_logits("gauss", 20, 7, 6, 5, 12)
    sl = L.ragged_lengths(20, 7, 12)
    status, got = _run((x, sl, 5, 2, False, 5, -1, "local"))
    assert status == "ok"
    want = _oracle_raw(x, sl, 5, 2, False, 5, -1)
    for g in range(6):
        for p in range(2):
            np.testing.assert_array_equal(got[g][p], want[g][p])
    np.testing.assert_array_equal(got[6], want[6])


def test_two_ranks_error_propagates():
    x = L.make_logits("gauss", 1, 4, 3, 0, 2)
    sl = np.ones(4, np.int32)
    status, msg = _run((x, sl, 8, 8, False, 0, -1))  # only 3 leaves after one frame
    assert status == "err" and "Less leaves" in msg
