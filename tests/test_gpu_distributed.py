"""One process per GPU (torchrun, NCCL): batch blocks decoded per rank, host-side gather to rank 0, no
collective on the decode path. Uses two GPUs when the box has them, else one
(the NCCL plumbing and the gather are the same)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_decode_distributed_under_torchrun_nccl():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    n = min(2, torch.cuda.device_count())
    env = dict(os.environ)
    env.pop("CTCX_BEAM_IMPL", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tests", "dist_worker.py")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
