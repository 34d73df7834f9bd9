"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on the same
seeded inputs. Labels, alignments and the float32 bits of log_probability must be IDENTICAL: the
device arithmetic reproduces the host libm bit for bit and the beam order is the oracle's stable
order (DESIGN.md "Numerics" / "Tie policy")."""
import numpy as np
import pytest

import ctcx_testlib as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["fast", "generic"])
def op(request):
    """Both beam kernels: the default dispatch (v2 fast path where it applies) and the generic
    kernel forced through CTCX_BEAM_IMPL=generic."""
    import os
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ctc_beam_search_op_b200 as m
    old = os.environ.get("CTCX_BEAM_IMPL")
    if request.param == "generic":
        os.environ["CTCX_BEAM_IMPL"] = "generic"
    else:
        os.environ.pop("CTCX_BEAM_IMPL", None)
    yield m
    if old is None:
        os.environ.pop("CTCX_BEAM_IMPL", None)
    else:
        os.environ["CTCX_BEAM_IMPL"] = old


def _dense_from_raw(raw, B, P, T):
    """Raw sparse outputs -> per (b,p) python lists."""
    dec = [[[] for _ in range(P)] for _ in range(B)]
    ali = [[[] for _ in range(P)] for _ in range(B)]
    for p in range(P):
        for (b, i), v in zip(np.asarray(raw[0][p]).tolist(), np.asarray(raw[1][p]).tolist()):
            assert i == len(dec[b][p])
            dec[b][p].append(v)
        for (b, i), v in zip(np.asarray(raw[3][p]).tolist(), np.asarray(raw[4][p]).tolist()):
            assert i == len(ali[b][p])
            ali[b][p].append(v)
    return dec, ali


def check_against_oracle(op, x, sl, W, P, merge, blank, blank_label, expect_all=True):
    T, B, C = x.shape
    want = L.oracle_decode(x, sl, W, P, merge, blank, blank_label)
    raw = op.ctc_ext_beam_search_decoder_raw(x, sl, beam_width=W, top_paths=P, merge_repeated=merge,
                                             blank_index=blank, blank_label=blank_label)
    packed = L.pack_sparse(want)
    bad = []
    dec, ali = _dense_from_raw(raw, B, P, T)
    for b in range(B):
        for p in range(P):
            if dec[b][p] != want.decoded(b, p) or ali[b][p] != want.alignment(b, p) or \
                    np.float32(raw[6][b, p]).view(np.uint32) != np.float32(want.logp[b, p]).view(np.uint32):
                bad.append((b, p))
    if expect_all:
        assert not bad, "mismatching (utterance, path): %s" % bad[:10]
        for g in range(6):
            for p in range(P):
                np.testing.assert_array_equal(np.asarray(raw[g][p]), packed[g][p])
    return bad


CASES = [
    # kind, T, B, C, W, P, merge, blank, blank_label, ragged
    ("gauss", 50, 8, 29, 10, 3, False, 28, -1, False),   # BASELINE cfg1
    ("peaky", 50, 8, 29, 10, 3, False, 28, -1, True),
    ("peaky", 60, 6, 29, 100, 1, True, 28, -1, False),   # cfg2 shape, short
    ("gauss", 60, 6, 29, 100, 1, True, 28, -1, True),
    ("gauss", 40, 4, 32, 64, 4, False, 31, -1, False),   # cfg3 shape, short
    ("peaky", 30, 3, 1024, 16, 1, False, 1023, -1, False),  # cfg4 shape, short (streaming mode)
    ("gauss", 30, 16, 6, 4, 4, True, 3, 1, True),
    ("gauss", 20, 16, 3, 2, 2, False, 1, 9, False),
    ("gauss", 20, 16, 5, 1, 1, True, 0, -1, False),
    ("gauss", 25, 4, 40, 200, 2, False, 7, -1, False),   # WMAX=256 tier
    ("gauss", 12, 2, 12, 300, 3, False, 0, -1, False),   # WMAX=1024 tier
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-T%d-B%d-C%d-W%d-P%d" % c[:6])
def test_parity_small(op, case):
    kind, T, B, C, W, P, merge, blank, blank_label, ragged = case
    x = L.make_logits(kind, T, B, C, blank, seed=11)
    sl = L.ragged_lengths(T, B, 11) if ragged else np.full(B, T, np.int32)
    check_against_oracle(op, x, sl, W, P, merge, blank, blank_label)


def test_constant_logits_pathological_ties(op):
    """All classes equally likely: every child of a row ties exactly; the cut through the tied
    scores must follow the stable order (members, then (row, label))."""
    for (T, B, C, W, P) in [(12, 2, 8, 10, 3), (10, 2, 29, 100, 2), (8, 1, 4, 3, 3)]:
        x = np.zeros((T, B, C), np.float32)
        check_against_oracle(op, x, np.full(B, T, np.int32), W, P, False, C - 1, -1)
    # two distinct values only
    rng = np.random.default_rng(3)
    x = rng.integers(0, 2, (15, 4, 12)).astype(np.float32)
    check_against_oracle(op, x, np.full(4, 15, np.int32), 20, 4, True, 0, -1)


def test_device_math_is_bit_exact(op):
    """ExpfExact / Log1pfExact / LogfExact on the device == the portable twin == host libm."""
    import ctypes
    import torch
    from ctc_beam_search_op_b200 import _lib
    lib = _lib.load()
    cpu = ctypes.CDLL(L.ORACLE_SO)
    rng = np.random.default_rng(5)
    fp = ctypes.POINTER(ctypes.c_float)
    doms = [(0, -np.abs(rng.standard_normal(200000) * 30).astype(np.float32), cpu.ctcx_libm_expf_v),
            (1, rng.random(200000).astype(np.float32), cpu.ctcx_libm_log1pf_v),
            (2, (1 + rng.random(200000) * 1100).astype(np.float32), cpu.ctcx_libm_logf_v)]
    for op_id, xs, fn in doms:
        xs = np.concatenate([xs, np.array([0.0, 1.0] if op_id else [0.0, -1e-30, -87.5, -103.0, -104.5, -np.inf], np.float32)])
        if op_id == 2:
            xs = np.maximum(xs, 1.0)
        want = np.empty_like(xs)
        fn(xs.ctypes.data_as(fp), want.ctypes.data_as(fp), len(xs))
        xd = torch.from_numpy(xs).cuda()
        yd = torch.empty_like(xd)
        assert lib.ctcx_debug_math_f32(op_id, xd.data_ptr(), yd.data_ptr(), len(xs), None) == 0
        torch.cuda.synchronize()
        got = yd.cpu().numpy()
        np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
