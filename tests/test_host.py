"""CPU suite, part 2: host logic and the C-ABI surface. No compute calls (no GPU here): the library
must load, export every symbol include/ctcx.h declares, map return codes to the reference's
messages, and the Python host must validate exactly like the reference op (kernels.cc:97-160).
"""
import ctypes
import os
import re

import numpy as np
import pytest

import ctcx_testlib as L

ROOT = L.ROOT


@pytest.fixture(scope="module")
def lib():
    from ctc_beam_search_op_b200 import _lib
    return _lib.load()


def test_cabi_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "ctcx.h")).read()
    declared = set(re.findall(r"\b(ctcx_[a-z0-9_]+)\s*\(", hdr))
    assert {"ctcx_decode_f32", "ctcx_pack_f32", "ctcx_decode_host_f32", "ctcx_workspace_bytes"} <= declared
    for sym in sorted(declared):
        assert hasattr(lib, sym), "libctcx.so does not export %s" % sym
    from ctc_beam_search_op_b200 import _lib
    assert declared == set(_lib.EXPORTS)


def test_error_messages_are_the_references(lib):
    want = {1: "inputs is not a 3-Tensor", 2: "max_time is 0", 3: "sequence_length is not a vector",
            6: "requested more paths than the beam width.",
            7: "Less leaves in the beam search than requested."}
    for code, msg in want.items():
        assert lib.ctcx_strerror(code).decode() == msg
    assert lib.ctcx_strerror(4).decode().startswith("len(sequence_length) != batch_size.")


def test_workspace_and_limits(lib):
    from ctc_beam_search_op_b200 import _lib
    lim = _lib.CtcxLimits()
    assert lib.ctcx_get_limits(ctypes.byref(lim)) == 100  # built for sm_100
    assert lim.max_beam_width >= 256 and lim.max_classes >= 1024
    small = lib.ctcx_workspace_bytes(50, 8, 29, 10, 3)
    big = lib.ctcx_workspace_bytes(500, 256, 29, 100, 1)
    assert 0 < small < big
    assert big >= 500 * 256 * 100 * 4  # back-pointer records (4 B per slot and frame here) dominate
    assert lib.ctcx_workspace_bytes(0, 8, 29, 10, 3) == 0


def test_cabi_validation_without_device(lib):
    """Argument errors are reported before any CUDA call."""
    from ctc_beam_search_op_b200 import _lib
    sizes = _lib.CtcxSizes()
    args = lambda T, B, C, W, P, blank: (None, T, B, C, None, W, P, 0, blank, -1, None, 0, None,  # noqa: E731
                                         ctypes.byref(sizes), None)
    assert lib.ctcx_decode_f32(*args(0, 1, 3, 2, 1, 0)) == 2     # max_time is 0
    assert lib.ctcx_decode_f32(*args(5, 1, 3, 0, 1, 0)) == 8     # beam_width < 1
    assert lib.ctcx_decode_f32(*args(5, 1, 3, 2, 1, 3)) == 8     # blank_index out of range
    assert lib.ctcx_decode_f32(*args(5, 1, 3, 5000, 1, 0)) == 9  # beyond this build's limits
    assert lib.ctcx_decode_f32(*args(5, 1, 3, 2, 1, 0)) == 10    # no workspace


def test_python_host_validation_order():
    import ctc_beam_search_op_b200 as op
    x = np.zeros((4, 2, 3), np.float32)
    with pytest.raises(ValueError, match="beam_width"):
        op.ctc_ext_beam_search_decoder(x, [4, 4], beam_width=0, top_paths=1)
    with pytest.raises(ValueError, match="top_paths"):
        op.ctc_ext_beam_search_decoder(x, [4, 4], beam_width=1, top_paths=0)
    with pytest.raises(op.InvalidArgumentError, match="inputs is not a 3-Tensor"):
        op.ctc_ext_beam_search_decoder(x[0], [4, 4], beam_width=2, top_paths=1)
    with pytest.raises(op.InvalidArgumentError, match="max_time is 0"):
        op.ctc_ext_beam_search_decoder(x[:0], [4, 4], beam_width=2, top_paths=1)
    with pytest.raises(op.InvalidArgumentError, match="sequence_length is not a vector"):
        op.ctc_ext_beam_search_decoder(x, [[4, 4]], beam_width=2, top_paths=1)
    with pytest.raises(op.FailedPreconditionError, match=r"len\(sequence_length\) != batch_size"):
        op.ctc_ext_beam_search_decoder(x, [4], beam_width=2, top_paths=1)
    with pytest.raises(TypeError):
        op.ctc_ext_beam_search_decoder(x.astype(np.int32), [4, 4], beam_width=2, top_paths=1)


def test_no_cpu_fallback():
    """Without a CUDA device the op must raise, not compute on the CPU."""
    import torch
    import ctc_beam_search_op_b200 as op
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        op.ctc_ext_beam_search_decoder(np.zeros((4, 2, 3), np.float32), [4, 4], beam_width=2, top_paths=1)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing in the product package may import, include, link
    or dlopen anything under oracle/ (a product path that routes through it would void parity)."""
    pkg = os.path.join(ROOT, "ctc-beam-search-op_b200")
    loaders = ("import", "#include", "CDLL", "dlopen", "load_library", "subprocess", "open(")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            for ln in open(os.path.join(dirpath, f)).read().splitlines():
                if "oracle" in ln.lower():
                    assert not any(k in ln for k in loaders), "%s: %s" % (f, ln.strip())


def test_shard_bounds_and_merge():
    import ctc_beam_search_op_b200 as op
    assert op.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert op.shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    x = L.make_logits("peaky", 20, 7, 6, 5, 3)
    sl = L.ragged_lengths(20, 7, 3)
    whole = L.pack_sparse(L.oracle_decode(x, sl, 4, 2, True, 5, -1))
    bounds = op.shard_bounds(7, 3)
    shards = [L.pack_sparse(L.oracle_decode(x[:, b0:b1], sl[b0:b1], 4, 2, True, 5, -1)) for b0, b1 in bounds]
    merged = op.merge_raw(shards, bounds, 7)
    for g in range(6):
        for p in range(2):
            np.testing.assert_array_equal(merged[g][p], whole[g][p])
    np.testing.assert_array_equal(merged[6], whole[6])


def test_torch_library_registration_traces_with_fake_tensors():
    """torch.ops.ctcx.ctc_ext_beam_search_decoder exists, has the reference op's argument order
    (ops.cc:10-16) and infers its output shapes without a device (ops.cc:41-61: data-dependent
    numbers of sparse entries, [batch, top_paths] log-probabilities)."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    from torch.fx.experimental.symbolic_shapes import ShapeEnv
    import ctc_beam_search_op_b200 as m  # noqa: F401
    op = torch.ops.ctcx.ctc_ext_beam_search_decoder
    schema = str(op.default._schema)
    for a in ("Tensor inputs", "Tensor sequence_length", "Int beam_width", "Int top_paths",
              "bool merge_repeated=False", "Int blank_index=0", "Int blank_label=-1"):
        assert a in schema, schema
    with FakeTensorMode(shape_env=ShapeEnv()):
        for dt in (torch.float32, torch.float64):
            x = torch.empty((50, 8, 29), device="cuda", dtype=dt)
            sl = torch.empty((8,), dtype=torch.int32, device="cuda")
            out = op(x, sl, 10, 3, True, 28, -1)
            assert len(out) == 19
            assert out[0].shape[1] == 2 and out[0].dtype == torch.int64 and out[6].shape == (2,)
            assert out[18].shape == (8, 3) and out[18].dtype == dt
    # the flat list maps back onto the raw namedtuple
    r = m.torch_op.unflatten(list(range(19)), 3)
    assert r.decoded_indices == [0, 1, 2] and r.alignment_shape == [15, 16, 17] and r.log_probability == 18


def test_error_classes_survive_pickling():
    """decode_distributed ships a rank's exception to the gathering rank: code and text must survive."""
    import pickle
    import ctc_beam_search_op_b200 as m
    for cls, code in ((m.InvalidArgumentError, 6), (m.FailedPreconditionError, 5), (m.UnsupportedError, 9)):
        e = pickle.loads(pickle.dumps(cls(code, "sequence_length(3) <= 8")))
        assert type(e) is cls and e.code == code and str(e) == "sequence_length(3) <= 8"


def test_packed_output_buffer_layout():
    """All 6*P sparse tensors + log_probability are views of ONE int64 buffer (one D2H copy for host
    callers): the carve-up must be gap-free, non-overlapping and typed as ops.cc:17-23 says."""
    import torch
    from ctc_beam_search_op_b200 import decoder as D
    for f64 in (False, True):
        B, P = 5, 3
        counts = ([7, 0, 12], [40, 40, 33])
        n = D._pack_elems(B, P, counts, f64)
        buf = torch.arange(n, dtype=torch.int64)
        groups, lp = D._carve(buf, B, P, counts, f64)
        seen = torch.zeros(n, dtype=torch.int32)
        for g in range(6):
            for p in range(P):
                t = groups[g][p]
                nd = counts[0][p] if g < 3 else counts[1][p]
                want = {0: (nd, 2), 1: (nd,), 2: (2,)}[g % 3]
                assert tuple(t.shape) == want and t.dtype == torch.int64 and t.is_contiguous()
                if t.numel():
                    seen[t.reshape(-1)] += 1  # values are the buffer offsets themselves
        assert lp.shape == (B, P) and lp.dtype == (torch.float64 if f64 else torch.float32)
        lp_words = B * P if f64 else (B * P + 1) // 2
        assert int(seen.sum()) == n - lp_words and int(seen.max()) == 1
        assert lp.data_ptr() == buf.data_ptr() + 8 * (n - lp_words)
