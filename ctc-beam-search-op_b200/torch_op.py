"""`torch.library` registration of the raw op, so that it can be called as
`torch.ops.ctcx.ctc_ext_beam_search_decoder(...)` from PyTorch programs, traced with fake tensors
and kept as one opaque node by graph capture -- the PyTorch counterpart of the reference's
`REGISTER_OP("CTCExtBeamSearchDecoder")` (cc/ops/ctc_ext_beam_search_decoder_ops.cc:9-24).

The op returns the reference's seven output groups flattened into one list of 6*top_paths + 1
tensors, group-major: decoded_indices[0..P), decoded_values[0..P), decoded_shape[0..P),
alignment_indices[0..P), alignment_values[0..P), alignment_shape[0..P), log_probability.
`unflatten(outputs, top_paths)` rebuilds the namedtuple of `ctc_ext_beam_search_decoder_raw`.
"""
from typing import List

import torch

from . import decoder as _decoder

_OP_NAME = "ctcx::ctc_ext_beam_search_decoder"


@torch.library.custom_op(_OP_NAME, mutates_args=(), device_types="cuda")
def _ctc_ext_beam_search_decoder(inputs: torch.Tensor, sequence_length: torch.Tensor, beam_width: int,
                                 top_paths: int, merge_repeated: bool = False, blank_index: int = 0,
                                 blank_label: int = -1) -> List[torch.Tensor]:
    raw = _decoder.ctc_ext_beam_search_decoder_raw(inputs, sequence_length, beam_width=beam_width,
                                                   top_paths=top_paths, merge_repeated=merge_repeated,
                                                   blank_index=blank_index, blank_label=blank_label)
    # the raw entry returns views of one packed buffer; a registered op's outputs must not alias
    flat = []
    for g in range(6):
        flat.extend(t.clone() for t in raw[g])
    flat.append(raw[6].clone())
    return flat


@_ctc_ext_beam_search_decoder.register_fake
def _(inputs, sequence_length, beam_width, top_paths, merge_repeated=False, blank_index=0, blank_label=-1):
    # shapes as cc/ops/ctc_ext_beam_search_decoder_ops.cc:41-61 infers them: the numbers of sparse
    # entries are data dependent
    ctx = torch.library.get_ctx()
    batch = inputs.shape[1]
    i64 = dict(dtype=torch.int64, device=inputs.device)
    n_dec = [ctx.new_dynamic_size() for _ in range(top_paths)]
    n_ali = [ctx.new_dynamic_size() for _ in range(top_paths)]
    out = [torch.empty((n, 2), **i64) for n in n_dec]
    out += [torch.empty((n,), **i64) for n in n_dec]
    out += [torch.empty((2,), **i64) for _ in range(top_paths)]
    out += [torch.empty((n, 2), **i64) for n in n_ali]
    out += [torch.empty((n,), **i64) for n in n_ali]
    out += [torch.empty((2,), **i64) for _ in range(top_paths)]
    lp_dtype = torch.float64 if inputs.dtype == torch.float64 else torch.float32
    out.append(torch.empty((batch, top_paths), dtype=lp_dtype, device=inputs.device))
    return out


def unflatten(outputs, top_paths):
    """The op's flat output list -> the 7-field namedtuple of ctc_ext_beam_search_decoder_raw."""
    P = int(top_paths)
    groups = [list(outputs[g * P:(g + 1) * P]) for g in range(6)]
    return _decoder.CTCExtBeamSearchDecoder(*groups, outputs[6 * P])


ctc_ext_beam_search_decoder_op = _ctc_ext_beam_search_decoder
