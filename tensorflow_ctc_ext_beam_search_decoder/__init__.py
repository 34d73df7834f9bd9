"""Compatibility import path of the reference package (its __init__.py:19 re-exports the op):

    from tensorflow_ctc_ext_beam_search_decoder import ctc_ext_beam_search_decoder

resolves to the B200 implementation. As in the reference, the symbol is the RAW op wrapper: keyword
call, result indexable [0..6] = decoded_indices, decoded_values, decoded_shape, alignment_indices,
alignment_values, alignment_shape, log_probability (python/ops/ctc_ext_beam_search_decoder_ops.py:12).
"""
from tensorflow_ctc_ext_beam_search_decoder.python.ops.ctc_ext_beam_search_decoder_ops import (  # noqa: F401
    ctc_ext_beam_search_decoder)
