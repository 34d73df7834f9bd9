"""ctcx: B200-native (sm_100a) CTC "extended" beam-search decode -- a drop-in for
prouast/ctc-beam-search-op's `ctc_ext_beam_search_decoder` (CTC beam search that also returns the
best alignment of every returned path). See DESIGN.md / INTEGRATION.md at the repository root.

    from ctc_beam_search_op_b200 import ctc_ext_beam_search_decoder
    decoded, alignment, log_probability = ctc_ext_beam_search_decoder(
        inputs, sequence_length, beam_width=100, top_paths=1, merge_repeated=True,
        blank_index=28, blank_label=-1)
"""
from .decoder import (PendingDecode, CTCExtBeamSearchDecoder, CTCExtBeamSearchDecoderStream, CtcxError, DecodeResult,  # noqa: F401
                      FailedPreconditionError, FLAG_ROUNDING_ANOMALY, set_beam_impl,
                      InvalidArgumentError, SparseTensor, UnsupportedError,
                      ctc_ext_beam_search_decoder, ctc_ext_beam_search_decoder_raw,
                      decode_host_cabi)

from . import torch_op  # noqa: F401,E402  (registers torch.ops.ctcx.ctc_ext_beam_search_decoder)
from .sharding import bind_host_to_device, decode_distributed, decode_multi_device, merge_raw, shard_bounds  # noqa: F401,E402

__all__ = ["PendingDecode", "CTCExtBeamSearchDecoderStream", "DecodeResult", "FLAG_ROUNDING_ANOMALY", "set_beam_impl", "decode_multi_device", "decode_distributed", "bind_host_to_device", "shard_bounds", "merge_raw","ctc_ext_beam_search_decoder", "ctc_ext_beam_search_decoder_raw", "decode_host_cabi",
           "SparseTensor", "CTCExtBeamSearchDecoder", "CtcxError", "InvalidArgumentError",
           "FailedPreconditionError", "UnsupportedError"]
