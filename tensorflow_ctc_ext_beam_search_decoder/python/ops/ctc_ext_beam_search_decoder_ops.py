"""Same module path and symbol as the reference's
tensorflow_ctc_ext_beam_search_decoder/python/ops/ctc_ext_beam_search_decoder_ops.py:10-12, bound to
the CUDA library instead of a TensorFlow op library."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _root not in sys.path:
    sys.path.insert(0, _root)

from ctc_beam_search_op_b200 import ctc_ext_beam_search_decoder_raw as ctc_ext_beam_search_decoder  # noqa: E402,F401
