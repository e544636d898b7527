"""gnn_matlang_b200 -- B200-native GNNML3 hot path (SpectConv / ML3Layer / SpectralDesign) behind the
reference's own module API.  Host code is Python/PyTorch plumbing; all arithmetic runs in hand-written
sm_100a CUDA kernels inside libgnnml3_b200.so (C ABI, include/gnnml3_b200.h).  No CPU fallback."""
from . import _lib  # noqa: F401
from .libs.spect_conv import SpectConv, ML3Layer  # noqa: F401

__all__ = ["SpectConv", "ML3Layer"]
