"""B200-native contrastive hot path of JingwWu/advise-video-ssl.

Python surface mirroring the reference's `models/contrastive.py`,
`models/losses.py` (ContrastiveLoss) and `utils/distributed.py` helpers, on top of
hand-written sm_100a CUDA kernels reached through the C-ABI in
`include/avssl_b200.h` (`libavssl_b200.so`).  No CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the shared library is missing)

__all__ = ["_lib"]
