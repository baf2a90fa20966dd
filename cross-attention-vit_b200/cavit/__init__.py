"""cavit — B200-native (sm_100a) cross-attention ViT hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all device work is done
by hand-written CUDA kernels in libcavit_sm100a.so reached through a C ABI (include/cavit.h).
"""
from . import _abi  # noqa: F401
from ._abi import CavitError  # noqa: F401

__all__ = ["CavitError"]
