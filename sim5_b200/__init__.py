"""sim5_b200 -- B200-native (sm_100a, FP64) implementation of the SIM5 per-pixel photon hot path.

csrc/    hand-written CUDA kernels + the C-ABI (libsim5b200.so, built in-tree)
abi.py   ctypes mirror of include/sim5_b200.h, BASELINE config presets
api.py   numpy binding of the batched entry (sim5_trace_image) and of the element-wise entries
dist.py  row-block split / gather helpers for one-process-per-GPU runs (torch.distributed)
"""
from . import abi  # noqa: F401

__all__ = ["abi", "api"]
