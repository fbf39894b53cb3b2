"""spinn_b200 -- B200-native implementation of SPINN's training hot path.

Package contents (only what the path needs):
  csrc/      CUDA kernels + C ABI (include/spinn_b200.h), built to lib/libspinn_b200.so
  _cabi.py   ctypes loader (no fallback)
  ops.py     torch.autograd.Function over the C ABI + the derivative "tower"
  common.py spinn1d.py spinn2d.py   host-side mirror of the reference's plug-in surface
  problems/  problem bases and the BASELINE configs (callers of the path)
  distributed.py   point sharding + one NCCL all-reduce per step
"""
__version__ = "0.1.0"
