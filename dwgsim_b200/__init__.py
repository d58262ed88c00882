"""dwgsim_b200: the dwgsim_core read-pair loop (nh13/DWGSIM src/dwgsim.c:636-1099) as sm_100a CUDA kernels.

The product is libdwgsim_b200.so and its C ABI (include/dwgsim_gpu.h).  This package is the thin Python
host mirror used by tests and bench.py: it marshals arguments over ctypes and never computes reads itself.
"""
from .api import DwgsimGpu, DwgsimGpuError, params_from_options  # noqa: F401
