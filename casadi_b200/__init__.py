"""casadi_b200: a B200-native (sm_100a) evaluator for CasADi's `Function::map(N, "cuda")`.

Package layout
  csrc/      CUDA kernels, tape compiler and the C ABI  -> lib/libcasadi_cuda.so
  host/      `CudaMap`, the C++ subclass of casadi::Map, and the reference-side patch
  capi.py    ctypes binding of the C ABI
  cuda_map.py  Python mirror of the Map interface (used by tests and bench.py)
"""
from .capi import CcuError, LAYOUT_AOS, LAYOUT_SOA  # noqa: F401
from .cuda_map import CudaMap, CudaMultiMap, CudaTape  # noqa: F401
from .linsol import CudaLinsol  # noqa: F401
from .tapeio import load_case, load_tape  # noqa: F401
