"""vofod_b200 — B200-native implementation of VoFOD's per-scan volumetric hot path.

The product is the C-ABI library libvofod_cuda.so (include/vofod_cuda.h; sources in vofod_b200/csrc).
`capi` is the ctypes view of it used by the tests and bench.py; `synth` generates the synthetic scans.
"""
from . import abi  # noqa: F401
