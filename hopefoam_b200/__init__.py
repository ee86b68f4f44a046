"""hopefoam_b200 - B200-native (sm_100a, FP64) explicit nodal-DG RHS + RK stage behind HopeFOAM's DG operator API.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + C ABI, built into libhopedg.so),
`include/` (the C++ facade mirroring dgm::/dgc::/dg::solveEquation), `capi.py` (ctypes binding used by the
tests and bench.py) and `meshgen.py` (synthetic benchmark meshes).
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
