"""ctypes driver of oracle/ref_cpu.c (CPU baseline in reference-faithful mode).  TEST INFRASTRUCTURE ONLY:
imported by tests/ and by bench.py's cpu_baseline / --impl reference legs, never by the product."""
from __future__ import annotations

import ctypes as C
import subprocess
import time
from pathlib import Path

import numpy as np

from . import dg_oracle as o

_HERE = Path(__file__).resolve().parent
_LIB = None


class _RefCase(C.Structure):
    _fields_ = [("K", C.c_int), ("F", C.c_int), ("Np", C.c_int), ("Ng", C.c_int), ("Nfp", C.c_int), ("Nfg", C.c_int)] + \
               [(n, C.c_void_p) for n in ("Vg", "If", "D1x", "D1y", "WJ", "M", "Lchol", "fnx", "fny", "fWJ", "mapO", "mapN", "cellFace",
                                          "faceOwner", "faceLocO", "faceLocN", "f2c", "faceRot")]


def lib():
    global _LIB
    if _LIB is None:
        so = _HERE / "libref_cpu.so"
        if not so.exists():
            subprocess.run(["make", "-C", str(_HERE)], check=True)
        _LIB = C.CDLL(str(so))
        _LIB.refcpu_work_doubles.restype = C.c_size_t
        _LIB.refcpu_work_doubles.argtypes = [C.c_int] * 4
        _LIB.refcpu_euler_steps.restype = C.c_int
        _LIB.refcpu_euler_steps.argtypes = [C.POINTER(_RefCase), C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p]
        _LIB.refcpu_max_threads.restype = C.c_int
    return _LIB


class RefCpuCase:
    """Packs an oracle Case (periodic mesh: no patches) into the C struct."""

    def __init__(self, case: o.Case):
        assert len(case.mesh.patches) == 0, "the C port handles the periodic (all-interior) benchmark mesh"
        self.case = case
        ref, geo, m = case.ref, case.geo, case.mesh
        keep = {}
        def f64(a):
            a = np.ascontiguousarray(a, dtype=np.float64); keep[id(a)] = a; return a.ctypes.data
        def i32(a):
            a = np.ascontiguousarray(a, dtype=np.int32); keep[id(a)] = a; return a.ctypes.data
        L = np.linalg.cholesky(geo.M)
        self.c = _RefCase(m.K, m.F, ref.Np, ref.Ng, ref.Nfp, ref.Nfg, f64(ref.Vg), f64(ref.If), f64(geo.D1x), f64(geo.D1y), f64(geo.WJ),
                          f64(geo.M), f64(L), f64(geo.fnx[..., 0]), f64(geo.fnx[..., 1]), f64(geo.fWJ), i32(case.map_o), i32(case.map_n),
                          i32(m.cell_face), i32(m.face_owner), i32(m.face_loc_o), i32(m.face_loc_n), i32(ref.f2c), i32(m.face_rot))
        self._keep = keep
        self.work = np.empty(lib().refcpu_work_doubles(m.K, m.F, ref.Ng, ref.Nfg))
        self.tmp = np.empty(2 * m.K * ref.Np * 4)

    def steps(self, rho, rhoU, E, gamma, dt, nsteps, threads=0):
        rho = np.ascontiguousarray(rho, dtype=np.float64).copy()
        rhoU = np.ascontiguousarray(rhoU, dtype=np.float64).copy()
        E = np.ascontiguousarray(E, dtype=np.float64).copy()
        t0 = time.perf_counter()
        lib().refcpu_euler_steps(C.byref(self.c), rho.ctypes.data, rhoU.ctypes.data, E.ctypes.data, gamma, dt, nsteps, threads,
                                 self.work.ctypes.data, self.tmp.ctypes.data)
        return rho, rhoU, E, time.perf_counter() - t0


_CACHE = {}


def time_euler_steps(N=4, n=160, steps=4, threads=0, dt=1.28e-4, gamma=1.4):
    """Bounded CPU sample of the benchmark workload: n x n x 2 periodic jittered triangles, `steps` SSP-RK2 steps."""
    from hopefoam_b200 import meshgen          # mesh generator only (numpy); no product compute is involved
    key = (N, n)
    if key not in _CACHE:
        mg = meshgen.jittered_square(n, periodic=True)
        mesh = o.build_connectivity(mg["xy"], mg["tris"], [], [], point_equiv=mg["point_equiv"])
        _CACHE[key] = RefCpuCase(o.Case(mesh, N))
    rc = _CACHE[key]
    x, y = rc.case.geo.x[..., 0], rc.case.geo.x[..., 1]
    rho, ru, rv, E = o.vortex_exact(x, y, 0.0, gamma)
    nthreads = threads or lib().refcpu_max_threads()
    _, _, _, sec = rc.steps(rho, np.stack([ru, rv], -1), E, gamma, dt, steps, nthreads)
    K, Np = rc.case.mesh.K, rc.case.ref.Np
    return {"K": K, "threads": nthreads, "seconds": sec, "dof_updates_per_s": 2 * steps * 4 * Np * K / sec}
