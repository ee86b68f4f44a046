"""ctypes binding of libhopedg.so (include/hopedg.h) - used by tests/, bench.py and __graft_entry__.py.

This is plumbing only: every numerical call goes through the C ABI into the CUDA kernels.  There is no
Python/NumPy fallback; if the shared library or a CUDA device is missing the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_LIB = None
LIB_PATH = Path(__file__).resolve().parent / "libhopedg.so"

BC_FIXED_VALUE, BC_ZERO_GRADIENT, BC_REFLECTIVE, BC_PROCESSOR, BC_EMPTY = 0, 1, 2, 3, 4
FLUX_ROE, FLUX_LF, FLUX_AVERAGE, FLUX_NONE = 0, 1, 2, 3

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); the list is checked against include/hopedg.h by tests/test_capi_exports.py
SIGNATURES = {
    "hdg_create": (C.c_int, [C.c_int, _vpp]),
    "hdg_destroy": (None, [C.c_void_p]),
    "hdg_last_error": (C.c_char_p, [C.c_void_p]),
    "hdg_version": (C.c_char_p, []),
    "hdg_sync": (C.c_int, [C.c_void_p]),
    "hdg_set_order": (C.c_int, [C.c_void_p, C.c_int]),
    "hdg_get_sizes": (C.c_int, [C.c_void_p, _i32p, _i32p, _i32p, _i32p]),
    "hdg_get_operator": (C.c_int64, [C.c_void_p, C.c_char_p, _f64p, C.c_int64]),
    "hdg_get_face_to_cell_index": (C.c_int, [C.c_void_p, _i32p]),
    "hdg_set_mesh_triangles": (C.c_int, [C.c_void_p, C.c_int64, _f64p, C.c_int64, _i32p, _i32p, C.c_int32, _i32p, _i32p, _i32p]),
    "hdg_set_mesh_polymesh": (C.c_int, [C.c_void_p, C.c_char_p]),
    "hdg_decompose_simple": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, _i32p]),
    "hdg_decompose_graph": (C.c_int, [C.c_void_p, C.c_int32, _i32p]),
    "hdg_decompose_from_dict": (C.c_int, [C.c_void_p, C.c_char_p, _i32p, _i32p]),
    "hdg_mesh_decompose": (C.c_int, [C.c_void_p, C.c_int32, _i32p, C.c_int32, C.c_void_p]),
    "hdg_mesh_proc_addressing": (C.c_int, [C.c_void_p, _i32p, _i32p, _i32p, _i32p]),
    "hdg_mesh_num_points": (C.c_int64, [C.c_void_p]),
    "hdg_mesh_counts": (C.c_int, [C.c_void_p, _i64p, _i64p, _i32p, _i64p]),
    "hdg_mesh_get_faces": (C.c_int, [C.c_void_p, _i32p, _i32p, _i32p, _i32p, _i32p]),
    "hdg_mesh_get_cell_vertices": (C.c_int, [C.c_void_p, _i32p]),
    "hdg_mesh_patch_info": (C.c_int, [C.c_void_p, C.c_int32, C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, _i32p]),
    "hdg_mesh_patch_faces": (C.c_int, [C.c_void_p, C.c_int32, _i32p]),
    "hdg_mesh_node_coords": (C.c_int, [C.c_void_p, _f64p]),
    "hdg_mesh_conn_codes": (C.c_int, [C.c_void_p, _i32p, C.c_int32, _i32p]),
    "hdg_mesh_boundary_slots": (C.c_int, [C.c_void_p, _i32p, _i32p]),
    "hdg_get_node_table": (C.c_int, [C.c_void_p, _i32p, C.c_int32]),
    "hdg_euler_limit": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double]),
    "hdg_limiter_weights": (C.c_int, [C.c_void_p, _f64p]),
    "hdg_state_freeze_traces": (C.c_int, [C.c_void_p, C.c_int32]),
    "hdg_state_thaw": (C.c_int, [C.c_void_p, C.c_int32]),
    "hdg_mesh_set_curved_patch": (C.c_int, [C.c_void_p, C.c_int32, _f64p]),
    "hdg_mesh_patch_node_coords": (C.c_int, [C.c_void_p, C.c_int32, _f64p]),
    "hdg_state_create": (C.c_int, [C.c_void_p, C.c_int32, _i32p]),
    "hdg_state_destroy": (C.c_int, [C.c_void_p, C.c_int32]),
    "hdg_state_upload": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    "hdg_state_download": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    "hdg_state_upload_async": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    "hdg_state_download_async": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    "hdg_state_set_patch_kind": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "hdg_state_set_patch_values": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    "hdg_euler_stage": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_double, C.c_double]),
    "hdg_euler_stage_range": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_double, C.c_double,
                                         C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "hdg_stream_wait": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "hdg_euler_step_ssprk2": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_int32]),
    "hdg_euler_step_lserk45": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_int32]),
    "hdg_advect_stage": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32, C.c_double, C.c_double]),
    "hdg_advect_step_ssprk2": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_int32]),
    "hdg_advect_step_lserk45": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_int32]),
    "hdg_state_copy": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "hdg_euler_stage_fields": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_double,
                                          C.c_double, C.c_int32, C.c_int32, C.c_int32]),
    "hdg_euler_stage_fields_ex": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hdg_state_copy_ghosts": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "hdg_state_swap": (C.c_int, [C.c_void_p, C.c_int32]),
    "hdg_state_axpby": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_int32, C.c_double, C.c_int32]),
    "hdg_state_l1_diff": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, _f64p]),
    "hdg_halo_counts": (C.c_int, [C.c_void_p, C.c_int32, _i64p]),
    "hdg_halo_bind": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]),
    "hdg_halo_pack": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _vpp, _i64p]),
    "hdg_halo_recv_buffer": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _vpp, _i64p]),
    "hdg_halo_unpack": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "hdg_mesh_get_points": (C.c_int, [C.c_void_p, _f64p]),
    "hdg_comm_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p]),
    "hdg_comm_rank_size": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "hdg_halo_exchange": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "hdg_euler_step_ssprk2_parallel": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_int32]),
    "hdg_group_euler_step_ssprk2": (C.c_int, [C.POINTER(C.c_void_p), _i32p, C.c_int32, C.c_double, C.c_double, C.c_int32]),
    "hdg_mesh_set_patch_neighbour": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "hdg_par_counts": (C.c_int, [C.c_void_p, _i64p, _i64p, _i64p, _i32p]),
    "hdg_comm_allreduce_sum": (C.c_int, [C.c_void_p, _f64p, C.c_int32]),
    "hdg_comm_allgather_i64": (C.c_int, [C.c_void_p, C.c_int64, _i64p]),
    "hdg_stream": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "hdg_launch_count": (C.c_int64, [C.c_void_p]),
    "hdg_euler_stage_kernels": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
    "hdg_measure_fp64_peak": (C.c_int, [C.c_void_p, C.c_double, _f64p]),
    "hdg_state_device_ptr": (C.c_void_p, [C.c_void_p, C.c_int32, C.c_int32]),
    "hdg_layout": (C.c_int, [C.c_void_p, _i64p, _i32p, _i32p, _i64p, _i64p, _i32p, _i32p]),
}


class EulerFieldsStage(C.Structure):
    """hdg_euler_fields_stage of include/hopedg.h"""
    _fields_ = [("s", C.c_int32 * 3), ("src", C.c_int32 * 3), ("aux", C.c_int32 * 3), ("out2", C.c_int32 * 3), ("aux2", C.c_int32 * 3),
                ("gamma", C.c_double), ("dt", C.c_double), ("a", C.c_double), ("b", C.c_double), ("a2", C.c_double), ("b2", C.c_double),
                ("fluxKind", C.c_int32), ("exchange", C.c_int32)]


def load_library(path: os.PathLike | None = None):
    """dlopen libhopedg.so and set the prototypes.  Raises if the library was not built."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    if path is None and os.environ.get("HDG_LIB_PATH"):      # kernel-variant experiments (tools/build_variant.sh)
        path = os.environ["HDG_LIB_PATH"]
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise RuntimeError(f"{p} not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C hopefoam_b200/csrc).  There is no CPU fallback.")
    lib = C.CDLL(str(p))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


class HdgError(RuntimeError):
    pass


def _ptr(a, typ):
    return a.ctypes.data_as(typ)


def _as_f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


class Context:
    """One GPU context (one per rank).  Thin, argument-checked mirror of the C ABI."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.hdg_create(device, C.byref(h))
        if rc != 0:
            raise HdgError(f"hdg_create failed: {self.lib.hdg_last_error(None).decode()}")
        self.h = h
        self.N = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.hdg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise HdgError(self.lib.hdg_last_error(self.h).decode())

    # ---- element / mesh -------------------------------------------------------------------------
    def set_order(self, N: int):
        self._ck(self.lib.hdg_set_order(self.h, N))
        self.N = N
        v = [C.c_int32() for _ in range(4)]
        self._ck(self.lib.hdg_get_sizes(self.h, *[C.byref(x) for x in v]))
        self.Np, self.Nfp, self.Ng, self.Nfg = [x.value for x in v]

    def operator(self, what: str, shape=None):
        n = self.lib.hdg_get_operator(self.h, what.encode(), None, 0)
        if n < 0:
            raise HdgError(f"unknown operator {what}")
        out = np.empty(n)
        self.lib.hdg_get_operator(self.h, what.encode(), _ptr(out, _f64p), n)
        return out.reshape(shape) if shape else out

    def face_to_cell_index(self):
        out = np.empty(3 * 2 * self.Nfp, dtype=np.int32)
        self._ck(self.lib.hdg_get_face_to_cell_index(self.h, _ptr(out, _i32p)))
        return out.reshape(3, 2, self.Nfp)

    def set_mesh_triangles(self, xy, tris, point_equiv=None, patch_edges=None):
        """patch_edges: list (per patch) of int arrays (n,3) = (cell, pointA, pointB)."""
        xy = _as_f64(xy)
        tris = np.ascontiguousarray(tris, dtype=np.int32)
        pe = None if point_equiv is None else np.ascontiguousarray(point_equiv, dtype=np.int32)
        patch_edges = patch_edges or []
        start = np.zeros(len(patch_edges) + 1, dtype=np.int32)
        cells, pts = [], []
        for i, e in enumerate(patch_edges):
            e = np.asarray(e, dtype=np.int32).reshape(-1, 3)
            start[i + 1] = start[i] + e.shape[0]
            cells.append(e[:, 0])
            pts.append(e[:, 1:3])
        ecell = np.ascontiguousarray(np.concatenate(cells) if cells else np.zeros(0), dtype=np.int32)
        epts = np.ascontiguousarray(np.concatenate(pts) if pts else np.zeros((0, 2)), dtype=np.int32)
        self._ck(self.lib.hdg_set_mesh_triangles(
            self.h, xy.shape[0], _ptr(xy, _f64p), tris.shape[0], _ptr(tris, _i32p),
            None if pe is None else _ptr(pe, _i32p), len(patch_edges), _ptr(start, _i32p), _ptr(ecell, _i32p), _ptr(epts, _i32p)))
        self._after_mesh()

    def set_mesh_polymesh(self, path):
        self._ck(self.lib.hdg_set_mesh_polymesh(self.h, str(path).encode()))
        self._after_mesh()

    def _after_mesh(self):
        K, F, nG = C.c_int64(), C.c_int64(), C.c_int64()
        nP = C.c_int32()
        self._ck(self.lib.hdg_mesh_counts(self.h, C.byref(K), C.byref(F), C.byref(nP), C.byref(nG)))
        self.K, self.F, self.n_patches, self.n_ghost = K.value, F.value, nP.value, nG.value

    def decompose_simple(self, nx, ny, nz=1, delta=0.001):
        out = np.empty(self.K, dtype=np.int32)
        self._ck(self.lib.hdg_decompose_simple(self.h, nx, ny, nz, delta, _ptr(out, _i32p)))
        return out

    def decompose_graph(self, n_procs):
        """cellToProc of the native graph partitioner (`method scotch | metis`)."""
        out = np.empty(self.K, dtype=np.int32)
        self._ck(self.lib.hdg_decompose_graph(self.h, n_procs, _ptr(out, _i32p)))
        return out

    def decompose_from_dict(self, case_dir):
        """cellToProc as <case>/system/decomposeParDict prescribes (method simple | manual)."""
        out = np.empty(self.K, dtype=np.int32)
        n = C.c_int32()
        self._ck(self.lib.hdg_decompose_from_dict(self.h, str(case_dir).encode(), C.byref(n), _ptr(out, _i32p)))
        return n.value, out

    def set_mesh_from_decomposition(self, global_ctx, cell_to_proc, n_procs, rank):
        """Processor mesh of `rank` (dgDecomposePar rules) built from the global mesh held by another context."""
        c2p = np.ascontiguousarray(cell_to_proc, dtype=np.int32)
        rc = self.lib.hdg_mesh_decompose(global_ctx.h, n_procs, _ptr(c2p, _i32p), rank, self.h)
        if rc != 0:
            raise HdgError(self.lib.hdg_last_error(self.h).decode())
        self._after_mesh()

    def proc_addressing(self):
        npts = self.lib.hdg_mesh_num_points(self.h)
        cell, point = np.empty(self.K, dtype=np.int32), np.empty(npts, dtype=np.int32)
        nbr, pf = np.empty(self.n_patches, dtype=np.int32), np.empty(self.n_ghost, dtype=np.int32)
        self._ck(self.lib.hdg_mesh_proc_addressing(self.h, _ptr(cell, _i32p), _ptr(point, _i32p), _ptr(nbr, _i32p), _ptr(pf, _i32p)))
        return {"cell": cell, "point": point, "patch_nbr_proc": nbr, "patch_face_global": pf}

    def faces(self):
        arrs = [np.empty(self.F, dtype=np.int32) for _ in range(5)]
        self._ck(self.lib.hdg_mesh_get_faces(self.h, *[_ptr(a, _i32p) for a in arrs]))
        return dict(zip(("owner", "nbr", "loc_o", "loc_n", "rot"), arrs))

    def cell_vertices(self):
        out = np.empty((self.K, 3), dtype=np.int32)
        self._ck(self.lib.hdg_mesh_get_cell_vertices(self.h, _ptr(out, _i32p)))
        return out

    def patch_info(self, p):
        name, typ = C.create_string_buffer(128), C.create_string_buffer(64)
        n = C.c_int32()
        self._ck(self.lib.hdg_mesh_patch_info(self.h, p, name, 128, typ, 64, C.byref(n)))
        return name.value.decode(), typ.value.decode(), n.value

    def patch_faces(self, p):
        n = self.patch_info(p)[2]
        out = np.empty(n, dtype=np.int32)
        if n:
            self._ck(self.lib.hdg_mesh_patch_faces(self.h, p, _ptr(out, _i32p)))
        return out

    def node_coords(self):
        out = np.empty((self.K, self.Np, 2))
        self._ck(self.lib.hdg_mesh_node_coords(self.h, _ptr(out, _f64p)))
        return out

    def patch_node_coords(self, p):
        n = self.patch_info(p)[2]
        out = np.empty((n * self.Nfp, 2))
        if n:
            self._ck(self.lib.hdg_mesh_patch_node_coords(self.h, p, _ptr(out, _f64p)))
        return out

    def set_curved_patch(self, p, positions):
        """positions: (nFaces*Nfp, 2) nodes of the patch faces on the curve (arcDgPatch::positions), patch-dof order."""
        pos = _as_f64(positions)
        assert pos.shape == (self.patch_info(p)[2] * self.Nfp, 2)
        self._ck(self.lib.hdg_mesh_set_curved_patch(self.h, p, _ptr(pos, _f64p)))

    def conn_codes(self, kinds):
        """(K,4) int32: neighbour element / ghost slot per face + packed code bytes, for per-patch HDG_BC_* kinds."""
        k = np.ascontiguousarray(kinds, dtype=np.int32)
        out = np.empty((self.K, 4), dtype=np.int32)
        self._ck(self.lib.hdg_mesh_conn_codes(self.h, _ptr(k, _i32p), k.size, _ptr(out, _i32p)))
        return out

    def boundary_slots(self):
        n_ghost = self.n_ghost
        bslot, first = np.empty((self.K, 3), dtype=np.int32), np.empty(max(n_ghost, 1), dtype=np.int32)
        self._ck(self.lib.hdg_mesh_boundary_slots(self.h, _ptr(bslot, _i32p), _ptr(first, _i32p)))
        return bslot, first[:n_ghost]

    def node_table(self):
        n = self.lib.hdg_get_node_table(self.h, None, 0)
        out = np.empty(n, dtype=np.int32)
        assert self.lib.hdg_get_node_table(self.h, _ptr(out, _i32p), n) == n
        return out

    def limiter_weights(self):
        out = np.empty(self.Np)
        self._ck(self.lib.hdg_limiter_weights(self.h, _ptr(out, _f64p)))
        return out

    def layout(self):
        Kpad, ps, gb = C.c_int64(), C.c_int64(), C.c_int64()
        a, b, eg, ag = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self.lib.hdg_layout(self.h, C.byref(Kpad), C.byref(a), C.byref(b), C.byref(ps), C.byref(gb), C.byref(eg), C.byref(ag)))
        return dict(Kpad=Kpad.value, NpPad=a.value, NfpPad=b.value, planeStride=ps.value, ghostBase=gb.value,
                    eulerGrid=eg.value, advectGrid=ag.value)

    # ---- states ------------------------------------------------------------------------------------
    def state_create(self, n_planes: int) -> int:
        sid = C.c_int32()
        self._ck(self.lib.hdg_state_create(self.h, n_planes, C.byref(sid)))
        return sid.value

    def state_destroy(self, sid):
        self._ck(self.lib.hdg_state_destroy(self.h, sid))

    def upload(self, sid, plane0, host):
        """host: (K,Np) scalar or (K,Np,c) array; component i -> plane plane0+i (all c components)."""
        host = _as_f64(host)
        stride = 1 if host.ndim == 2 else host.shape[2]
        assert host.shape[0] == self.K and host.shape[1] == self.Np
        self._ck(self.lib.hdg_state_upload(self.h, sid, plane0, stride, host.ctypes.data, stride))

    def upload_ptr(self, sid, plane0, n_planes, ptr, stride):
        self._ck(self.lib.hdg_state_upload(self.h, sid, plane0, n_planes, ptr, stride))

    def download(self, sid, plane0, n_planes=1):
        out = np.empty((self.K, self.Np, n_planes))
        self._ck(self.lib.hdg_state_download(self.h, sid, plane0, n_planes, out.ctypes.data, n_planes))
        return out[..., 0] if n_planes == 1 else out

    def download_ptr(self, sid, plane0, n_planes, ptr, stride):
        self._ck(self.lib.hdg_state_download(self.h, sid, plane0, n_planes, ptr, stride))

    def upload_ptr_async(self, sid, plane0, n_planes, ptr, stride):
        self._ck(self.lib.hdg_state_upload_async(self.h, sid, plane0, n_planes, ptr, stride))

    def download_ptr_async(self, sid, plane0, n_planes, ptr, stride):
        self._ck(self.lib.hdg_state_download_async(self.h, sid, plane0, n_planes, ptr, stride))

    def set_patch_kind(self, sid, patch, kind):
        self._ck(self.lib.hdg_state_set_patch_kind(self.h, sid, patch, kind))

    def set_patch_values(self, sid, plane0, patch, values):
        values = _as_f64(values)
        stride = 1 if values.ndim == 1 else values.shape[1]
        self._ck(self.lib.hdg_state_set_patch_values(self.h, sid, plane0, stride, patch, values.ctypes.data, stride))

    def state_copy(self, dst, src):
        self._ck(self.lib.hdg_state_copy(self.h, dst, src))

    def euler_stage_fields(self, s_rho, s_rhou, s_e, gamma, dt, a=0.0, b=1.0, aux=(0, 0, 0), flux=FLUX_ROE):
        self._ck(self.lib.hdg_euler_stage_fields(self.h, s_rho, s_rhou, s_e, gamma, dt, flux, a, b, *aux))

    def euler_stage_fields_ex(self, s, gamma, dt, src=(-1, -1, -1), a=0.0, b=1.0, aux=(0, 0, 0), out2=(-1, -1, -1), a2=0.0, b2=0.0,
                              aux2=(0, 0, 0), flux=FLUX_ROE, exchange=0):
        """hdg_euler_stage_fields_ex: nodal data from src, boundary data of s, result -> stage copies of s, optional second result -> out2."""
        st = EulerFieldsStage((C.c_int32 * 3)(*s), (C.c_int32 * 3)(*src), (C.c_int32 * 3)(*aux), (C.c_int32 * 3)(*out2), (C.c_int32 * 3)(*aux2),
                              gamma, dt, a, b, a2, b2, flux, exchange)
        self._ck(self.lib.hdg_euler_stage_fields_ex(self.h, C.byref(st)))

    def state_copy_ghosts(self, dst, src):
        self._ck(self.lib.hdg_state_copy_ghosts(self.h, dst, src))

    def euler_limit(self, s_rho, s_rhou, s_e, gamma=1.4, eps=1e-10, tol=1e-2):
        """Godunov.limite(rho, rhoU, Ener) with the Triangle limiter, in place (Trianglelimite.C:61-864)."""
        self._ck(self.lib.hdg_euler_limit(self.h, s_rho, s_rhou, s_e, gamma, eps, tol))

    def freeze_traces(self, sid):
        self._ck(self.lib.hdg_state_freeze_traces(self.h, sid))

    def thaw(self, sid):
        self._ck(self.lib.hdg_state_thaw(self.h, sid))

    def state_swap(self, sid):
        self._ck(self.lib.hdg_state_swap(self.h, sid))

    def state_axpby(self, dst, a, x, b, y):
        self._ck(self.lib.hdg_state_axpby(self.h, dst, a, x, b, y))

    def l1_diff(self, sid, plane, ref):
        ref = _as_f64(ref)
        out = C.c_double()
        self._ck(self.lib.hdg_state_l1_diff(self.h, sid, plane, ref.ctypes.data, 1, C.byref(out)))
        return out.value

    # ---- hot path ----------------------------------------------------------------------------------
    def euler_stage(self, sid, gamma, dt, stage, a, b, flux=FLUX_ROE):
        self._ck(self.lib.hdg_euler_stage(self.h, sid, gamma, dt, flux, stage, a, b))

    def euler_stage_range(self, sid, gamma, dt, stage, a, b, elem_begin, elem_end, elem_begin2=0, elem_end2=0, flux=FLUX_ROE):
        self._ck(self.lib.hdg_euler_stage_range(self.h, sid, gamma, dt, flux, stage, a, b, elem_begin, elem_end, elem_begin2, elem_end2))

    def stream_wait(self, waiter, signaler):
        self._ck(self.lib.hdg_stream_wait(self.h, waiter, signaler))

    def euler_step_ssprk2(self, sid, gamma, dt, flux=FLUX_ROE):
        self._ck(self.lib.hdg_euler_step_ssprk2(self.h, sid, gamma, dt, flux))

    def euler_step_lserk45(self, sid, gamma, dt, flux=FLUX_ROE):
        self._ck(self.lib.hdg_euler_step_lserk45(self.h, sid, gamma, dt, flux))

    def advect_stage(self, sT, sU, dt, stage, a, b, flux=FLUX_LF):
        self._ck(self.lib.hdg_advect_stage(self.h, sT, sU, dt, flux, stage, a, b))

    def advect_step_ssprk2(self, sT, sU, dt, flux=FLUX_LF):
        self._ck(self.lib.hdg_advect_step_ssprk2(self.h, sT, sU, dt, flux))

    def advect_step_lserk45(self, sT, sU, dt, flux=FLUX_LF):
        self._ck(self.lib.hdg_advect_step_lserk45(self.h, sT, sU, dt, flux))

    def sync(self):
        self._ck(self.lib.hdg_sync(self.h))

    def launch_count(self):
        return self.lib.hdg_launch_count(self.h)

    def euler_stage_kernels(self):
        """Names of the kernels one full-mesh Euler stage launches at this order (split: face-flux + element kernel)."""
        buf = C.create_string_buffer(256)
        self.lib.hdg_euler_stage_kernels(self.h, buf, 256)
        return buf.value.decode().split("+")

    def measure_fp64_peak(self, seconds=1.0):
        out = C.c_double()
        self._ck(self.lib.hdg_measure_fp64_peak(self.h, seconds, C.byref(out)))
        return out.value

    def stream(self, which=0):
        return self.lib.hdg_stream(self.h, which)

    # ---- one process per GPU: communicator + overlapped exchange owned by the library --------------------
    def comm_init(self, rank, world, id_file):
        self._ck(self.lib.hdg_comm_init(self.h, rank, world, str(id_file).encode()))

    def halo_exchange(self, sid, which=0):
        self._ck(self.lib.hdg_halo_exchange(self.h, sid, which))

    def euler_step_ssprk2_parallel(self, sid, gamma, dt, flux=FLUX_ROE):
        self._ck(self.lib.hdg_euler_step_ssprk2_parallel(self.h, sid, gamma, dt, flux))

    def set_patch_neighbour(self, patch, rank, tag=0):
        self._ck(self.lib.hdg_mesh_set_patch_neighbour(self.h, patch, rank, tag))

    def par_counts(self):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        self._ck(self.lib.hdg_par_counts(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"proc_faces": a.value, "boundary_octets": b.value, "interior_octets": c.value, "neighbours": d.value}

    # ---- halo ---------------------------------------------------------------------------------------
    def halo_count(self, patch):
        n = C.c_int64()
        self._ck(self.lib.hdg_halo_counts(self.h, patch, C.byref(n)))
        return n.value

    def halo_bind(self, patch, send_ptr, recv_ptr, cap):
        self._ck(self.lib.hdg_halo_bind(self.h, patch, send_ptr, recv_ptr, cap))

    def halo_pack(self, sid, which, patch):
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self.lib.hdg_halo_pack(self.h, sid, which, patch, C.byref(p), C.byref(n)))
        return p.value, n.value

    def halo_unpack(self, sid, which, patch):
        self._ck(self.lib.hdg_halo_unpack(self.h, sid, which, patch))


def group_euler_step_ssprk2(ctxs, sids, gamma, dt, flux=FLUX_ROE):
    """hdg_group_euler_step_ssprk2: one SSP-RK2 step on the contexts of THIS process (index = processor number), peer-copy transport."""
    n = len(ctxs)
    arr = (C.c_void_p * n)(*[c.h for c in ctxs])
    ids = np.ascontiguousarray(sids, dtype=np.int32)
    rc = ctxs[0].lib.hdg_group_euler_step_ssprk2(arr, _ptr(ids, _i32p), n, gamma, dt, flux)
    if rc != 0:
        raise HdgError(ctxs[0].lib.hdg_last_error(ctxs[0].h).decode())
