"""The product's `Triangle` limiter arithmetic (hopefoam_b200/csrc/dg_limiter_core.hpp - the inline functions that the CUDA kernels of
dg_limiter.cu wrap one-to-one) run in HOST loops by a test harness (tests/native/limiter_host_check.cpp) on the product's own data layout
(padded planes + ghost traces) and the product's own topology arrays (hdg_mesh_conn_codes / hdg_mesh_boundary_slots / hdg_get_node_table
from a host-only context), compared with the numpy restatement of the reference (oracle.triangle_limit, Trianglelimite.C:61-864).

This verifies the arithmetic and all indexing of the device code on the CPU.  What it cannot verify is the launch glue of
hdg_euler_limit (buffer carving, stream order): that is tests/test_gpu_limiter.py."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from hopefoam_b200 import capi, meshgen
from oracle import dg_oracle as o
from tests import helpers as H

ROOT = Path(__file__).resolve().parent.parent
KIND = {o.BC_FIXED: capi.BC_FIXED_VALUE, o.BC_ZEROGRAD: capi.BC_ZERO_GRADIENT, o.BC_REFLECTIVE: capi.BC_REFLECTIVE}


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = tmp_path_factory.mktemp("limiter") / "liblimiter_host_check.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Werror", str(ROOT / "tests/native/limiter_host_check.cpp"),
                    "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    lib.limiter_host_run.restype = C.c_int
    lib.limiter_host_run.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, ip, ip, ip, ip,
                                     dp, dp, dp, dp, ip, dp, C.c_double, C.c_double, C.c_double, C.c_int]
    return lib


def _mesh(n, wall=False):
    mg = meshgen.jittered_square(n)
    if wall:
        e = mg["patch_edges"][0]
        om = o.build_connectivity(mg["xy"], mg["tris"], [[(int(c), (int(a), int(b))) for c, a, b in e]],
                                  [{"name": "wall", "type": "wall"}], point_equiv=mg["point_equiv"])
    else:
        om = H.oracle_mesh(mg)
    return mg, om


def _bvals(case, rho, U, E, fixed_fn=None):
    bR, bU, bE = [case.patch_internal(rho, 0)], [case.patch_internal(U, 0)], [case.patch_internal(E, 0)]
    if fixed_fn is not None:
        bR, bU, bE = fixed_fn(bR[0], bU[0], bE[0])
    case.evaluate_bc(rho, bR)
    case.evaluate_bc(U, bU, is_vector=True)
    case.evaluate_bc(E, bE)
    return bR, bU, bE


def run_product_core(lib, ctx, fields, bvals, kind, gamma=1.4, eps=1e-10, tol=1e-2, fused=0):
    """fields = (rho (K,Np), U (K,Np,2), E); bvals = per-patch lists as the oracle takes them.  Returns the limited fields."""
    L = ctx.layout()
    K, Np, Nfp, NpPad, NfpPad, gb = ctx.K, ctx.Np, ctx.Nfp, L["NpPad"], L["NfpPad"], L["ghostBase"]
    rho, U, E = fields
    planes = [np.full(L["planeStride"], np.nan) for _ in range(4)]          # NaN: anything read outside the written region shows up
    for pl, f in zip(planes, (rho, U[..., 0], U[..., 1], E)):
        v = pl[:L["Kpad"] * NpPad].reshape(L["Kpad"], NpPad)
        v[:K, :Np] = f
    kinds_o = list(kind) if isinstance(kind, (list, tuple)) else [kind] * ctx.n_patches
    bR, bU, bE = bvals
    start = 0
    for ip, k_o in enumerate(kinds_o):                                       # ghost traces, as hdg_state_set_patch_values lays them out
        nf = ctx.patch_info(ip)[2]
        if k_o == o.BC_FIXED and nf:
            for pl, b in zip(planes, (bR[ip], bU[ip][:, 0], bU[ip][:, 1], bE[ip])):
                g = pl[gb + start * NfpPad:gb + (start + nf) * NfpPad].reshape(nf, NfpPad)
                g[:, :Nfp] = b.reshape(nf, Nfp)
        start += nf
    kinds = [KIND[k_o] for k_o in kinds_o]
    conn = ctx.conn_codes(kinds)
    bslot, first = ctx.boundary_slots()
    first = np.ascontiguousarray(np.concatenate([first, [0]]), dtype=np.int32)
    tris = ctx.cell_vertices()
    verts = np.ascontiguousarray(ctx_points(ctx)[tris].reshape(K, 6))
    r, s, mpp, tab = ctx.operator("r"), ctx.operator("s"), ctx.limiter_weights(), ctx.node_table()
    work = np.full(56 * K + 16 * ctx.n_ghost, np.nan)
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    P = lambda a, t: a.ctypes.data_as(t)
    rc = lib.limiter_host_run(K, ctx.n_ghost, gb, Np, NpPad, Nfp, NfpPad, *(P(p, dp) for p in planes), P(conn, ip), P(conn, ip),
                              P(bslot, ip), P(first, ip), P(verts, dp), P(r, dp), P(s, dp), P(mpp, dp), P(tab, ip), P(work, dp),
                              gamma, eps, tol, fused)
    assert rc == 0
    out = [pl[:L["Kpad"] * NpPad].reshape(L["Kpad"], NpPad)[:K, :Np].copy() for pl in planes]
    return out[0], np.stack([out[1], out[2]], -1), out[3]


def ctx_points(ctx):
    n = ctx.lib.hdg_mesh_num_points(ctx.h)
    out = np.empty((n, 2))
    assert ctx.lib.hdg_mesh_get_points(ctx.h, out.ctypes.data_as(C.POINTER(C.c_double))) == 0
    return out


def _host_ctx(mg, N, wall=False):
    c = H.HostContext()
    c.set_order(N)
    pe = [mg["patch_edges"][0]] if wall else mg["patch_edges"]
    c.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], pe)
    return c


def _smooth_state(case, amp=0.2):
    x, y = case.geo.x[..., 0], case.geo.x[..., 1]
    rho = 1.0 + amp * np.sin(0.7 * x + 0.3) * np.cos(0.5 * y)
    uu, vv = 0.4 + amp * np.sin(0.4 * y), -0.2 + amp * np.cos(0.6 * x + 0.2 * y)
    p = 1.0 + amp * np.cos(0.3 * x) * np.sin(0.8 * y + 0.1)
    return rho, np.stack([rho * uu, rho * vv], -1), p / 0.4 + 0.5 * rho * (uu ** 2 + vv ** 2)


def _shock_state(case):
    x, y = case.geo.x[..., 0], case.geo.x[..., 1]
    left = (x + 0.3 * y) < 5.0
    rho = np.where(left, 8.0, 1.4)
    uu, vv = np.where(left, 7.1, 0.0), np.where(left, -4.1, 0.0)
    p = np.where(left, 116.5, 1.0)
    return rho, np.stack([rho * uu, rho * vv], -1), p / 0.4 + 0.5 * rho * (uu ** 2 + vv ** 2)


@pytest.mark.parametrize("N", [1, 2, 3, 4, 6])
@pytest.mark.parametrize("kind", [o.BC_ZEROGRAD, o.BC_FIXED, o.BC_REFLECTIVE])
def test_core_matches_oracle_smooth(built_library, harness, N, kind):
    wall = kind == o.BC_REFLECTIVE
    mg, om = _mesh(7, wall)
    case = o.Case(om, N, bc_kinds=[kind])
    ctx = _host_ctx(mg, N, wall)
    rho, U, E = _smooth_state(case)
    fixed = None
    if kind == o.BC_FIXED:                    # boundary data that differ from the interior trace, varying along the patch
        fixed = lambda r, u, e: ([r * 1.05 + 0.01], [u * 0.9 + 0.02], [e * 1.02])
    bv = _bvals(case, rho, U, E, fixed)
    want = o.triangle_limit(case, rho, U, E, *bv)
    got = run_product_core(harness, ctx, (rho, U, E), bv, kind)
    for g, w in zip(got, want):
        assert np.isfinite(g).all()
        assert np.abs(g - w).max() <= 2e-11 * np.abs(w).max()
    moved = max(np.abs(w - f).max() for w, f in zip(want, (rho, U, E)))
    assert moved > 1e-3                       # the limiter did something (P_N -> P1)


def test_fused_form_equals_the_five_pass_form(built_library, harness):
    """The device runs three launches (dg_limiter.cu): averages + vertex values + ghost cells | per-element face gradients in the dgFace
    owner's role + cell gradient | reconstruction.  Same arithmetic as the five passes up to the end-point coordinates (vertex
    coordinates instead of the affine map of the vertex nodes): equal to round-off, and equal to the oracle within the same bound."""
    mg, om = multi_patch_mesh(8)
    kinds = [o.BC_REFLECTIVE, o.BC_FIXED, o.BC_ZEROGRAD, o.BC_FIXED]
    for N in (1, 4, 5):
        case = o.Case(om, N, bc_kinds=kinds)
        ctx = H.HostContext()
        ctx.set_order(N)
        ctx.set_mesh_triangles(mg["xy"], mg["tris"], None, mg["patch_edges"])
        rho, U, E = _smooth_state(case, amp=0.3)
        bv = [[case.patch_internal(f, ip) for ip in range(4)] for f in (rho, U, E)]
        case.evaluate_bc(rho, bv[0]); case.evaluate_bc(U, bv[1], is_vector=True); case.evaluate_bc(E, bv[2])
        a = run_product_core(harness, ctx, (rho, U, E), bv, kinds)
        b = run_product_core(harness, ctx, (rho, U, E), bv, kinds, fused=1)
        want = o.triangle_limit(case, rho, U, E, *bv)
        for x, y, w in zip(a, b, want):
            assert np.abs(x - y).max() <= 1e-12 * np.abs(x).max()
            assert np.abs(y - w).max() <= 2e-11 * np.abs(w).max()


@pytest.mark.parametrize("kind", [o.BC_ZEROGRAD, o.BC_FIXED])
def test_core_matches_oracle_across_a_shock(built_library, harness, kind):
    mg, om = _mesh(9)
    case = o.Case(om, 3, bc_kinds=[kind])
    ctx = _host_ctx(mg, 3)
    rho, U, E = _shock_state(case)
    bv = _bvals(case, rho, U, E)
    want = o.triangle_limit(case, rho, U, E, *bv)
    got = run_product_core(harness, ctx, (rho, U, E), bv, kind)
    for g, w in zip(got, want):
        assert np.abs(g - w).max() <= 1e-10 * np.abs(w).max()


def test_core_density_floor_and_bounded_loop(built_library, harness):
    """Density floor as in the oracle; a cell whose MEAN is below tol makes the reference (and the oracle) loop forever - the product
    bounds the loop and gives the cell a zero density slope."""
    mg, om = _mesh(4)
    case = o.Case(om, 2, bc_kinds=[o.BC_ZEROGRAD])
    ctx = _host_ctx(mg, 2)
    x = case.geo.x[..., 0]
    rho = 0.012 + 0.04 * np.maximum(x - 5.0, 0.0)
    rho[case.mesh.K // 2] *= 0.9
    U, E = np.zeros(x.shape + (2,)), np.full_like(x, 2.5)
    bv = _bvals(case, rho, U, E)
    want = o.triangle_limit(case, rho, U, E, *bv)
    got = run_product_core(harness, ctx, (rho, U, E), bv, o.BC_ZEROGRAD)
    assert got[0].min() >= 1e-2 - 1e-15
    for g, w in zip(got, want):
        assert np.abs(g - w).max() <= 1e-11 * max(np.abs(w).max(), 1.0)
    rho2 = rho.copy()
    rho2[3] = 0.004                                             # mean below tol: oracle would hang, so no oracle call here
    got2 = run_product_core(harness, ctx, (rho2, U, E), _bvals(case, rho2, U, E), o.BC_ZEROGRAD)
    assert np.isfinite(got2[0]).all() and np.abs(got2[0][3] - 0.004).max() < 1e-15


def multi_patch_mesh(n):
    """jittered square with its four sides as separate patches (bottom = wall)."""
    mg = meshgen.jittered_square(n)
    e = mg["patch_edges"][0]
    sides = [e[i * n:(i + 1) * n] for i in range(4)]
    names = ["wall", "outlet", "far", "inlet"]
    om = o.build_connectivity(mg["xy"], mg["tris"], [[(int(c), (int(a), int(b))) for c, a, b in sd] for sd in sides],
                              [{"name": nm, "type": "wall" if nm == "wall" else "patch"} for nm in names], point_equiv=None)
    mg = dict(mg, patch_edges=sides)
    return mg, om


@pytest.mark.parametrize("N", [2, 4])
def test_core_matches_oracle_mixed_patch_kinds(built_library, harness, N):
    """doubleMach-like boundary set: reflective wall, two fixedValue patches with different data, one zeroGradient patch - exercises the
    per-patch ghost-slot offsets (first value of EACH fixedValue patch, Trianglelimite.C:222-225)."""
    mg, om = multi_patch_mesh(8)
    kinds = [o.BC_REFLECTIVE, o.BC_FIXED, o.BC_ZEROGRAD, o.BC_FIXED]
    case = o.Case(om, N, bc_kinds=kinds)
    ctx = H.HostContext()
    ctx.set_order(N)
    ctx.set_mesh_triangles(mg["xy"], mg["tris"], None, mg["patch_edges"])
    assert ctx.n_patches == 4
    rho, U, E = _smooth_state(case, amp=0.3)
    bR = [case.patch_internal(rho, ip) for ip in range(4)]
    bU = [case.patch_internal(U, ip) for ip in range(4)]
    bE = [case.patch_internal(E, ip) for ip in range(4)]
    bR[1], bU[1], bE[1] = bR[1] * 1.1, bU[1] * 0.8 + 0.05, bE[1] * 1.05
    bR[3], bU[3], bE[3] = bR[3] * 0.9 + 0.02, bU[3] * 1.2, bE[3] * 0.97
    case.evaluate_bc(rho, bR)
    case.evaluate_bc(U, bU, is_vector=True)
    case.evaluate_bc(E, bE)
    want = o.triangle_limit(case, rho, U, E, bR, bU, bE)
    got = run_product_core(harness, ctx, (rho, U, E), (bR, bU, bE), kinds)
    for g, w in zip(got, want):
        assert np.abs(g - w).max() <= 2e-11 * np.abs(w).max()
    _, first = ctx.boundary_slots()
    assert sorted(set(first.tolist())) == [0, 8, 16, 24]


@pytest.mark.parametrize("periodic", [False, True])
def test_conn_codes_match_an_independent_construction(built_library, periodic):
    """hdg_mesh_conn_codes (the array the stage kernels and the limiter read) against the oracle's connectivity."""
    mg = meshgen.jittered_square(8, periodic=periodic)
    om = H.oracle_mesh(mg)
    ctx = H.HostContext()
    ctx.set_order(2)
    ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    for kind in (capi.BC_FIXED_VALUE, capi.BC_ZERO_GRADIENT, capi.BC_REFLECTIVE):
        conn = ctx.conn_codes([kind] * ctx.n_patches)
        bslot, first = ctx.boundary_slots()
        ghost = {}
        g = 0
        for p in om.patches:
            for f in p["faces"]:
                ghost[int(f)] = g
                g += 1
        for k in range(om.K):
            for lf in range(3):
                f = int(om.cell_face[k, lf])
                code = (int(conn[k, 3]) >> (8 * lf)) & 0xff
                owner = om.face_owner[f] == k and om.face_loc_o[f] == lf
                assert bool(code & 0x20) == owner
                if om.face_nbr[f] >= 0:
                    assert bslot[k, lf] == -1
                    assert conn[k, lf] == (om.face_nbr[f] if owner else om.face_owner[f])
                    assert (code & 3) == (om.face_loc_n[f] if owner else om.face_loc_o[f])
                    assert bool(code & 4) == (om.face_rot[f] == 1)
                    assert not code & 0x18
                else:
                    assert bslot[k, lf] == ghost[f]
                    if kind == capi.BC_FIXED_VALUE:
                        assert code & 8 and conn[k, lf] == ghost[f]
                    else:
                        assert conn[k, lf] == k and (code & 3) == lf and bool(code & 0x10) == (kind == capi.BC_REFLECTIVE)
        if not periodic:
            assert (first == 0).all() or ctx.n_patches > 1
            # frozen variant (hdg_state_freeze_traces): trace-type patches read their ghost slots, the reflective flag stays
            if kind != capi.BC_FIXED_VALUE:
                fr = ctx.conn_codes([kind | 0x100] * ctx.n_patches)
                for k in range(om.K):
                    for lf in range(3):
                        f = int(om.cell_face[k, lf])
                        code = (int(fr[k, 3]) >> (8 * lf)) & 0xff
                        if om.face_nbr[f] >= 0:
                            assert fr[k, lf] == conn[k, lf] and code == ((int(conn[k, 3]) >> (8 * lf)) & 0xff)
                        else:
                            assert fr[k, lf] == ghost[f] and code & 8 and code & 0x20
                            assert bool(code & 0x10) == (kind == capi.BC_REFLECTIVE)
