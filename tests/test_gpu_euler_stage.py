"""GPU parity (through the C ABI) of the fused Euler stage against the oracle: <= 1e-12 relative L2 per stage
(BASELINE.json north_star tolerance), for every supported order and every boundary kind."""
import numpy as np
import pytest

from hopefoam_b200 import capi, meshgen
from oracle import dg_oracle as o
from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL_STAGE = 1e-12


def _case(N, n=6, periodic=False, kinds=None):
    mg = meshgen.jittered_square(n, periodic=periodic)
    om = H.oracle_mesh(mg)
    case = o.Case(om, N, bc_kinds=kinds)
    return mg, case


def _run_stage(ctx, mg, case, t0=0.3, dt=1e-3, gamma=1.4):
    ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    x, y = case.geo.x[..., 0], case.geo.x[..., 1]
    rho, rhoU, E = H.vortex_state(x, y, t0, gamma)
    npatch = len(case.mesh.patches)
    bR, bU, bE = [], [], []
    for ip in range(npatch):
        xy = case.patch_internal(case.geo.x, ip)
        r, u, e = H.vortex_state(xy[:, 0], xy[:, 1], t0 + 0.01, gamma)     # boundary data differs from the trace
        bR.append(r), bU.append(u), bE.append(e)
    # BC evaluate on the oracle side for the non-fixed kinds (fields are "already corrected")
    case.evaluate_bc(rho, bR)
    case.evaluate_bc(rhoU, bU, is_vector=True)
    case.evaluate_bc(E, bE)
    sid = H.setup_euler(ctx, case, rho, rhoU, E, bR, bU, bE, case.bc_kinds)
    # the boundary lists are updated in place by correctBoundaryConditions() after each solve (dgMatrixSolve.C:209):
    # fixedValue keeps its (stale, t_n) data, zeroGradient/reflective follow the new interior trace
    r1, u1, e1 = o.euler_stage(case, rho, rhoU, E, bR, bU, bE, gamma, dt)
    ctx.euler_stage(sid, gamma, dt, 0, 0.0, 1.0)
    # stage result lives in the stage copy; run the second SSP stage too and compare the combined step
    r2, u2, e2 = o.euler_stage(case, r1, u1, e1, bR, bU, bE, gamma, dt)
    ctx.euler_stage(sid, gamma, dt, 1, 0.5, 0.5)
    ctx.sync()
    g = H.download_euler(ctx, sid)
    ref = (0.5 * rho + 0.5 * r2, 0.5 * rhoU + 0.5 * u2, 0.5 * E + 0.5 * e2)
    # increments are what the kernel computes: compare q^{n+1}-q^n too (tighter than the fields themselves)
    errs = [H.rel_l2(a, b) for a, b in zip(g, ref)]
    inc = [H.rel_l2(a - q0, b - q0) for a, b, q0 in zip(g, ref, (rho, rhoU, E))]
    ctx.state_destroy(sid)
    return errs, inc


@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10])      # 9, 10: beyond the reference's cubature table, own collapsed rule
def test_euler_step_fixed_value_all_orders(gpu_ctx_factory, N):
    ctx = gpu_ctx_factory(N)
    mg, case = _case(N, n=5)
    errs, inc = _run_stage(ctx, mg, case)
    assert max(errs) <= TOL_STAGE, (errs, inc)
    # the O(dt) increment itself (what the kernel computes): dt = 1e-3 puts the round-off floor of q (1e-16 x the conditioning of the
    # order-N operators) at ~1e-12 (N + 1)^2 relative to the increment
    assert max(inc) <= 2e-12 * (N + 1) ** 2, inc
    ctx.close()


@pytest.mark.parametrize("N", [2, 4])
def test_euler_step_periodic(gpu_ctx_factory, N):
    ctx = gpu_ctx_factory(N)
    mg, case = _case(N, n=6, periodic=True)
    assert len(case.mesh.patches) == 0 and (case.mesh.face_nbr >= 0).all()
    errs, inc = _run_stage(ctx, mg, case)
    assert max(errs) <= TOL_STAGE, (errs, inc)
    ctx.close()


@pytest.mark.parametrize("kind", [o.BC_ZEROGRAD, o.BC_REFLECTIVE])
def test_euler_step_wall_kinds(gpu_ctx_factory, kind):
    ctx = gpu_ctx_factory(4)
    mg, case = _case(4, n=5, kinds=[kind])
    errs, inc = _run_stage(ctx, mg, case)
    assert max(errs) <= TOL_STAGE, (errs, inc)
    ctx.close()


def test_ragged_element_count(gpu_ctx_factory):
    """K not a multiple of the 8-element warp tile: n=3 -> 18 triangles."""
    ctx = gpu_ctx_factory(4)
    mg, case = _case(4, n=3)
    assert case.mesh.K % 8 != 0
    errs, inc = _run_stage(ctx, mg, case)
    assert max(errs) <= TOL_STAGE, (errs, inc)
    ctx.close()


def test_cylinder_style_mixed_patches_N6(gpu_ctx_factory):
    """BASELINE configs[4] boundary set at its order (N=6): a reflective (slip) wall patch and fixedValue far-field patches on
    the same mesh, each with its own patch-dof numbering."""
    N = 6
    ctx = gpu_ctx_factory(N)
    mg = meshgen.jittered_square(5)
    e = mg["patch_edges"][0]
    n = 5
    mg["patch_edges"] = [e[0:n], e[n:4 * n]]            # bottom side = wall, the other three = far field
    om = H.oracle_mesh(mg)
    case = o.Case(om, N, bc_kinds=[o.BC_REFLECTIVE, o.BC_FIXED])
    errs, inc = _run_stage(ctx, mg, case)
    assert max(errs) <= TOL_STAGE, (errs, inc)
    ctx.close()


def test_two_triangle_mesh(gpu_ctx_factory):
    """Smallest mesh: one quad split into two triangles, five of the six faces on the patch (K < 8: a single partial octet)."""
    ctx = gpu_ctx_factory(4)
    mg, case = _case(4, n=1)
    assert case.mesh.K == 2 and case.mesh.F == 5 and int((case.mesh.face_nbr >= 0).sum()) == 1
    errs, inc = _run_stage(ctx, mg, case, dt=1e-3)
    assert max(errs) <= TOL_STAGE, (errs, inc)
    ctx.close()


@pytest.mark.parametrize("n", [1, 2, 3])
def test_euler_smallest_meshes(gpu_ctx_factory, n):
    """2, 8 and 18 triangles: one ragged / exactly full octet and two octets with a ragged tail, all elements on the boundary."""
    ctx = gpu_ctx_factory(4)
    mg, case = _case(4, n=n)
    errs, inc = _run_stage(ctx, mg, case)
    assert max(errs) <= TOL_STAGE, (errs, inc)
