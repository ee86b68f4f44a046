"""hdg_euler_limit (Godunov.limite with `limiteScheme Triangle`) on the GPU against oracle.triangle_limit, through the C ABI with the
fields held as three states (rho | rhoU | Ener) as the facade keeps them.

The arithmetic these kernels run is also verified on the host (tests/test_limiter_core_host.py); this module checks the device launch
path (first green run on a B200: profiles/limiter_gpu_r01.txt)."""
import numpy as np
import pytest

from hopefoam_b200 import capi
from oracle import dg_oracle as o
from tests import helpers as H
from tests import test_limiter_core_host as T

pytestmark = pytest.mark.gpu


def _gpu_limit(ctx, case, fields, bv, kind):
    rho, U, E = fields
    sid = [ctx.state_create(1), ctx.state_create(2), ctx.state_create(1)]
    for s, f in zip(sid, (rho, U, E)):
        ctx.upload(s, 0, f)
        for ip in range(ctx.n_patches):
            ctx.set_patch_kind(s, ip, T.KIND[kind])
    if kind == o.BC_FIXED:
        for s, b in zip(sid, bv):
            ctx.set_patch_values(s, 0, 0, b[0])
    ctx.euler_limit(*sid)
    out = ctx.download(sid[0], 0), ctx.download(sid[1], 0, 2), ctx.download(sid[2], 0)
    for s in sid:
        ctx.state_destroy(s)
    return out


@pytest.mark.parametrize("N", [1, 3, 4])
@pytest.mark.parametrize("kind", [o.BC_ZEROGRAD, o.BC_FIXED, o.BC_REFLECTIVE])
def test_limit_matches_oracle(gpu_ctx_factory, N, kind):
    wall = kind == o.BC_REFLECTIVE
    mg, om = T._mesh(7, wall)
    case = o.Case(om, N, bc_kinds=[kind])
    ctx = gpu_ctx_factory(N)
    ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], [mg["patch_edges"][0]] if wall else mg["patch_edges"])
    rho, U, E = T._smooth_state(case)
    fixed = (lambda r, u, e: ([r * 1.05 + 0.01], [u * 0.9 + 0.02], [e * 1.02])) if kind == o.BC_FIXED else None
    bv = T._bvals(case, rho, U, E, fixed)
    want = o.triangle_limit(case, rho, U, E, *bv)
    n0 = ctx.launch_count()
    got = _gpu_limit(ctx, case, (rho, U, E), bv, kind)
    assert ctx.launch_count() - n0 >= 3
    for g, w in zip(got, want):
        assert np.abs(g - w).max() <= 2e-11 * np.abs(w).max()


def test_limit_then_stage_keeps_running(gpu_ctx_factory):
    """limit -> Euler stage -> limit on a discontinuous state: finite, and the cell means of rho survive each limit call."""
    mg, om = T._mesh(9)
    case = o.Case(om, 3, bc_kinds=[o.BC_ZEROGRAD])
    ctx = gpu_ctx_factory(3)
    ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    rho, U, E = T._shock_state(case)
    sid = [ctx.state_create(1), ctx.state_create(2), ctx.state_create(1)]
    for s, f in zip(sid, (rho, U, E)):
        ctx.upload(s, 0, f)
        ctx.set_patch_kind(s, 0, capi.BC_ZERO_GRADIENT)
    w = ctx.limiter_weights()
    for it in range(3):
        before = ctx.download(sid[0], 0) @ w
        ctx.euler_limit(*sid)
        after = ctx.download(sid[0], 0)
        assert np.isfinite(after).all() and np.abs(after @ w - before).max() < 1e-12 * np.abs(before).max()
        ctx.euler_stage_fields(*sid, 1.4, 1e-4)
        for s in sid:
            ctx.state_swap(s)
    assert np.isfinite(ctx.download(sid[2], 0)).all()


@pytest.mark.parametrize("kind", [o.BC_ZEROGRAD, o.BC_REFLECTIVE])
def test_frozen_traces_reproduce_the_lagging_boundary_data(gpu_ctx_factory, kind):
    """freeze -> limit -> stage: the stage must see the boundary data of the UNLIMITED field (what the reference's second RK stage sees,
    doubleMach/dgEulerFoam/dgEulerFoam.C:92-107), not the trace of the limited one."""
    wall = kind == o.BC_REFLECTIVE
    mg, om = T._mesh(7, wall)
    N = 3
    case = o.Case(om, N, bc_kinds=[kind])
    ctx = gpu_ctx_factory(N)
    ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], [mg["patch_edges"][0]] if wall else mg["patch_edges"])
    rho, U, E = T._smooth_state(case, amp=0.3)
    bv = T._bvals(case, rho, U, E)                                 # evaluated from the unlimited field
    lim = o.triangle_limit(case, rho, U, E, *bv)
    dt = 1e-3
    stale = o.euler_stage(case, *lim, *[[b[0].copy()] for b in bv], 1.4, dt)
    fresh = o.euler_stage(case, *lim, *T._bvals(case, *lim), 1.4, dt)
    assert H.rel_l2(stale[0], fresh[0]) > 1e-7                     # the two semantics differ measurably on this state
    sid = [ctx.state_create(1), ctx.state_create(2), ctx.state_create(1)]
    for s, f in zip(sid, (rho, U, E)):
        ctx.upload(s, 0, f)
        ctx.set_patch_kind(s, 0, T.KIND[kind])
    for s in sid:
        ctx.freeze_traces(s)
    ctx.euler_limit(*sid)
    ctx.euler_stage_fields(*sid, 1.4, dt)
    for s in sid:
        ctx.state_swap(s)
    got = ctx.download(sid[0], 0), ctx.download(sid[1], 0, 2), ctx.download(sid[2], 0)
    for g, w in zip(got, stale):
        assert H.rel_l2(g, w) <= 1e-12
    # the stage thawed the states: the next one evaluates from the field again
    ctx.euler_stage_fields(*sid, 1.4, dt)
    for s in sid:
        ctx.state_swap(s)
    nxt = o.euler_stage(case, *stale, *T._bvals(case, *stale), 1.4, dt)
    assert H.rel_l2(ctx.download(sid[0], 0), nxt[0]) <= 1e-12
