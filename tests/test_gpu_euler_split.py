"""The split Euler stage (dg_euler_split.cu: ONE Roe flux per dgFace in the owner's orientation, as
defaultConvectionScheme.C:114-127 hands one flux to both cells) against the fused stage kernel (dg_kernels.cu: the flux evaluated on
both sides), through the C ABI on the same inputs - and the dispatch rule between the two.  Both are held to the oracle by
tests/test_gpu_euler_stage.py; here they are held to EACH OTHER at rounding level, for every order that has both and every boundary kind."""
import os

import numpy as np
import pytest

from hopefoam_b200 import capi, meshgen
from oracle import dg_oracle as o
from tests import helpers as H

pytestmark = pytest.mark.gpu
GAMMA = 1.4


def _ctx(N, split):
    """A context whose stage implementation is fixed by HDG_EULER_SPLIT (read once per context, at its first stage)."""
    old = os.environ.get("HDG_EULER_SPLIT")
    os.environ["HDG_EULER_SPLIT"] = "1" if split else "0"
    try:
        c = capi.Context(int(os.environ.get("HDG_TEST_DEVICE", "0")))
        c.set_order(N)
        names = c.euler_stage_kernels()          # forces the switch to be read now
    finally:
        if old is None:
            del os.environ["HDG_EULER_SPLIT"]
        else:
            os.environ["HDG_EULER_SPLIT"] = old
    return c, names


def _mixed_case(N, n=6):
    """bottom side reflective, right side zeroGradient, the other two fixedValue: every exterior-trace source on one mesh"""
    mg = meshgen.jittered_square(n)
    e = mg["patch_edges"][0]
    mg["patch_edges"] = [e[0:n], e[n:2 * n], e[2 * n:4 * n]]
    case = o.Case(H.oracle_mesh(mg), N, bc_kinds=[o.BC_REFLECTIVE, o.BC_ZEROGRAD, o.BC_FIXED])
    return mg, case


def _advance(ctx, mg, case, rk, steps=3, dt=1e-3):
    ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    x, y = case.geo.x[..., 0], case.geo.x[..., 1]
    rho, rhoU, E = H.vortex_state(x, y, 0.3, GAMMA)
    bR, bU, bE = [], [], []
    for ip in range(len(case.mesh.patches)):
        xy = case.patch_internal(case.geo.x, ip)
        r, u, e = H.vortex_state(xy[:, 0], xy[:, 1], 0.31, GAMMA)
        bR.append(r), bU.append(u), bE.append(e)
    sid = H.setup_euler(ctx, case, rho, rhoU, E, bR, bU, bE, case.bc_kinds)
    l0 = ctx.launch_count()
    for _ in range(steps):
        (ctx.euler_step_lserk45 if rk == "lserk45" else ctx.euler_step_ssprk2)(sid, GAMMA, dt)
    ctx.sync()
    launches = ctx.launch_count() - l0
    g = H.download_euler(ctx, sid)
    return (rho, rhoU, E), g, launches


@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
def test_split_equals_fused_all_orders(built_library, N):
    mg, case = _mixed_case(N)
    cs, names_s = _ctx(N, True)
    cf, names_f = _ctx(N, False)
    face = "eulerFacePairFluxKernel" if N <= 2 else "eulerFaceFluxKernel"        # N = 1, 2: two faces per DMMA row
    assert names_s == [f"{face}<{N}>", f"eulerElemKernel<{N}>"] and names_f == [f"eulerStageKernel<{N}>"]
    q0, gs, ls = _advance(cs, mg, case, "ssprk2")
    _, gf, lf = _advance(cf, mg, case, "ssprk2")
    assert (ls, lf) == (3 * 2 * 2, 3 * 2)                    # two launches per stage against one
    for a, b, q in zip(gs, gf, q0):
        assert np.isfinite(a).all()
        assert H.rel_l2(a - q, b - q) <= 2e-12 * max(1, N - 3), (N, H.rel_l2(a - q, b - q))     # the increments, not the fields
    cs.close(), cf.close()


@pytest.mark.parametrize("periodic", [True, False])
def test_split_equals_fused_lserk45(built_library, periodic):
    """The low-storage RK path (residual read and written by the element kernel) and the periodic glue."""
    N = 4
    if periodic:
        mg = meshgen.jittered_square(6, periodic=True)
        case = o.Case(H.oracle_mesh(mg), N)
    else:
        mg, case = _mixed_case(N)
    cs, _ = _ctx(N, True)
    cf, _ = _ctx(N, False)
    q0, gs, _ = _advance(cs, mg, case, "lserk45", steps=2)
    _, gf, _ = _advance(cf, mg, case, "lserk45", steps=2)
    for a, b, q in zip(gs, gf, q0):
        assert H.rel_l2(a - q, b - q) <= 2e-12
    cs.close(), cf.close()


def test_thin_launches_stay_fused(built_library):
    """A launch over at most half of the mesh (the rows next to processor patches) must not pay for a pass over all faces; a step
    assembled from a thin (fused) and a wide (split) launch per stage equals the step of one full launch per stage."""
    N, dt = 4, 1e-3
    mg = meshgen.jittered_square(8, periodic=True)            # 128 triangles = 16 octets
    c, _ = _ctx(N, True)
    c.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    x, y = np.moveaxis(c.node_coords(), -1, 0)
    rho, rhoU, E = H.vortex_state(x, y, 0.3, GAMMA)
    q0 = np.concatenate([rho[..., None], rhoU, E[..., None]], -1)
    s_parts, s_full = c.state_create(4), c.state_create(4)
    c.upload(s_parts, 0, q0)
    c.upload(s_full, 0, q0)
    for stage, (a, b) in enumerate([(0.0, 1.0), (0.5, 0.5)]):
        l0 = c.launch_count()
        c.euler_stage_range(s_parts, GAMMA, dt, stage, a, b, 0, 32)             # 4 octets of 16: fused
        assert c.launch_count() - l0 == 1
        c.euler_stage_range(s_parts, GAMMA, dt, stage, a, b, 32, 128)           # 12 octets of 16: face kernel + element kernel
        assert c.launch_count() - l0 == 3
    c.euler_step_ssprk2(s_full, GAMMA, dt)
    c.sync()
    got, ref = c.download(s_parts, 0, 4), c.download(s_full, 0, 4)
    for f in range(4):
        assert H.rel_l2(got[..., f] - q0[..., f], ref[..., f] - q0[..., f]) <= 2e-12
    c.close()


def test_fields_stage_exchange_flag_without_a_communicator(built_library):
    """hdg_euler_stage_fields_ex(exchange = 1) on an undecomposed mesh (no hdg_comm_init): nothing to exchange, the call is the plain
    stage on the three separate fields - bit-identical to exchange = 0 and equal to the 4-plane stage.  (With a communicator the flag
    hides the halo behind the interior octets: tests/mgpu_facade_parallel.py, two GPUs.)"""
    N, dt = 4, 1e-3
    mg = meshgen.jittered_square(8, periodic=True)
    c, _ = _ctx(N, True)
    c.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    x, y = np.moveaxis(c.node_coords(), -1, 0)
    rho, rhoU, E = H.vortex_state(x, y, 0.3, GAMMA)
    res = []
    for ex in (0, 1):
        s = (c.state_create(1), c.state_create(2), c.state_create(1))
        c.upload(s[0], 0, rho), c.upload(s[1], 0, rhoU), c.upload(s[2], 0, E)
        c.euler_stage_fields_ex(s, GAMMA, dt, exchange=ex)
        for sid in s:
            c.state_swap(sid)
        c.sync()
        res.append((c.download(s[0], 0), c.download(s[1], 0, 2), c.download(s[2], 0)))
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    s4 = c.state_create(4)
    c.upload(s4, 0, np.concatenate([rho[..., None], rhoU, E[..., None]], -1))
    c.euler_stage(s4, GAMMA, dt, 0, 0.0, 1.0)
    c.state_swap(s4)
    c.sync()
    q4 = c.download(s4, 0, 4)
    assert np.array_equal(q4[..., 0], res[0][0]) and np.array_equal(q4[..., 1:3], res[0][1]) and np.array_equal(q4[..., 3], res[0][2])
    c.close()


@pytest.mark.parametrize("N", [2, 4, 7])
@pytest.mark.parametrize("split", [True, False])
def test_lax_friedrichs_flux_option(built_library, N, split):
    """HDG_FLUX_LF on the Euler entry points: point-wise local Lax-Friedrichs (Rusanov) flux - an extension (the reference's godunovScheme
    knows Roe only; oracle.rusanov_flux is a definition, parity unpinned).  Both stage implementations against the oracle, <= 1e-12."""
    mg, case = _mixed_case(N)
    c, _ = _ctx(N, split)
    c.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    x, y = case.geo.x[..., 0], case.geo.x[..., 1]
    rho, rhoU, E = H.vortex_state(x, y, 0.3, GAMMA)
    bR, bU, bE = [], [], []
    for ip in range(len(case.mesh.patches)):
        xy = case.patch_internal(case.geo.x, ip)
        r, u, e = H.vortex_state(xy[:, 0], xy[:, 1], 0.31, GAMMA)
        bR.append(r), bU.append(u), bE.append(e)
    case.evaluate_bc(rho, bR)
    case.evaluate_bc(rhoU, bU, is_vector=True)
    case.evaluate_bc(E, bE)
    sid = H.setup_euler(c, case, rho, rhoU, E, bR, bU, bE, case.bc_kinds)
    dt = 1e-3
    r1, u1, e1 = o.euler_stage(case, rho, rhoU, E, bR, bU, bE, GAMMA, dt, flux="LF")
    r2, u2, e2 = o.euler_stage(case, r1, u1, e1, bR, bU, bE, GAMMA, dt, flux="LF")
    c.euler_step_ssprk2(sid, GAMMA, dt, flux=capi.FLUX_LF)
    c.sync()
    got = H.download_euler(c, sid)
    ref = (0.5 * rho + 0.5 * r2, 0.5 * rhoU + 0.5 * u2, 0.5 * E + 0.5 * e2)
    for a, b, q in zip(got, ref, (rho, rhoU, E)):
        assert H.rel_l2(a, b) <= 1e-12
        assert H.rel_l2(a - q, b - q) <= 1e-10          # the O(dt) increment itself
    # and it IS a different flux: the Roe result differs measurably
    s2 = H.setup_euler(c, case, rho, rhoU, E, bR, bU, bE, case.bc_kinds)
    c.euler_step_ssprk2(s2, GAMMA, dt)
    c.sync()
    assert H.rel_l2(H.download_euler(c, s2)[0] - rho, got[0] - rho) > 1e-4
    # any other scheme is refused
    with pytest.raises(capi.HdgError):
        c.euler_step_ssprk2(sid, GAMMA, dt, flux=capi.FLUX_AVERAGE)
    c.close()
