"""Longer GPU runs through the C ABI: 1000 SSP-RK2 steps against the oracle (1e-10, north_star), the golden vortex
error norms on the tutorial mesh shape (when the fixture is available), LSERK45, the L1 error reduction."""
import numpy as np
import pytest

from hopefoam_b200 import capi, meshgen
from oracle import dg_oracle as o
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _vortex_run(ctx, N, n, dt, periodic=False):
    mg = meshgen.jittered_square(n, periodic=periodic)
    case = o.Case(H.oracle_mesh(mg), N)
    ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    run = o.VortexRun(case, dt)
    run.set_boundary_values(0.0)
    sid = H.setup_euler(ctx, case, run.rho, run.rhoU, run.E, run.bR, run.bU, run.bE, case.bc_kinds)
    return case, run, sid


def test_euler_1000_steps_vs_oracle(gpu_ctx_factory):
    ctx = gpu_ctx_factory(4)
    case, run, sid = _vortex_run(ctx, 4, 6, 1e-3)
    for _ in range(1000):
        # solver hook setBoundaryValues(t_n) every step (dgEulerFoam.C:73), through the C ABI like the facade does
        run.set_boundary_values(run.t)
        for ip in range(len(case.mesh.patches)):
            ctx.set_patch_values(sid, 0, ip, run.bR[ip])
            ctx.set_patch_values(sid, 1, ip, run.bU[ip])
            ctx.set_patch_values(sid, 3, ip, run.bE[ip])
        ctx.euler_step_ssprk2(sid, 1.4, run.dt)
        run.step()
    ctx.sync()
    got = H.download_euler(ctx, sid)
    errs = [H.rel_l2(a, b) for a, b in zip(got, (run.rho, run.rhoU, run.E))]
    assert max(errs) <= 1e-10, errs
    ctx.close()


def test_l1_diff_matches_numpy(gpu_ctx_factory):
    ctx = gpu_ctx_factory(3)
    case, run, sid = _vortex_run(ctx, 3, 5, 1e-3)
    ref = run.rho + 1e-3 * np.sin(case.geo.x[..., 0])
    got = ctx.l1_diff(sid, 0, ref)
    want = np.abs(run.rho - ref).sum()
    assert abs(got - want) <= 1e-12 * want
    ctx.close()


def test_lserk45_matches_unfused_reference(gpu_ctx_factory):
    """LSERK(5,4) with the coefficients of createFields.H:119-131: compare with the same scheme driven stage by stage on
    the oracle's RHS (L = (forward-Euler result - q)/dt)."""
    ctx = gpu_ctx_factory(4)
    case, run, sid = _vortex_run(ctx, 4, 5, 2e-3, periodic=True)
    a = [0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0, -3550918686646.0 / 2091501179385.0,
         -1275806237668.0 / 842570457699.0]
    b = [1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0, 1720146321549.0 / 2090206949498.0,
         3134564353537.0 / 4481467310338.0, 2277821191437.0 / 14882151754819.0]
    q = [run.rho.copy(), run.rhoU.copy(), run.E.copy()]
    dt = run.dt
    for _ in range(3):
        res = [np.zeros_like(x) for x in q]
        for s in range(5):
            q1 = o.euler_stage(case, q[0], q[1], q[2], [], [], [], 1.4, dt)
            for i in range(3):
                res[i] = a[s] * res[i] + (q1[i] - q[i])         # dt*L(q)
                q[i] = q[i] + b[s] * res[i]
        ctx.euler_step_lserk45(sid, 1.4, dt)
    ctx.sync()
    got = H.download_euler(ctx, sid)
    errs = [H.rel_l2(x, y) for x, y in zip(got, q)]
    assert max(errs) <= 1e-11, errs
    ctx.close()


@pytest.mark.parametrize("n", [160, 707])
def test_large_mesh_properties(gpu_ctx_factory, n):
    """BASELINE-size check through size-independent properties (n=707: the 999 698-triangle benchmark mesh): on a periodic
    mesh the scheme conserves mass, momentum and energy to round-off (sum_k J_k w^T V q), and a uniform state is a fixed point."""
    ctx = gpu_ctx_factory(4)
    mg = meshgen.jittered_square(n, periodic=True)
    ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    xy = ctx.node_coords()
    r, u, e = H.vortex_state(xy[..., 0], xy[..., 1])
    sid = ctx.state_create(4)
    ctx.upload(sid, 0, r); ctx.upload(sid, 1, u); ctx.upload(sid, 3, e)
    ref = o.RefElement(4)
    wnode = ref.Vg.T @ ref.gw                                  # nodal quadrature weights of the reference element
    v = mg["xy"][ctx.cell_vertices()]
    J = 0.25 * ((v[:, 1, 0] - v[:, 0, 0]) * (v[:, 2, 1] - v[:, 0, 1]) - (v[:, 1, 1] - v[:, 0, 1]) * (v[:, 2, 0] - v[:, 0, 0]))
    total = lambda q: float(((q @ wnode) * J).sum())
    before = [total(r), total(u[..., 0]), total(u[..., 1]), total(e)]
    dt = 0.09 / n                                             # stable step for N=4 (User Guide Table 1.1 scaled with h)
    for _ in range(20):
        ctx.euler_step_ssprk2(sid, 1.4, dt)
    ctx.sync()
    r2, u2, e2 = H.download_euler(ctx, sid)
    after = [total(r2), total(u2[..., 0]), total(u2[..., 1]), total(e2)]
    for a_, b_ in zip(after, before):
        assert abs(a_ - b_) <= 1e-11 * max(1.0, abs(b_)), (after, before)
    assert np.isfinite(r2).all() and r2.min() > 0
    # uniform state is preserved exactly up to round-off
    ctx.upload(sid, 0, np.full_like(r, 1.3)); ctx.upload(sid, 1, np.stack([np.full_like(r, 0.4), np.full_like(r, -0.2)], -1))
    ctx.upload(sid, 3, np.full_like(r, 2.5))
    ctx.euler_step_ssprk2(sid, 1.4, dt)
    ctx.sync()
    r3, u3, e3 = H.download_euler(ctx, sid)
    assert np.abs(r3 - 1.3).max() < 1e-12 and np.abs(u3[..., 0] - 0.4).max() < 1e-12 and np.abs(e3 - 2.5).max() < 1e-12
    ctx.close()


def test_published_vortex_errors_on_gpu(gpu_ctx_factory):
    """The CUDA path itself against the reference's PUBLISHED numbers (User Guide §1.8: N=4, vortex1024.msh, dt=0.004, t=2,
    exact-solution fixedValue boundary refreshed at t_n every step): rhoError 8.807979526797244e-06, rhoUError 1.865574862711117e-05."""
    import json
    from pathlib import Path
    gold = Path(__file__).resolve().parent / "golden"
    g = json.loads((gold / "golden_errors.json").read_text())["user_guide"]
    d = np.load(gold / f"{g['mesh']}.npz")
    ctx = gpu_ctx_factory(g["N"])
    ctx.set_mesh_triangles(d["xy"], d["tris"], None, [d["patch_edges"]])
    xy, pxy = ctx.node_coords(), ctx.patch_node_coords(0)
    r, u, e = H.vortex_state(xy[..., 0], xy[..., 1], 0.0)
    sid = ctx.state_create(4)
    ctx.upload(sid, 0, r); ctx.upload(sid, 1, u); ctx.upload(sid, 3, e)
    dt, t = g["dt"], 0.0
    for _ in range(int(round(g["endTime"] / dt))):
        br, bu, be = H.vortex_state(pxy[:, 0], pxy[:, 1], t)
        ctx.set_patch_values(sid, 0, 0, br); ctx.set_patch_values(sid, 1, 0, bu); ctx.set_patch_values(sid, 3, 0, be)
        ctx.euler_step_ssprk2(sid, 1.4, dt)
        t += dt
    ctx.sync()
    rx, ux, _ = H.vortex_state(xy[..., 0], xy[..., 1], t)
    err_rho = ctx.l1_diff(sid, 0, rx) / rx.size                    # eulerError.H:32-38 through the C ABI
    rho, rhoU, _ = H.download_euler(ctx, sid)
    err_rhou = np.sqrt(((rhoU - ux) ** 2).sum(-1)).sum() / rx.size
    assert abs(err_rho - g["rhoError"]) <= 1e-9 * g["rhoError"], err_rho
    assert abs(err_rhou - g["rhoUError"]) <= 1e-9 * g["rhoUError"], err_rhou
    assert abs(np.abs(rho - rx).sum() / rx.size - err_rho) <= 1e-12 * err_rho
    ctx.close()


def _published_table():
    import json
    from pathlib import Path
    T = json.loads((Path(__file__).resolve().parent / "golden" / "golden_errors.json").read_text())["slide18_full"]
    return [(N, mi, mesh) for N in range(1, 7) for mi, mesh in enumerate(T["meshes"])], T


@pytest.mark.parametrize("N,mi,mesh", _published_table()[0], ids=lambda v: str(v))
def test_published_convergence_table_on_gpu(gpu_ctx_factory, N, mi, mesh):
    """All 36 numbers of the reference's published accuracy table (workshop deck slide 18: rho and rhoU error at t=2, N=1..6 on
    vortex0256/1024/4096, dt of User Guide Table 1.1) reproduced by the CUDA path to the 4 printed digits."""
    from pathlib import Path
    T = _published_table()[1]
    d = np.load(Path(__file__).resolve().parent / "golden" / f"{mesh}.npz")
    ctx = gpu_ctx_factory(N)
    ctx.set_mesh_triangles(d["xy"], d["tris"], None, [d["patch_edges"]])
    xy, pxy = ctx.node_coords(), ctx.patch_node_coords(0)
    r, u, e = H.vortex_state(xy[..., 0], xy[..., 1], 0.0)
    sid = ctx.state_create(4)
    ctx.upload(sid, 0, r); ctx.upload(sid, 1, u); ctx.upload(sid, 3, e)
    dt, t = T["dt"][str(N)][mi], 0.0
    for _ in range(int(round(2.0 / dt))):
        br, bu, be = H.vortex_state(pxy[:, 0], pxy[:, 1], t)
        ctx.set_patch_values(sid, 0, 0, br); ctx.set_patch_values(sid, 1, 0, bu); ctx.set_patch_values(sid, 3, 0, be)
        ctx.euler_step_ssprk2(sid, 1.4, dt)
        t += dt
    ctx.sync()
    rx, ux, _ = H.vortex_state(xy[..., 0], xy[..., 1], t)
    rho, rhoU, _ = H.download_euler(ctx, sid)
    er = np.abs(rho - rx).sum() / rx.size
    eu = np.sqrt(((rhoU - ux) ** 2).sum(-1)).sum() / rx.size
    pr, pu = T["rho"][str(N)][mi], T["rhoU"][str(N)][mi]
    assert abs(er - pr) <= 6e-4 * pr, (er, pr)                       # 4 significant digits printed
    assert abs(eu - pu) <= 6e-4 * pu, (eu, pu)
    ctx.close()
