"""GPU parity of the fused scalar-advection stage (dgc::div(U,T) + LF flux) against the oracle."""
import numpy as np
import pytest

from hopefoam_b200 import capi, meshgen
from oracle import dg_oracle as o
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _setup(ctx, N, n, periodic, uniform=True, kinds=None):
    mg = meshgen.jittered_square(n, x0=-1, x1=1, y0=-1, y1=1, periodic=periodic)
    case = o.Case(H.oracle_mesh(mg), N, bc_kinds=kinds)
    ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    x, y = case.geo.x[..., 0], case.geo.x[..., 1]
    T = np.exp(-((x + 0.3) ** 2 + (y + 0.3) ** 2) / (2 * 0.1 ** 2)) + 0.1 * np.sin(3 * x) * np.cos(2 * y)
    if uniform:
        Ux, Uy = np.full_like(x, 1.0), np.full_like(x, 0.5)
    else:
        Ux, Uy = 1.0 + 0.3 * y, 0.5 - 0.2 * x        # linear (divergence-free) nodal velocity
    npatch = len(case.mesh.patches)
    bT, bUx, bUy = [], [], []
    for ip in range(npatch):
        xy = case.patch_internal(case.geo.x, ip)
        bT.append(np.exp(-((xy[:, 0] + 0.29) ** 2 + (xy[:, 1] + 0.295) ** 2) / (2 * 0.1 ** 2)))
        bUx.append(case.patch_internal(Ux, ip).copy())
        bUy.append(case.patch_internal(Uy, ip).copy())
    sT, sU = ctx.state_create(1), ctx.state_create(2)
    ctx.upload(sT, 0, T)
    ctx.upload(sU, 0, np.stack([Ux, Uy], -1))
    for ip in range(npatch):
        kind = case.bc_kinds[ip]
        ck = {o.BC_FIXED: capi.BC_FIXED_VALUE, o.BC_ZEROGRAD: capi.BC_ZERO_GRADIENT}[kind]
        ctx.set_patch_kind(sT, ip, ck)
        ctx.set_patch_kind(sU, ip, capi.BC_FIXED_VALUE)
        if kind == o.BC_FIXED:
            ctx.set_patch_values(sT, 0, ip, bT[ip])
        ctx.set_patch_values(sU, 0, ip, np.stack([bUx[ip], bUy[ip]], -1))
    return case, (T, Ux, Uy, bT, bUx, bUy), (sT, sU)


@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
@pytest.mark.parametrize("uniform", [True, False])
def test_advect_ssprk2_fixed_value(gpu_ctx_factory, N, uniform):
    ctx = gpu_ctx_factory(N)
    case, (T, Ux, Uy, bT, bUx, bUy), (sT, sU) = _setup(ctx, N, 5, False, uniform)
    dt = 2e-3
    T1 = o.advect_stage(case, T, Ux, Uy, bT, bUx, bUy, dt)
    T2 = o.advect_stage(case, T1, Ux, Uy, bT, bUx, bUy, dt)
    ref = 0.5 * T + 0.5 * T2
    ctx.advect_step_ssprk2(sT, sU, dt, capi.FLUX_LF)
    ctx.sync()
    got = ctx.download(sT, 0)
    assert H.rel_l2(got, ref) <= 1e-12
    assert H.rel_l2(got - T, ref - T) <= 1e-11      # the increment (what the kernel computes), not only the field
    ctx.close()


@pytest.mark.parametrize("N", [4, 5, 6])
@pytest.mark.parametrize("periodic,kinds", [(True, None), (False, [o.BC_ZEROGRAD])])
def test_advect_periodic_and_zero_gradient(gpu_ctx_factory, periodic, kinds, N):
    ctx = gpu_ctx_factory(N)
    case, (T, Ux, Uy, bT, bUx, bUy), (sT, sU) = _setup(ctx, N, 6, periodic, False, kinds)
    if kinds:
        case.evaluate_bc(T, bT)
    dt = 1e-3
    T1 = o.advect_stage(case, T, Ux, Uy, bT, bUx, bUy, dt)
    ctx.advect_stage(sT, sU, dt, 0, 0.0, 1.0, capi.FLUX_LF)
    T2 = o.advect_stage(case, T1, Ux, Uy, bT, bUx, bUy, dt)
    ctx.advect_stage(sT, sU, dt, 1, 0.5, 0.5, capi.FLUX_LF)
    ctx.sync()
    assert H.rel_l2(ctx.download(sT, 0), 0.5 * T + 0.5 * T2) <= 1e-12
    ctx.close()


def test_advect_1000_steps_gaussian(gpu_ctx_factory):
    """Config-1 style run (Gaussian pulse, uniform U, LF, N=4, SSP-RK2): 1e-10 after 1000 steps vs the oracle."""
    ctx = gpu_ctx_factory(4)
    case, (T, Ux, Uy, bT, bUx, bUy), (sT, sU) = _setup(ctx, 4, 8, True, True)
    dt = 1e-3
    Tn = T
    for _ in range(1000):
        T1 = o.advect_stage(case, Tn, Ux, Uy, bT, bUx, bUy, dt)
        T2 = o.advect_stage(case, T1, Ux, Uy, bT, bUx, bUy, dt)
        Tn = 0.5 * Tn + 0.5 * T2
        ctx.advect_step_ssprk2(sT, sU, dt, capi.FLUX_LF)
    ctx.sync()
    assert H.rel_l2(ctx.download(sT, 0), Tn) <= 1e-10
    ctx.close()


def test_advect_config1_gaussian_fixed_value(gpu_ctx_factory):
    """BASELINE configs[0]: Gaussian pulse, uniform U=(1,0.5), 71x71x2 = 10 082 jittered triangles on [-1,1]^2, one fixedValue
    patch holding the translated exact Gaussian (refreshed every step at t_n like setBoundaryValues), LF, N=4, SSP-RK2,
    dt = 0.1 h_min/(|U|(N+1)^2), the 1000 steps SURVEY §8-d prescribes: <= 1e-12 after 40 steps, <= 1e-10 after 1000 (north_star)."""
    N, n = 4, 71
    ctx = gpu_ctx_factory(N)
    mg = meshgen.jittered_square(n, x0=-1, x1=1, y0=-1, y1=1)
    assert mg["tris"].shape[0] == 10082
    case = o.Case(H.oracle_mesh(mg), N)
    ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    x, y = case.geo.x[..., 0], case.geo.x[..., 1]
    ux, uy = 1.0, 0.5
    exact = lambda xx, yy, t: np.exp(-((xx + 0.3 - ux * t) ** 2 + (yy + 0.3 - uy * t) ** 2) / (2 * 0.1 ** 2))
    T = exact(x, y, 0.0)
    Ux, Uy = np.full_like(x, ux), np.full_like(x, uy)
    h = 2.0 / n
    dt = 0.1 * (0.6 * h) / (np.hypot(ux, uy) * (N + 1) ** 2)
    pxy = case.patch_internal(case.geo.x, 0)
    bUx, bUy = [np.full(pxy.shape[0], ux)], [np.full(pxy.shape[0], uy)]
    sT, sU = ctx.state_create(1), ctx.state_create(2)
    ctx.upload(sT, 0, T)
    ctx.upload(sU, 0, np.stack([Ux, Uy], -1))
    ctx.set_patch_values(sU, 0, 0, np.stack([bUx[0], bUy[0]], -1))
    Tn, t = T, 0.0
    for step in range(1000):
        if step == 40:
            ctx.sync()
            assert H.rel_l2(ctx.download(sT, 0), Tn) <= 1e-12
        bT = [exact(pxy[:, 0], pxy[:, 1], t)]
        ctx.set_patch_values(sT, 0, 0, bT[0])
        T1 = o.advect_stage(case, Tn, Ux, Uy, bT, bUx, bUy, dt)
        T2 = o.advect_stage(case, T1, Ux, Uy, bT, bUx, bUy, dt)
        Tn = 0.5 * Tn + 0.5 * T2
        ctx.advect_step_ssprk2(sT, sU, dt, capi.FLUX_LF)
        t += dt
    ctx.sync()
    got = ctx.download(sT, 0)
    err = H.rel_l2(got, Tn)
    print("config 1, 1000 steps: rel-L2 vs oracle", err)
    assert err <= 1e-10
    assert np.abs(got - exact(x, y, t)).max() < 5e-3          # and it is actually advecting the pulse
    ctx.close()


@pytest.mark.parametrize("N", [3, 5, 6])
@pytest.mark.parametrize("kind,flux", [("average", capi.FLUX_AVERAGE), ("none", capi.FLUX_NONE)])
def test_advect_average_and_none_flux(gpu_ctx_factory, kind, flux, N):
    """The other run-time selectable fluxCalcSchemes of the reference: `average` (averageFlux.C:95-190) and `none` (noneFlux.C:45-97)."""
    ctx = gpu_ctx_factory(N)
    case, (T, Ux, Uy, bT, bUx, bUy), (sT, sU) = _setup(ctx, N, 5, False, False)
    dt = 1e-3
    T1 = o.advect_stage(case, T, Ux, Uy, bT, bUx, bUy, dt, kind)
    T2 = o.advect_stage(case, T1, Ux, Uy, bT, bUx, bUy, dt, kind)
    ctx.advect_step_ssprk2(sT, sU, dt, flux)
    ctx.sync()
    assert H.rel_l2(ctx.download(sT, 0), 0.5 * T + 0.5 * T2) <= 1e-12
    ctx.close()


# LSERK(5,4) coefficients as declared in TUT/isentropicVortex/dgEulerFoam/createFields.H:119-131
RK4A = [0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0, -3550918686646.0 / 2091501179385.0,
        -1275806237668.0 / 842570457699.0]
RK4B = [1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0, 1720146321549.0 / 2090206949498.0,
        3134564353537.0 / 4481467310338.0, 2277821191437.0 / 14882151754819.0]


@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 6, 7])
def test_advect_lserk45(gpu_ctx_factory, N):
    """Low-storage RK(5,4) driver on the fused advection stage (residual read + written every stage), two steps vs the oracle's
    operator (L(T) recovered from its forward-Euler stage)."""
    ctx = gpu_ctx_factory(N)
    case, (T, Ux, Uy, bT, bUx, bUy), (sT, sU) = _setup(ctx, N, 6, False, False)
    dt = 2e-3
    Tn, res = T.copy(), np.zeros_like(T)
    for _ in range(2):
        for a, b in zip(RK4A, RK4B):
            L = (o.advect_stage(case, Tn, Ux, Uy, bT, bUx, bUy, 1.0) - Tn)
            res = a * res + dt * L
            Tn = Tn + b * res
        ctx.advect_step_lserk45(sT, sU, dt, capi.FLUX_LF)
    ctx.sync()
    got = ctx.download(sT, 0)
    assert H.rel_l2(got, Tn) <= 1e-12
    assert H.rel_l2(got - T, Tn - T) <= 1e-11
    ctx.close()


@pytest.mark.parametrize("N", [1, 2, 4, 5, 6, 7])
@pytest.mark.parametrize("n", [3, 5, 9])
def test_advect_ragged_octets(gpu_ctx_factory, n, N):
    """Element counts that are not a multiple of 8 (18, 50, 162 triangles: the last octet is ragged, its padding rows must stay
    zero); N = 4 (128-B rows), and the wide-row kernel: 1, 2 (64-B rows), 5 (192-B rows, no swizzle), 6 (256-B rows as two swizzled
    lines), 7 (320-B rows)."""
    ctx = gpu_ctx_factory(N)
    case, (T, Ux, Uy, bT, bUx, bUy), (sT, sU) = _setup(ctx, N, n, False, False)
    assert case.mesh.K % 8 != 0
    dt = 1e-3
    Tn = T
    for _ in range(3):
        T1 = o.advect_stage(case, Tn, Ux, Uy, bT, bUx, bUy, dt)
        T2 = o.advect_stage(case, T1, Ux, Uy, bT, bUx, bUy, dt)
        Tn = 0.5 * Tn + 0.5 * T2
        ctx.advect_step_ssprk2(sT, sU, dt, capi.FLUX_LF)
    ctx.sync()
    assert H.rel_l2(ctx.download(sT, 0), Tn) <= 1e-12
    ctx.close()


def test_advect_velocity_changes_between_steps(gpu_ctx_factory):
    """The TMA kernel reads the velocity from a derived (x,y)-pair copy kept by the library; every write to the velocity state
    (upload, boundary values) must be seen by the next stage."""
    ctx = gpu_ctx_factory(4)
    case, (T, Ux, Uy, bT, bUx, bUy), (sT, sU) = _setup(ctx, 4, 6, False, False)
    dt = 1e-3
    npatch = len(case.mesh.patches)

    def oracle_step(Tn, ux, uy, bux, buy):
        T1 = o.advect_stage(case, Tn, ux, uy, bT, bux, buy, dt)
        T2 = o.advect_stage(case, T1, ux, uy, bT, bux, buy, dt)
        return 0.5 * Tn + 0.5 * T2

    Tn = oracle_step(T, Ux, Uy, bUx, bUy)
    ctx.advect_step_ssprk2(sT, sU, dt, capi.FLUX_LF)
    # new interior velocity, same boundary data
    Ux2, Uy2 = 0.7 * Ux - 0.1, 1.3 * Uy + 0.2
    ctx.upload(sU, 0, np.stack([Ux2, Uy2], -1))
    Tn = oracle_step(Tn, Ux2, Uy2, bUx, bUy)
    ctx.advect_step_ssprk2(sT, sU, dt, capi.FLUX_LF)
    # new boundary velocity only
    bUx3 = [0.5 * b for b in bUx]
    bUy3 = [b + 0.3 for b in bUy]
    for ip in range(npatch):
        ctx.set_patch_values(sU, 0, ip, np.stack([bUx3[ip], bUy3[ip]], -1))
    Tn = oracle_step(Tn, Ux2, Uy2, bUx3, bUy3)
    ctx.advect_step_ssprk2(sT, sU, dt, capi.FLUX_LF)
    ctx.sync()
    assert H.rel_l2(ctx.download(sT, 0), Tn) <= 1e-12
    ctx.close()


@pytest.mark.parametrize("n", [1, 2])
def test_advect_smallest_meshes(gpu_ctx_factory, n):
    """2 and 8 triangles: one (ragged / exactly full) octet, every element touches the boundary."""
    ctx = gpu_ctx_factory(4)
    case, (T, Ux, Uy, bT, bUx, bUy), (sT, sU) = _setup(ctx, 4, n, False, False)
    dt = 1e-3
    T1 = o.advect_stage(case, T, Ux, Uy, bT, bUx, bUy, dt)
    T2 = o.advect_stage(case, T1, Ux, Uy, bT, bUx, bUy, dt)
    ctx.advect_step_ssprk2(sT, sU, dt, capi.FLUX_LF)
    ctx.sync()
    assert H.rel_l2(ctx.download(sT, 0), 0.5 * T + 0.5 * T2) <= 1e-12
    ctx.close()


@pytest.mark.parametrize("cfg", ["0", "2", "3"])
def test_alternate_kernel_configurations(cfg):
    """HDG_ADV_CFG selects the data path of the advection stage when the library is loaded: 0 = the first kernel for every order,
    2 = TMA pipeline with the result leaving through a shared-memory tile + TMA store, 3 = three blocks of 4 warps per SM instead of
    one block of 12 (N = 3, 4).  All must pass the same parity tests (run in a child process, because the choice is latched on
    first use)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, HDG_ADV_CFG=cfg)
    sel = "(periodic_and_zero or ragged or lserk or fixed_value or average) and not config1"
    out = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-x", "-m", "gpu", "-k", sel], env=env, capture_output=True, text=True,
                         timeout=600, cwd=str(__import__("pathlib").Path(__file__).resolve().parent.parent))
    assert out.returncode == 0, out.stdout[-3000:]
    assert " passed" in out.stdout


@pytest.mark.parametrize("cfg", ["1", "2", "3"])
def test_wide_kernel_configurations(cfg):
    """HDG_ADVW_CFG selects stages / warps per block / resident blocks of the wide-row TMA kernel (N = 1, 2, 5, 6, 7); every
    configuration passes the same parity tests (child process: the choice is latched on first use)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, HDG_ADVW_CFG=cfg)
    sel = "ragged or lserk or periodic_and_zero"
    out = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-x", "-m", "gpu", "-k", sel], env=env, capture_output=True, text=True,
                         timeout=600, cwd=str(__import__("pathlib").Path(__file__).resolve().parent.parent))
    assert out.returncode == 0, out.stdout[-3000:]
    assert " passed" in out.stdout


@pytest.mark.parametrize("n,N", [(160, 4), (707, 4), (160, 2), (160, 5), (160, 6), (500, 5)])
def test_advect_large_mesh_properties(gpu_ctx_factory, n, N):
    """BASELINE-size check of the advection stage through size-independent properties (n=707: 999 698 triangles, the bench mesh; the
    wide-row kernel of N = 2, 5, 6 on 51 200 and 500 000 triangles): on a periodic mesh with a divergence-free nodal velocity the
    total of T (sum_k J_k w^T V T) is conserved to round-off by the LF flux form, a constant T stays constant, and the LSERK(5,4) and
    SSP-RK2 drivers agree to their truncation error."""
    ctx = gpu_ctx_factory(N)
    mg = meshgen.jittered_square(n, x0=-1, x1=1, y0=-1, y1=1, periodic=True)
    ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    xy = ctx.node_coords()
    x, y = xy[..., 0], xy[..., 1]
    T0 = np.exp(-((x + 0.3) ** 2 + (y + 0.3) ** 2) / (2 * 0.1 ** 2)) + 0.2
    U = np.stack([np.full_like(x, 1.0), np.full_like(x, 0.5)], -1)
    sT, sU = ctx.state_create(1), ctx.state_create(2)
    ctx.upload(sT, 0, T0)
    ctx.upload(sU, 0, U)
    ref = o.RefElement(N)
    wnode = ref.Vg.T @ ref.gw
    v = mg["xy"][ctx.cell_vertices()]
    J = 0.25 * ((v[:, 1, 0] - v[:, 0, 0]) * (v[:, 2, 1] - v[:, 0, 1]) - (v[:, 1, 1] - v[:, 0, 1]) * (v[:, 2, 0] - v[:, 0, 0]))
    total = lambda q: float(((q @ wnode) * J).sum())
    dt = 0.04 / n * (5.0 / (N + 1)) ** 2
    for _ in range(20):
        ctx.advect_step_ssprk2(sT, sU, dt, capi.FLUX_LF)
    ctx.sync()
    T1 = ctx.download(sT, 0)
    assert np.isfinite(T1).all()
    assert abs(total(T1) - total(T0)) <= 1e-12 * abs(total(T0)), (total(T1), total(T0))
    assert 0.1 < T1.min() and T1.max() < 1.3
    # the same 20 steps with the low-storage RK(5,4): a different time integrator, same operator -> close, not equal
    ctx.upload(sT, 0, T0)
    for _ in range(20):
        ctx.advect_step_lserk45(sT, sU, dt, capi.FLUX_LF)
    ctx.sync()
    T2 = ctx.download(sT, 0)
    assert abs(total(T2) - total(T0)) <= 1e-12 * abs(total(T0))
    assert H.rel_l2(T2, T1) < 1e-4
    # constant field: fixed point
    ctx.upload(sT, 0, np.full_like(T0, 0.7))
    ctx.advect_step_ssprk2(sT, sU, dt, capi.FLUX_LF)
    ctx.sync()
    assert np.abs(ctx.download(sT, 0) - 0.7).max() < 1e-13
    ctx.close()
