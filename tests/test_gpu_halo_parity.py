"""Processor-patch halo parity on ONE GPU, through the C ABI (replaces processorDgPatchField::initEvaluate/evaluate,
processorDgPatchField.C:235-331).

The processors of a decomposition (hdg_decompose_simple / a manual cellToProc + hdg_mesh_decompose = the repo's dgDecomposePar path) are n
contexts of this process on device 0; hdg_group_euler_step_ssprk2 advances them with the library's overlapped exchange (octets that own a
processor face -> pack kernel -> peer copy -> unpack kernel under the launch over the other octets).  It is the code path of the one-
process-per-GPU step hdg_euler_step_ssprk2_parallel except for the transport (cudaMemcpyPeerAsync instead of ncclSend/ncclRecv).
The result, put back through cellProcAddressing, is held to the ORACLE on the undecomposed mesh (<= 1e-12 per step) and to one context
on the undecomposed mesh (pure data movement: <= 1e-14)."""
import os

import numpy as np
import pytest

from hopefoam_b200 import capi, meshgen
from oracle import dg_oracle as o
from tests import helpers as H

pytestmark = pytest.mark.gpu

GAMMA = 1.4


def _global_ctx(N, mg):
    g = capi.Context(0)
    g.set_order(N)
    g.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    return g


def _processors(g, N, c2p, nproc):
    ctxs = []
    for r in range(nproc):
        c = capi.Context(0)
        c.set_order(N)
        c.set_mesh_from_decomposition(g, c2p, nproc, r)
        ctxs.append(c)
    return ctxs


def _setup_state(c, q, kinds, patch_values):
    """q: (K,Np,4); kinds: per ORIGINAL patch capi.BC_*; patch_values(ctx, patch) -> (n,4) for fixedValue patches."""
    sid = c.state_create(4)
    c.upload(sid, 0, q)
    nbr = c.proc_addressing()["patch_nbr_proc"]
    for p in range(c.n_patches):
        if nbr[p] >= 0:
            c.set_patch_kind(sid, p, capi.BC_PROCESSOR)
            continue
        c.set_patch_kind(sid, p, kinds[p])
        if kinds[p] == capi.BC_FIXED_VALUE and c.patch_info(p)[2]:
            c.set_patch_values(sid, 0, p, patch_values(c, p))
    return sid


def _gather(ctxs, sids, K, Np):
    full = np.zeros((K, Np, 4))
    for c, s in zip(ctxs, sids):
        full[c.proc_addressing()["cell"]] = c.download(s, 0, 4)
    return full


def _vortex4(xy, t=0.0):
    r, u, e = H.vortex_state(xy[..., 0], xy[..., 1], t, GAMMA)
    return np.concatenate([r[..., None], u, e[..., None]], -1)


def _oracle_steps(case, q0, bvals, dt, steps):
    rho, rhoU, E = q0[..., 0].copy(), q0[..., 1:3].copy(), q0[..., 3].copy()
    bR, bU, bE = bvals
    case.evaluate_bc(rho, bR)
    case.evaluate_bc(rhoU, bU, is_vector=True)
    case.evaluate_bc(E, bE)
    for _ in range(steps):
        r1, u1, e1 = o.euler_stage(case, rho, rhoU, E, bR, bU, bE, GAMMA, dt)
        r2, u2, e2 = o.euler_stage(case, r1, u1, e1, bR, bU, bE, GAMMA, dt)
        rho, rhoU, E = 0.5 * rho + 0.5 * r2, 0.5 * rhoU + 0.5 * u2, 0.5 * E + 0.5 * e2
        case.evaluate_bc(rho, bR)
        case.evaluate_bc(rhoU, bU, is_vector=True)
        case.evaluate_bc(E, bE)
    return np.concatenate([rho[..., None], rhoU, E[..., None]], -1)


def _stage_launches(pc):
    """Kernel launches of one parallel stage: boundary octets, pack, unpack, interior octets.  A launch over more than half of the mesh
    runs as face-flux kernel + element kernel (the split stage, dg_euler_split.cu), the thin one as one fused kernel."""
    nb, ni = pc["boundary_octets"], pc["interior_octets"]
    split_on = os.environ.get("HDG_EULER_SPLIT", "1") != "0"
    stage = lambda n: 0 if n == 0 else (2 if split_on and 2 * n > nb + ni else 1)
    return stage(nb) + 2 + stage(ni)


def _run_case(N, mg, div, kinds_oracle, steps=3, dt=1e-3, c2p=None, nproc=None):
    """Decompose `mg`, advance `steps` SSP-RK2 steps on the processors and on the undecomposed mesh; returns the three fields."""
    kind_map = {o.BC_FIXED: capi.BC_FIXED_VALUE, o.BC_ZEROGRAD: capi.BC_ZERO_GRADIENT, o.BC_REFLECTIVE: capi.BC_REFLECTIVE}
    g = _global_ctx(N, mg)
    if c2p is None:
        c2p = g.decompose_simple(*div)
        nproc = int(np.prod(div))
    kinds = [kind_map[k] for k in kinds_oracle]
    bval = lambda c, p: _vortex4(c.patch_node_coords(p), 0.01)            # boundary data differs from the trace
    ctxs = _processors(g, N, c2p, nproc)
    sids = [_setup_state(c, _vortex4(c.node_coords()), kinds, bval) for c in ctxs]
    l0 = [c.launch_count() for c in ctxs]
    for _ in range(steps):
        capi.group_euler_step_ssprk2(ctxs, sids, GAMMA, dt)
    for c in ctxs:
        c.sync()
    # per step: 2 stages x (boundary launch + pack + unpack + interior launch) + one priming exchange (pack + unpack) at the first step
    for c, l in zip(ctxs, l0):
        pc = c.par_counts()
        assert pc["neighbours"] >= 1 and pc["proc_faces"] > 0 and pc["boundary_octets"] > 0
        assert c.launch_count() - l == steps * 2 * _stage_launches(pc) + 2, (c.launch_count() - l, pc)
    got = _gather(ctxs, sids, g.K, g.Np)
    # one context, undecomposed
    q0 = _vortex4(g.node_coords())
    s1 = _setup_state(g, q0, kinds, bval)
    for _ in range(steps):
        g.euler_step_ssprk2(s1, GAMMA, dt)
    g.sync()
    single = g.download(s1, 0, 4)
    # oracle, undecomposed
    om = H.oracle_mesh(mg)
    case = o.Case(om, N, bc_kinds=kinds_oracle if kinds_oracle else None)
    bv = ([], [], [])
    for ip in range(len(om.patches)):
        b = _vortex4(case.patch_internal(case.geo.x, ip), 0.01)
        bv[0].append(b[:, 0]); bv[1].append(b[:, 1:3]); bv[2].append(b[:, 3])
    want = _oracle_steps(case, q0, bv, dt, steps)
    parts = np.bincount(c2p, minlength=nproc).tolist()
    for c in ctxs:
        c.close()
    g.close()
    return got, single, want, parts


def _check(got, single, want, steps):
    for f in range(4):
        assert H.rel_l2(got[..., f], single[..., f]) <= 1e-14, ("vs one context", f)
        assert H.rel_l2(got[..., f], want[..., f]) <= 1e-12 * steps, ("vs oracle", f)


@pytest.mark.parametrize("N", [2, 4])
def test_strips_of_the_periodic_square(built_library, N):
    """BASELINE configs[2] partition: `simple (1 2 1)` strips of the doubly periodic square; every cut face (also across the periodic
    wrap) is a processor face."""
    mg = meshgen.jittered_square(8, periodic=True)
    got, single, want, parts = _run_case(N, mg, (1, 2, 1), [])
    assert parts == [64, 64]
    _check(got, single, want, 3)


def test_simple_2x2_with_fixed_value_boundary(built_library):
    """`simple (2 2 1)` (four processors, up to three neighbours each) on the square with the reference's fixedValue boundary."""
    mg = meshgen.jittered_square(9)
    got, single, want, parts = _run_case(4, mg, (2, 2, 1), [o.BC_FIXED])
    assert len(parts) == 4 and min(parts) >= 40
    _check(got, single, want, 3)


def test_graph_partition_five_processors(built_library):
    """`method scotch` stand-in (hdg_decompose_graph): five irregular parts with several neighbours each."""
    mg = meshgen.jittered_square(10)
    g = H.HostContext()
    g.set_order(3)
    g.set_mesh_triangles(mg["xy"], mg["tris"], None, mg["patch_edges"])
    c2p = g.decompose_graph(5)
    got, single, want, parts = _run_case(3, mg, None, [o.BC_FIXED], c2p=c2p, nproc=5)
    assert parts == [40] * 5
    _check(got, single, want, 3)


def test_three_strips_ragged_octets(built_library):
    """Three processors, element counts not multiples of 8 (ragged last octet on every processor)."""
    mg = meshgen.jittered_square(7, periodic=True)
    got, single, want, parts = _run_case(3, mg, (3, 1, 1), [])
    assert sum(parts) == 98 and any(p % 8 for p in parts)
    _check(got, single, want, 3)


def test_cylinder_sectors_N6(built_library):
    """BASELINE configs[4] shape: closed O-grid around the cylinder (reflective wall + fixedValue far field), N=6, four angular sectors
    given as a manual cellToProc (the reference's `manual` method)."""
    n_r, n_th = 5, 16
    mg = meshgen.ogrid_sector(n_r, n_th, 0.0, 2 * np.pi, closed=True)
    mg["patch_edges"] = [mg["sides"]["left"], mg["sides"]["right"]]
    K = mg["tris"].shape[0]
    c2p = (np.arange(K) // (2 * n_r) * 4 // n_th).astype(np.int32)      # element rows are rings of constant theta index
    got, single, want, parts = _run_case(6, mg, None, [o.BC_REFLECTIVE, o.BC_FIXED], steps=2, dt=2e-4, c2p=c2p, nproc=4)
    assert parts == [K // 4] * 4
    _check(got, single, want, 2)


def test_state_written_between_steps_triggers_a_fresh_exchange(built_library):
    """The ghosts of a state are current after a parallel step; an upload in between must be followed by a new (blocking) exchange."""
    N, dt = 3, 1e-3
    mg = meshgen.jittered_square(6, periodic=True)
    g = _global_ctx(N, mg)
    c2p = g.decompose_simple(2, 1, 1)
    ctxs = _processors(g, N, c2p, 2)
    sids = [_setup_state(c, _vortex4(c.node_coords()), [], None) for c in ctxs]
    capi.group_euler_step_ssprk2(ctxs, sids, GAMMA, dt)
    for c, s in zip(ctxs, sids):                                         # restart from a different field
        c.upload(s, 0, _vortex4(c.node_coords(), 0.2))
    l0 = ctxs[0].launch_count()
    capi.group_euler_step_ssprk2(ctxs, sids, GAMMA, dt)
    capi.group_euler_step_ssprk2(ctxs, sids, GAMMA, dt)
    per_step = 2 * _stage_launches(ctxs[0].par_counts())
    assert ctxs[0].launch_count() - l0 == 2 * per_step + 2              # exactly one priming exchange
    got = _gather(ctxs, sids, g.K, g.Np)
    s1 = g.state_create(4)
    g.upload(s1, 0, _vortex4(g.node_coords(), 0.2))
    g.euler_step_ssprk2(s1, GAMMA, dt)
    g.euler_step_ssprk2(s1, GAMMA, dt)
    assert H.rel_l2(got, g.download(s1, 0, 4)) <= 1e-14


def test_legacy_pack_copy_unpack_entry_points(built_library):
    """hdg_halo_pack -> device copy -> hdg_halo_unpack driven by the caller (the entry points an MPI / NCCL host code binds), two
    processors, serial exchange before each stage."""
    import ctypes as C
    N, dt = 4, 1e-3
    mg = meshgen.jittered_square(6, periodic=True)
    g = _global_ctx(N, mg)
    c2p = g.decompose_simple(1, 2, 1)
    ctxs = _processors(g, N, c2p, 2)
    sids = [_setup_state(c, _vortex4(c.node_coords()), [], None) for c in ctxs]
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    pp = [[p for p, q in enumerate(c.proc_addressing()["patch_nbr_proc"]) if q >= 0] for c in ctxs]
    assert len(pp[0]) == 1 and len(pp[1]) == 1

    def exchange(which):
        bufs = []
        for c, s, ps in zip(ctxs, sids, pp):
            c.stream_wait(1, 0)
            ptr, n = c.halo_pack(s, which, ps[0])
            bufs.append((ptr, n))
        for c in ctxs:
            c.sync()
        for r, (c, s, ps) in enumerate(zip(ctxs, sids, pp)):
            recv = C.c_void_p()
            n = C.c_int64()
            c._ck(c.lib.hdg_halo_recv_buffer(c.h, s, ps[0], C.byref(recv), C.byref(n)))
            assert n.value == bufs[1 - r][1]
            assert rt.cudaMemcpy(recv, C.c_void_p(bufs[1 - r][0]), n.value * 8, 3) == 0      # device to device
            c.halo_unpack(s, which, ps[0])
            c.stream_wait(0, 1)

    for _ in range(2):
        exchange(0)
        for c, s in zip(ctxs, sids):
            c.euler_stage(s, GAMMA, dt, 0, 0.0, 1.0)
        exchange(1)
        for c, s in zip(ctxs, sids):
            c.euler_stage(s, GAMMA, dt, 1, 0.5, 0.5)
    for c in ctxs:
        c.sync()
    got = _gather(ctxs, sids, g.K, g.Np)
    s1 = g.state_create(4)
    g.upload(s1, 0, _vortex4(g.node_coords()))
    g.euler_step_ssprk2(s1, GAMMA, dt)
    g.euler_step_ssprk2(s1, GAMMA, dt)
    assert H.rel_l2(got, g.download(s1, 0, 4)) <= 1e-14
