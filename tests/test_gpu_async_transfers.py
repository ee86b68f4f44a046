"""Asynchronous state transfers (hdg_state_upload_async / hdg_state_download_async): two jobs alternating on one GPU with
upload(n+1) | step(n) | download(n-1) on three streams give bit-identical results to the synchronous upload -> step -> download
sequence, including the re-use of host buffers and device planes across iterations."""
import numpy as np
import pytest

from hopefoam_b200 import meshgen
from tests import helpers as H

pytestmark = pytest.mark.gpu

GAMMA, DT = 1.4, 1e-3


def _pinned(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy()).pin_memory()
    return t


@pytest.mark.parametrize("N,n", [(4, 24), (2, 16)])
def test_pipelined_jobs_equal_serial(gpu_ctx_factory, N, n):
    ctx = gpu_ctx_factory(N)
    mg = meshgen.jittered_square(n, periodic=True)
    ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    xy = ctx.node_coords()
    jobs = []
    for t0 in (0.0, 0.37):
        rho, rhoU, E = H.vortex_state(xy[..., 0], xy[..., 1], t0, GAMMA)
        rhoU3 = np.concatenate([rhoU, np.zeros_like(rhoU[..., :1])], axis=-1)          # the reference's 3-vector layout
        jobs.append((rho, rhoU3, E))
    nsteps = 5      # per job

    # serial reference: one state, synchronous copies
    sid = ctx.state_create(4)
    serial = []
    for rho, rhoU3, E in jobs:
        h = [_pinned(rho), _pinned(rhoU3), _pinned(E)]
        for _ in range(nsteps):
            ctx.upload_ptr(sid, 0, 1, h[0].data_ptr(), 1)
            ctx.upload_ptr(sid, 1, 2, h[1].data_ptr(), 3)
            ctx.upload_ptr(sid, 3, 1, h[2].data_ptr(), 1)
            ctx.euler_step_ssprk2(sid, GAMMA, DT)
            ctx.download_ptr(sid, 0, 1, h[0].data_ptr(), 1)
            ctx.download_ptr(sid, 1, 2, h[1].data_ptr(), 3)
            ctx.download_ptr(sid, 3, 1, h[2].data_ptr(), 1)
        serial.append([t.numpy().copy() for t in h])

    # pipelined: two states, two host buffer sets, asynchronous copies
    sids = [sid, ctx.state_create(4)]
    hosts = [[_pinned(a) for a in job] for job in jobs]

    def up(j):
        ctx.upload_ptr_async(sids[j], 0, 1, hosts[j][0].data_ptr(), 1)
        ctx.upload_ptr_async(sids[j], 1, 2, hosts[j][1].data_ptr(), 3)
        ctx.upload_ptr_async(sids[j], 3, 1, hosts[j][2].data_ptr(), 1)

    def down(j):
        ctx.download_ptr_async(sids[j], 0, 1, hosts[j][0].data_ptr(), 1)
        ctx.download_ptr_async(sids[j], 1, 2, hosts[j][1].data_ptr(), 3)
        ctx.download_ptr_async(sids[j], 3, 1, hosts[j][2].data_ptr(), 1)

    total = 2 * nsteps
    up(0)
    for i in range(total):
        j = i & 1
        ctx.euler_step_ssprk2(sids[j], GAMMA, DT)
        if i + 1 < total:
            up(1 - j)
        down(j)
    ctx.sync()
    for j in range(2):
        for got, ref in zip(hosts[j], serial[j]):
            assert np.array_equal(got.numpy(), ref)
        assert np.all(hosts[j][1].numpy()[..., 2] == 0.0)          # z of the 2-D momentum comes back zero-filled
    ctx.close()


def test_async_then_sync_calls_are_ordered(gpu_ctx_factory):
    """A synchronous download right after an asynchronous upload sees the uploaded data (the compute stream orders itself
    after the pending upload), and a synchronous upload after an asynchronous download does not disturb it."""
    ctx = gpu_ctx_factory(3)
    mg = meshgen.jittered_square(10, periodic=True)
    ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    rng = np.random.default_rng(3)
    a = rng.standard_normal((ctx.K, ctx.Np))
    b = rng.standard_normal((ctx.K, ctx.Np))
    sid = ctx.state_create(1)
    ha, out = _pinned(a), _pinned(np.zeros_like(a))
    ctx.upload_ptr_async(sid, 0, 1, ha.data_ptr(), 1)
    assert np.array_equal(ctx.download(sid, 0), a)
    ctx.download_ptr_async(sid, 0, 1, out.data_ptr(), 1)
    ctx.upload(sid, 0, b)
    ctx.sync()
    assert np.array_equal(out.numpy(), a)
    assert np.array_equal(ctx.download(sid, 0), b)
    ctx.close()
