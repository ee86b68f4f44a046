"""N>1 host logic on CPU: world_size 2 and 3 over the gloo backend.  Each rank builds its strip partition (host-only
context), "packs" the node coordinates of its processor-patch traces exactly as the device pack kernel does (owner
trace reversed per face, processorDgPatchField.C:253-260), runs the SAME message schedule as the GPU halo exchange
(partition.exchange_messages) and checks that what lands in each ghost slot is the coordinate of its own face node."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, N, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hopefoam_b200 import partition
        from tests import helpers as H
        part = partition.strip_partition(n, world, rank)
        c = H.HostContext()
        c.set_order(N)
        c.set_mesh_triangles(part["xy"], part["tris"], part["point_equiv"], part["patch_edges"])
        send, recv, mine = [], [], []
        for p in (0, 1):
            xy = c.patch_node_coords(p).reshape(n, c.Nfp, 2)
            mine.append(xy)
            send.append(torch.from_numpy(np.ascontiguousarray(xy[:, ::-1, :])).clone())     # sender-side reversal
            recv.append(torch.zeros_like(send[-1]))
        partition.exchange_messages(dist, send, recv, part["peers"])
        period = 10.0 * world
        ok = True
        for p in (0, 1):
            d = recv[p].numpy() - mine[p]
            d[..., 1] -= period * np.round(d[..., 1] / period)                              # periodic wrap in y
            ok = ok and bool(np.abs(d).max() < 1e-12)
        out[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_schedule_gloo(built_library, world):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, 5, 3, out), nprocs=world, join=True)
    assert dict(out) == {r: True for r in range(world)}
