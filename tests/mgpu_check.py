#!/usr/bin/env python3
"""Multi-GPU parity check, launched under torchrun (one rank per GPU, NCCL):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Every rank advances its strip with per-stage halo exchange; rank 0 also advances the undecomposed mesh on its own GPU and
compares strip by strip (<= 1e-13 relative: the halo is pure data movement; only the flux orientation on cut faces differs)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hopefoam_b200 import capi, partition  # noqa: E402
from bench import vortex_fields  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    N, n, dt, steps = 4, 24, 2e-3, 25
    ctx = capi.Context(lr)
    ctx.set_order(N)
    part = partition.strip_partition(n, world, rank)
    ctx.set_mesh_triangles(part["xy"], part["tris"], part["point_equiv"], part["patch_edges"])
    xy = ctx.node_coords()
    y_per = 10.0 * world

    def init(x, y):            # one vortex in the global domain, made periodic in y by hand (cheap images)
        tot = None
        for s in (-y_per, 0.0, y_per):
            r, ru, rv, e = vortex_fields(x, y - 5.0 - s)      # centred ON the cut between strips 0 and 1
            q = np.stack([r - 1.0, ru - 1.0, rv, e - (1 / 0.4 + 0.5)], -1)
            tot = q if tot is None else tot + q
        return tot + np.array([1.0, 1.0, 0.0, 1 / 0.4 + 0.5])
    q0 = init(xy[..., 0], xy[..., 1])
    sid = ctx.state_create(4)
    ctx.upload(sid, 0, q0)
    halo = partition.HaloExchanger(ctx, sid, part, dist, torch)
    overlap = os.environ.get("HDG_MGPU_OVERLAP", "1") == "1"
    for _ in range(steps):
        if overlap:                      # boundary rows first, exchange for the next stage under the interior launch
            halo.step_ssprk2(1.4, dt)
        else:
            halo.exchange(0)
            ctx.euler_stage(sid, 1.4, dt, 0, 0.0, 1.0)
            halo.exchange(1)
            ctx.euler_stage(sid, 1.4, dt, 1, 0.5, 0.5)
    ctx.sync()
    mine = torch.from_numpy(ctx.download(sid, 0, 4)).cuda()
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    ok = True
    if rank == 0:
        g = partition.global_mesh(n, world)
        c1 = capi.Context(lr)
        c1.set_order(N)
        c1.set_mesh_triangles(g["xy"], g["tris"], g["point_equiv"], [])
        gxy = c1.node_coords()
        s1 = c1.state_create(4)
        c1.upload(s1, 0, init(gxy[..., 0], gxy[..., 1]))
        for _ in range(steps):
            c1.euler_step_ssprk2(s1, 1.4, dt)
        c1.sync()
        ref = c1.download(s1, 0, 4)
        K = ctx.K
        moved_max = 0.0
        for r in range(world):
            a, b = gathered[r].cpu().numpy(), ref[r * K:(r + 1) * K]
            err = np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())
            moved = np.linalg.norm((b - init(gxy[..., 0], gxy[..., 1])[r * K:(r + 1) * K]).ravel())
            print(f"strip {r}: rel-L2 vs single GPU {err:.3e} (state moved by {moved:.3e})", flush=True)
            ok = ok and err <= 1e-13
            moved_max = max(moved_max, moved)
        ok = ok and moved_max > 1e-6          # the vortex sits on the cut between strips 0 and 1: the test is not vacuous
        print("MGPU_CHECK", "overlap" if overlap else "serial", "PASS" if ok else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
