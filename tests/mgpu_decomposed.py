#!/usr/bin/env python3
"""Multi-GPU run on a mesh decomposed by the library's dgDecomposePar restatement (`method simple`), torchrun + NCCL:
every rank builds its processor mesh from the global mesh + cellToProc, advances the isentropic vortex with fixedValue far-field
data, and the result scattered back through cellProcAddressing is compared with one GPU on the undecomposed mesh (<= 1e-13)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hopefoam_b200 import capi, meshgen, partition  # noqa: E402
from bench import vortex_fields  # noqa: E402

DIV = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (4, 2, 1)}


def state_of(xy, t=0.0):
    r, ru, rv, e = vortex_fields(xy[..., 0], xy[..., 1], t)
    return np.stack([r, ru, rv, e], -1)


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    N, n, dt, steps = 4, 20, 2e-3, 25
    mg = meshgen.jittered_square(n)
    glob = capi.Context(lr)
    glob.set_order(N)
    glob.set_mesh_triangles(mg["xy"], mg["tris"], None, mg["patch_edges"])
    c2p = glob.decompose_simple(*DIV[world], 0.001)
    ctx = capi.Context(lr)
    ctx.set_order(N)
    ctx.set_mesh_from_decomposition(glob, c2p, world, rank)
    addr = ctx.proc_addressing()
    sid = ctx.state_create(4)
    ctx.upload(sid, 0, state_of(ctx.node_coords()))
    if ctx.patch_info(0)[2]:
        ctx.set_patch_values(sid, 0, 0, state_of(ctx.patch_node_coords(0)))
    halo = partition.GeneralHalo(ctx, sid, dist, torch)
    for _ in range(steps):
        halo.step_ssprk2(1.4, dt)
    ctx.sync()
    # scatter into the global numbering on every rank through cellProcAddressing and sum
    full = torch.zeros((glob.K, ctx.Np, 4), dtype=torch.float64, device="cuda")
    full[torch.from_numpy(addr["cell"].astype(np.int64)).cuda()] = torch.from_numpy(ctx.download(sid, 0, 4)).cuda()
    dist.all_reduce(full)
    ok = True
    if rank == 0:
        s1 = glob.state_create(4)
        glob.upload(s1, 0, state_of(glob.node_coords()))
        glob.set_patch_values(s1, 0, 0, state_of(glob.patch_node_coords(0)))
        for _ in range(steps):
            glob.euler_step_ssprk2(s1, 1.4, dt)
        glob.sync()
        ref = glob.download(s1, 0, 4)
        err = np.linalg.norm((full.cpu().numpy() - ref).ravel()) / np.linalg.norm(ref.ravel())
        nproc_patches = int((addr["patch_nbr_proc"] >= 0).sum())
        print(f"simple{DIV[world]}: cells per rank {np.bincount(c2p).tolist()}, rank 0 has {nproc_patches} processor patches, "
              f"rel-L2 vs single GPU {err:.3e}", flush=True)
        ok = err <= 1e-13
        print("MGPU_DECOMPOSED", "PASS" if ok else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
