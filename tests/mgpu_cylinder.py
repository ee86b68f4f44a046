#!/usr/bin/env python3
"""BASELINE configs[4] on N GPUs (torchrun, NCCL): subsonic Euler flow past a cylinder, O-grid of triangles, N=6, reflective (slip)
wall on the cylinder, fixedValue free stream on the far field, angular-sector partition with per-stage halo exchange.

  --check : small mesh, every sector compared with ONE GPU advancing the closed annulus (<= 1e-11 relative: the sector vertices are
            computed from their own theta range, so coordinates differ in the last bit)
  --perf  : 2000 x 1000 x 2 = 4.0 M triangles in total (configs[4] size), device-timed throughput
"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hopefoam_b200 import capi, partition  # noqa: E402

GAMMA, MACH = 1.4, 0.38


def free_stream(shape):
    rho = np.ones(shape)
    ru = np.full(shape, MACH)          # p = 1/gamma -> c = 1
    rv = np.zeros(shape)
    E = np.full(shape, 1.0 / (GAMMA * (GAMMA - 1.0)) + 0.5 * MACH * MACH)
    return np.stack([rho, ru, rv, E], -1)


def setup(ctx, part):
    ctx.set_mesh_triangles(part["xy"], part["tris"], part["point_equiv"], part["patch_edges"])
    sid = ctx.state_create(4)
    ctx.upload(sid, 0, free_stream((ctx.K, ctx.Np)))
    ctx.set_patch_kind(sid, part["wall_patch"], capi.BC_REFLECTIVE)
    ff = part["farfield_patch"]
    ctx.set_patch_kind(sid, ff, capi.BC_FIXED_VALUE)
    ctx.set_patch_values(sid, 0, ff, free_stream((ctx.patch_info(ff)[2] * ctx.Nfp,)))
    return sid


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--perf", action="store_true")
    ap.add_argument("--steps", type=int, default=40)
    args = ap.parse_args()
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    N = 6
    ok = True
    if args.check:
        n_r, n_th = 12, 8 * world
        dt, steps = 2e-4, 25
        ctx = capi.Context(lr)
        ctx.set_order(N)
        part = partition.sector_partition(n_r, n_th // world, world, rank)
        sid = setup(ctx, part)
        halo = partition.HaloExchanger(ctx, sid, part, dist, torch)
        for _ in range(steps):
            halo.step_ssprk2(GAMMA, dt)
        ctx.sync()
        mine = torch.from_numpy(ctx.download(sid, 0, 4)).cuda()
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        if rank == 0:
            c1 = capi.Context(lr)
            c1.set_order(N)
            g = partition.sector_partition(n_r, n_th, 1, 0)
            s1 = setup(c1, g)
            for _ in range(steps):
                c1.euler_step_ssprk2(s1, GAMMA, dt)
            c1.sync()
            ref = c1.download(s1, 0, 4)
            K = ctx.K
            dev = np.abs(ref - free_stream((c1.K, c1.Np))).max()
            for r in range(world):
                a, b = gathered[r].cpu().numpy(), ref[r * K:(r + 1) * K]
                err = np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())
                print(f"sector {r}: rel-L2 vs single GPU {err:.3e}", flush=True)
                ok = ok and err <= 1e-11
            ok = ok and dev > 1e-3 and np.isfinite(ref).all()       # the wall really deflects the flow
            print(f"max deviation from the free stream {dev:.3e}", flush=True)
            print("MGPU_CYLINDER_CHECK", "PASS" if ok else "FAIL", flush=True)
    if args.perf:
        n_r, n_th_total = 1000, 2000
        ctx = capi.Context(lr)
        ctx.set_order(N)
        part = partition.sector_partition(n_r, n_th_total // world, world, rank)
        sid = setup(ctx, part)
        halo = partition.HaloExchanger(ctx, sid, part, dist, torch) if world > 1 else None
        dt = 1e-6
        step = (lambda: halo.step_ssprk2(GAMMA, dt)) if halo else (lambda: ctx.euler_step_ssprk2(sid, GAMMA, dt))
        for _ in range(3):
            step()
        ctx.sync(); torch.cuda.synchronize(); dist.barrier()
        stream = torch.cuda.ExternalStream(ctx.stream(0))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(args.steps):
                step()
            e1.record(stream)
        ctx.sync(); torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        k = torch.tensor([float(ctx.K)], dtype=torch.float64, device="cuda")
        dist.all_reduce(k)
        finite = bool(np.isfinite(ctx.download(sid, 0, 1)).all())
        if rank == 0:
            ms = t.item() / args.steps
            print(f"CYLINDER_PERF n_gpus {world} triangles {int(k.item())} N {N} ms_per_step {ms:.4f} "
                  f"GDOF/s {2 * 4 * ctx.Np * k.item() / (ms * 1e-3) / 1e9:.2f} finite {finite}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
