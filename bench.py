#!/usr/bin/env python3
"""bench.py - headline benchmark of the fused FP64 DG Euler stage (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W        # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference ...                   # the restated reference CPU path (oracle port) on host cores

Workload (BASELINE.json configs[1]): 2-D compressible Euler isentropic vortex, periodic square [0,10]x[-5,5],
707x707x2 = 999 698 jittered triangles per GPU, N=4, FP64, Roe flux, SSP-RK2 (the reference solver's scheme).
A "step" is one time step = 2 fused RK stages.  value = DOF-updates per second per RK stage summed over all
ranks (DOF-update = one nodal value of one conserved scalar advanced by one stage: 4*Np*K per stage).
Multi-GPU: weak scaling, one strip partition of 999 698 triangles per rank, per-stage halo exchange of the cut-face
traces over NCCL (torch.distributed) between neighbouring strips.

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "FP64 GDOF-updates/s per RK stage (2-D Euler, N=4)"
UNIT = "GDOF/s"
GAMMA = 1.4


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


FP64_PEAK_TFLOPS = 37.1      # measured on this pool with tools/microbench/fp64_peak.cu (profiles/fp64_peak_r01.txt)


def ncu_traffic(order, K):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE stage-kernel launch from the committed ncu capture (per launch, like
    `achieved`); only valid for the configuration it was captured on."""
    p = ROOT / "profiles" / "ncu_traffic_r01.json"
    if not p.exists() or order != 4 or K != 999698:
        return None
    return json.loads(p.read_text())["dram_bytes_per_launch"]


def algorithmic_bytes_per_element_stage(Np, n_scalars=4):
    """SSP-RK2 mean (r+w) = 2.5 state passes + 128 B geometry/connectivity."""
    return n_scalars * Np * 8 * 2.5 + 128


def algorithmic_flops_per_element_stage(Np, Ng, Nfg, Nfp):
    """SURVEY.md §8-d Fl(N)."""
    return 24 * Ng * Np + 24 * Np * Nfg + 48 * Nfg * Nfp + 54 * Ng + 480 * Nfg


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index=0):
        self.rows = []
        self.proc = None
        self.device_index = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------------------

def vortex_fields(x, y, t=0.0, gamma=GAMMA, beta=5.0):
    """Isentropic vortex (TUT/isentropicVortex/dgEulerFoam/setNonUniformInlet.H:19-27); product-side initial data."""
    r = (x - 5.0 - t) ** 2 + y ** 2
    rho = np.power(1.0 - (gamma - 1.0) * beta * beta * np.exp(2.0 * (1.0 - r)) / (16.0 * gamma * np.pi * np.pi), 1.0 / (gamma - 1.0))
    ru = (1 - beta * np.exp(1 - r) * y / (2.0 * np.pi)) * rho
    rv = (beta * np.exp(1 - r) * (x - 5 - t) / (2.0 * np.pi)) * rho
    E = np.power(rho, gamma) / (gamma - 1.0) + 0.5 * (ru * ru + rv * rv) / rho
    return rho, ru, rv, E


ADVECT_NCU_TRAFFIC = {4: 643491328}      # dram__bytes_read + dram__bytes_write of one advectStageTmaKernel<4> launch at 999 698 triangles
                                          # (profiles/ncu_advect_r01g.md: 529.1 MB + 114.4 MB)


def run_advection(args, quiet=False):
    """Secondary measurement (not the headline line): fused scalar-advection stage (BASELINE configs[0] physics at 1 M triangles),
    the HBM-bound sibling of the Euler stage.  Algorithmic bytes per element-stage: 20*Np + 16*Np (nodal U read) + 128."""
    import torch
    from hopefoam_b200 import capi, meshgen
    ctx = capi.Context(int(os.environ.get("LOCAL_RANK", "0")))
    N = args.order
    ctx.set_order(N)
    mg = meshgen.jittered_square(args.n, x0=-1, x1=1, y0=-1, y1=1, periodic=True)
    ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    xy = ctx.node_coords()
    T = np.exp(-((xy[..., 0] + 0.3) ** 2 + (xy[..., 1] + 0.3) ** 2) / (2 * 0.1 ** 2))
    U = np.stack([np.full_like(T, 1.0), np.full_like(T, 0.5)], -1)
    sT, sU = ctx.state_create(1), ctx.state_create(2)
    ctx.upload(sT, 0, T)
    ctx.upload(sU, 0, U)
    dt = 1e-5
    steps = min(args.steps, 200)
    stream = torch.cuda.ExternalStream(ctx.stream(0))
    for _ in range(max(args.warmup, 3)):
        ctx.advect_step_ssprk2(sT, sU, dt)
    ctx.sync()
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(steps):
            ctx.advect_step_ssprk2(sT, sU, dt)
        ev1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    launches = ctx.launch_count() - l0
    K, Np = ctx.K, ctx.Np
    hbm_peak, src = measured_peaks()
    bytes_stage = (20 * Np + 16 * Np + 128) * K
    tma = N in (3, 4) and os.environ.get("HDG_ADV_CFG", "1") != "0"
    out = {"metric": "FP64 GDOF-updates/s per RK stage (2-D scalar advection, LF)", "value": 2 * Np * K / (ms * 1e-3) / 1e9, "unit": UNIT,
           "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"2-D scalar advection, nodal U, LF flux, periodic, {K} triangles, N={N}, SSP-RK2 (2 fused stages per step)",
                      "l2_policy": "T, U and geometry streams of one stage (%.0f MB) exceed the 126 MB L2; no explicit flush" % (bytes_stage / 1e6)},
           "roofline": {"bound": "hbm", "achieved": bytes_stage / (ms / 2 * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": bytes_stage / (ms / 2 * 1e-3) / 1e9 / hbm_peak,
                        "traffic": ADVECT_NCU_TRAFFIC.get(N) if (tma and abs(K - 999698) < 1000) else None, "peak_source": src,
                        "kernel": f"advectStageTmaKernel<{N}>" if tma else f"advectStageKernel<{N}>", "kernel_ms": ms / 2,
                        "algorithmic_bytes_per_element_stage": 36 * Np + 128, "algorithmic_bytes_per_launch": bytes_stage},
           "gpu_launches": int(launches)}
    if not quiet:
        print(json.dumps(out), flush=True)
    ctx.close()
    return out


def run_gpu(args):
    import torch
    from hopefoam_b200 import capi, meshgen
    from hopefoam_b200 import partition

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    N = args.order
    n = args.n
    ctx = capi.Context(local_rank)
    ctx.set_order(N)
    # weak scaling: every rank owns one [0,10]x[-5,5]-sized strip of n x n x 2 triangles of a global periodic mesh that is
    # `world` strips tall; cut faces between strips are processor patches exchanged every stage.
    part = partition.strip_partition(n, world, rank)
    ctx.set_mesh_triangles(part["xy"], part["tris"], part["point_equiv"], part["patch_edges"])
    K, Np = ctx.K, ctx.Np
    xy = ctx.node_coords()
    rho, ru, rv, E = vortex_fields(xy[..., 0], xy[..., 1] - part["y_shift"])
    # pinned host staging of the whole state in the reference's AoS layout (rho | rhoU as 3-vectors | Ener)
    h_rho = torch.from_numpy(rho).pin_memory()
    h_rhoU = torch.from_numpy(np.stack([ru, rv, np.zeros_like(ru)], axis=-1)).pin_memory()
    h_E = torch.from_numpy(E).pin_memory()
    sid = ctx.state_create(4)

    def upload():
        ctx.upload_ptr(sid, 0, 1, h_rho.data_ptr(), 1)
        ctx.upload_ptr(sid, 1, 2, h_rhoU.data_ptr(), 3)
        ctx.upload_ptr(sid, 3, 1, h_E.data_ptr(), 1)

    def download():
        ctx.download_ptr(sid, 0, 1, h_rho.data_ptr(), 1)
        ctx.download_ptr(sid, 1, 2, h_rhoU.data_ptr(), 3)
        ctx.download_ptr(sid, 3, 1, h_E.data_ptr(), 1)

    if args.rk == "lserk45" and world > 1:
        raise SystemExit("--rk lserk45 is measured on one GPU")
    upload()
    halo = partition.HaloExchanger(ctx, sid, part, dist, torch) if world > 1 else None
    dt = args.dt
    stream = torch.cuda.ExternalStream(ctx.stream(0))

    def step():
        if args.rk == "lserk45":
            ctx.euler_step_lserk45(sid, GAMMA, dt)          # single GPU only: 5 fused stages, 2N-storage
        elif halo is None:
            ctx.euler_step_ssprk2(sid, GAMMA, dt)
        elif args.no_overlap:
            halo.exchange(0)
            ctx.euler_stage(sid, GAMMA, dt, 0, 0.0, 1.0)
            halo.exchange(1)
            ctx.euler_stage(sid, GAMMA, dt, 1, 0.5, 0.5)
        else:
            halo.step_ssprk2(GAMMA, dt)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput (value) ---------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - l0
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    k_total = torch.tensor([float(K)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(k_total)
    K_all = float(k_total.item())
    stages = 5 if args.rk == "lserk45" else 2
    dof_per_step = stages * 4 * Np * K_all
    value = dof_per_step / (ms_step * 1e-3) / 1e9

    # ---- kernel-level roofline: the stage kernel alone, timed live with CUDA events on its stream ---------
    barrier()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(2 * min(args.steps, 50))]
    with torch.cuda.stream(stream):
        for i, (a, b) in enumerate(kev):
            a.record(stream)
            ctx.euler_stage(sid, GAMMA, dt, i & 1, 0.0 if (i & 1) == 0 else 0.5, 1.0 if (i & 1) == 0 else 0.5)
            b.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    hbm_peak, peak_src = measured_peaks()
    bytes_per_launch = algorithmic_bytes_per_element_stage(Np) * K
    achieved_gbs = bytes_per_launch / (k_ms * 1e-3) / 1e9
    flops_per_launch = algorithmic_flops_per_element_stage(Np, ctx.Ng, ctx.Nfg, ctx.Nfp) * K
    achieved_tf = flops_per_launch / (k_ms * 1e-3) / 1e12

    # ---- end-to-end through the C ABI with HOST buffers: every step uploads its input state from pinned host memory, advances it
    # by one SSP-RK2 step and downloads the result.  On one GPU two independent jobs (two host buffer sets, two device states)
    # alternate, so that through the asynchronous transfer entry points the upload of job B overlaps the stage kernels of job A
    # and the download of the job before (PCIe is full duplex); every step still moves its own input and its own result.
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    pipelined = args.rk == "ssprk2" and not args.e2e_serial and not (world > 1 and args.no_overlap)
    if pipelined:
        e2e_steps = max(e2e_steps, 24)      # amortises the fill and the drain of the three-stream pipeline
        hosts = [(h_rho, h_rhoU, h_E), (h_rho.clone().pin_memory(), h_rhoU.clone().pin_memory(), h_E.clone().pin_memory())]
        sids = [sid, ctx.state_create(4)]
        halos = [halo, None]
        if halo is not None:
            # second job on every rank: its own state (ghost traces included), the context's message buffers shared; its halo
            # is primed with synchronous calls so that the first exchange cannot overtake the first asynchronous upload
            ctx.upload_ptr(sids[1], 0, 1, hosts[1][0].data_ptr(), 1)
            ctx.upload_ptr(sids[1], 1, 2, hosts[1][1].data_ptr(), 3)
            ctx.upload_ptr(sids[1], 3, 1, hosts[1][2].data_ptr(), 1)
            halos[1] = partition.HaloExchanger(ctx, sids[1], part, dist, torch, share=halo)
            barrier()
            halos[1].exchange(0)
            halos[1]._primed = True
            barrier()

        def upload_async(j):
            r, u, e = hosts[j]
            ctx.upload_ptr_async(sids[j], 0, 1, r.data_ptr(), 1)
            ctx.upload_ptr_async(sids[j], 1, 2, u.data_ptr(), 3)
            ctx.upload_ptr_async(sids[j], 3, 1, e.data_ptr(), 1)

        def download_async(j):
            r, u, e = hosts[j]
            ctx.download_ptr_async(sids[j], 0, 1, r.data_ptr(), 1)
            ctx.download_ptr_async(sids[j], 1, 2, u.data_ptr(), 3)
            ctx.download_ptr_async(sids[j], 3, 1, e.data_ptr(), 1)

        def e2e_loop(nsteps):
            upload_async(0)
            for i in range(nsteps):
                j = i & 1
                if halos[j] is None:
                    ctx.euler_step_ssprk2(sids[j], GAMMA, dt)
                else:
                    halos[j].step_ssprk2(GAMMA, dt)
                    ctx.stream_wait(0, 1)      # the step's last exchange (halo stream) is ordered before the download / the next upload
                if i + 1 < nsteps:
                    upload_async(1 - j)
                download_async(j)
            ctx.sync()

        e2e_loop(2)          # warm-up: staging rings, second state
        barrier()
        t0 = time.perf_counter()
        e2e_loop(e2e_steps)
        barrier()
        e2e_s = time.perf_counter() - t0
    else:
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            upload()
            step()
            download()
        barrier()
        e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = dof_per_step * e2e_steps / float(te.item()) / 1e9
    state_bytes = K * Np * 8 * 5            # rho + rhoU(3) + E doubles per node, the reference's host layout
    finite = bool(np.isfinite(h_rho.numpy()).all())

    out = None
    if rank == 0:
        # on rank 0 at N=1 only: under torchrun the other ranks spin in the barrier and OMP_NUM_THREADS is forced to 1
        cpu = cpu_baseline(args, sample_only=True) if (not args.no_cpu and world == 1) else None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"2-D Euler isentropic vortex, periodic, {int(K_all)} jittered triangles ({K} per GPU), N={N}, "
                                   f"Roe flux, {'LSERK(5,4) (5 fused stages per step)' if args.rk == 'lserk45' else 'SSP-RK2 (2 fused stages per step)'}, dt={dt}",
                       "order": N, "elements_per_gpu": K, "stages_per_step": stages, "partition": ("strips, halo overlapped with interior" if not args.no_overlap else "strips, serial halo") if world > 1 else "none",
                       "l2_policy": "state per copy (%.0f MB) exceeds the 126 MB L2; no explicit flush" % (4 * K * 16 * 8 / 1e6)},
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                         "traffic": ncu_traffic(N, K), "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": peak_src, "kernel": "eulerStageKernel<4>", "kernel_ms": k_ms,
                         "algorithmic_bytes_per_element_stage": algorithmic_bytes_per_element_stage(Np),
                         "note": "the Euler stage with the reference's 3(N+1) cubature is FP64-pipe-bound (SURVEY §8-d); see fp64"},
            "fp64": {"achieved": achieved_tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved_tf / FP64_PEAK_TFLOPS,
                     "algorithmic_flops_per_element_stage": algorithmic_flops_per_element_stage(Np, ctx.Ng, ctx.Nfg, ctx.Nfp),
                     "peak_source": "measured DFMA/DMMA peak, profiles/fp64_peak_r01.txt"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes,
                    "steps": e2e_steps, "finite": finite,
                    "mode": ("two jobs alternating: upload(n+1) | step(n) | download(n-1) on three streams" if pipelined
                             else "serial upload -> step -> download")},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    if rank == 0:
        if world == 1 and args.rk == "ssprk2" and not args.no_advection:
            # the HBM-bound sibling of the headline kernel, measured in the same run (it is the kernel the 70 %-of-HBM target applies to)
            try:
                out["advection"] = run_advection(args, quiet=True)
            except Exception as ex:  # noqa: BLE001 - the headline line must survive a failure of the secondary measurement
                out["advection"] = {"error": str(ex)}
        print(json.dumps(out), flush=True)
    return out


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the restated reference CPU path (oracle) on the host cores
# ------------------------------------------------------------------------------------------------------------

def cpu_baseline(args, sample_only=False):
    """Times oracle/ref_cpu (C port of the reference's per-stage loop structure: AoS fields, stored per-element
    cellD1dx and mass matrices, three equation passes per stage) with all host threads on a bounded sample."""
    from oracle import ref_cpu
    N = args.order
    n = args.cpu_n
    threads = os.cpu_count() or 1
    res = ref_cpu.time_euler_steps(N=N, n=n, steps=args.cpu_steps, threads=threads, dt=args.dt)
    val = res["dof_updates_per_s"] / 1e9
    return {"value": val, "unit": UNIT, "cores": res["threads"], "kind": "port",
            "sample": f"{res['K']} triangles (same generator, {n}x{n}x2, periodic), N={N}, {args.cpu_steps} SSP-RK2 steps, "
                      f"{res['seconds']:.2f} s wall; restated reference CPU path (oracle/ref_cpu.c), not the HopeFOAM binary"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_cpu
    N = args.order
    threads = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        ref_cpu.time_euler_steps(N=N, n=args.cpu_n, steps=1, threads=threads, dt=args.dt)
    res = ref_cpu.time_euler_steps(N=N, n=args.cpu_n, steps=max(1, min(args.steps, args.cpu_steps)), threads=threads, dt=args.dt)
    val = res["dof_updates_per_s"] / 1e9
    nsteps = max(1, min(args.steps, args.cpu_steps))
    cb = {"value": val, "unit": UNIT, "cores": res["threads"], "kind": "port",
          "sample": f"{res['K']} triangles ({args.cpu_n}x{args.cpu_n}x2 periodic, same generator), N={N}, {nsteps} SSP-RK2 steps per timing"}
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": nsteps, "warmup": min(args.warmup, 1),
           "ms_per_step": res["seconds"] / nsteps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"2-D Euler isentropic vortex, periodic, N={N}, Roe flux, SSP-RK2; bounded sample of {res['K']} triangles "
                                  "per step on the host cores (per-element cost is size-independent)", "order": N},
           "cpu_baseline": cb,
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--mesh-n", dest="n", type=int, default=707, help="quads per side per GPU (707 -> 999 698 triangles)")
    ap.add_argument("--dt", type=float, default=1.28e-4)
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--e2e-serial", action="store_true", help="e2e without overlapping transfers and compute (one job, synchronous copies)")
    ap.add_argument("--cpu-n", type=int, default=200, help="CPU sample: quads per side (200 -> 80 000 triangles)")
    ap.add_argument("--cpu-steps", type=int, default=64, help="SSP-RK2 steps of the CPU sample (about 10 s on 16 cores)")
    ap.add_argument("--rk", default="ssprk2", choices=["ssprk2", "lserk45"], help="lserk45: the low-storage RK of createFields.H:119-138 (1 GPU)")
    ap.add_argument("--workload", default="euler", choices=["euler", "advection"], help="advection = secondary HBM-bound measurement")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-advection", action="store_true", help="skip the secondary scalar-advection measurement attached to the 1-GPU line")
    ap.add_argument("--no-overlap", action="store_true", help="multi-GPU: serialise halo exchange and stage (A/B of the overlap)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "advection":
        run_advection(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
