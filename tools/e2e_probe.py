#!/usr/bin/env python3
"""Where does bench.py's e2e leg lose against the raw link?  (run on the GPU box)  Times, with the same 1 M-triangle state and the
library's asynchronous transfer entry points: uploads only, downloads only, both without a step, and the full pipeline with 2 and 3 jobs."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hopefoam_b200 import capi, meshgen  # noqa: E402
from bench import vortex_fields  # noqa: E402

N, n = 4, 707
ctx = capi.Context(0)
ctx.set_order(N)
mg = meshgen.jittered_square(n, periodic=True)
ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
K, Np = ctx.K, ctx.Np
x, y = np.moveaxis(ctx.node_coords(), -1, 0)
rho, ru, rv, E = vortex_fields(x, y)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
mk = lambda: (pin(rho), pin(np.stack([ru, rv], -1)), pin(E))
J = 3
ins, outs = [mk() for _ in range(J)], [mk() for _ in range(J)]
sids = [ctx.state_create(4) for _ in range(J)]
bytes_step = K * Np * 8 * 4


def up(s, h):
    ctx.upload_ptr_async(s, 0, 1, h[0].data_ptr(), 1); ctx.upload_ptr_async(s, 1, 2, h[1].data_ptr(), 2); ctx.upload_ptr_async(s, 3, 1, h[2].data_ptr(), 1)


def down(s, h):
    ctx.download_ptr_async(s, 0, 1, h[0].data_ptr(), 1); ctx.download_ptr_async(s, 1, 2, h[1].data_ptr(), 2); ctx.download_ptr_async(s, 3, 1, h[2].data_ptr(), 1)


def timed(fn, reps):
    fn(2); ctx.sync(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(reps); ctx.sync(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def only_up(r):
    for i in range(r): up(sids[i % 2], ins[i % 2])


def only_down(r):
    for i in range(r): down(sids[i % 2], outs[i % 2])


def both(r):
    for i in range(r): up(sids[i % 2], ins[i % 2]); down(sids[(i + 1) % 2], outs[(i + 1) % 2])


def pipe(jobs):
    def run(r):
        up(sids[0], ins[0])
        for i in range(r):
            j = i % jobs
            if i + 1 < r: up(sids[(i + 1) % jobs], ins[(i + 1) % jobs])
            ctx.euler_step_ssprk2(sids[j], 1.4, 1.28e-4)
            down(sids[j], outs[j])
    return run


for name, fn in (("uploads only", only_up), ("downloads only", only_down), ("uploads + downloads, no step", both), ("pipeline, 2 jobs", pipe(2)), ("pipeline, 3 jobs", pipe(3))):
    ms = timed(fn, 12)
    print(f"{name:32s} {ms:7.2f} ms per step  = {bytes_step / ms / 1e6:5.1f} GB/s each way", flush=True)
