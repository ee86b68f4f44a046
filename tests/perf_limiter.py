"""Timing of hdg_euler_limit on one GPU (not a test): python tests/perf_limiter.py [n] [N].  Wall clock around a synchronised batch of
calls (three launches per call); prints ms per call and the algorithmic HBM traffic rate (4 planes read + written once = 64*Np B per
element) against the measured HBM peak."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hopefoam_b200 import capi, meshgen  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mg = meshgen.jittered_square(n)
ctx = capi.Context(0)
ctx.set_order(N)
ctx.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
xy = ctx.node_coords()
x, y = xy[..., 0], xy[..., 1]
left = (x + 0.3 * y) < 5.0
rho = np.where(left, 8.0, 1.4) + 0.01 * np.sin(x)
U = np.stack([np.where(left, 57.0, 0.0), np.where(left, -33.0, 0.0)], -1)
E = np.where(left, 116.5, 1.0) / 0.4 + 0.5 * (U ** 2).sum(-1) / rho
sid = [ctx.state_create(1), ctx.state_create(2), ctx.state_create(1)]
for s, f in zip(sid, (rho, U, E)):
    ctx.upload(s, 0, f)
    ctx.set_patch_kind(s, 0, capi.BC_ZERO_GRADIENT)
for _ in range(3):
    ctx.euler_limit(*sid)
ctx.sync()
reps = 200
n_l0 = ctx.launch_count()
t0 = time.perf_counter()
for _ in range(reps):
    ctx.euler_limit(*sid)
ctx.sync()
ms = (time.perf_counter() - t0) / reps * 1e3
L = ctx.layout()
bytes_alg = 2 * 4 * ctx.K * ctx.Np * 8
import json  # noqa: E402
pk = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
peak = json.loads(pk.read_text())["hbm_gbs"] if pk.exists() else 6650.0
print(f"hdg_euler_limit: K={ctx.K} N={N}: {ms:.4f} ms per call ({(ctx.launch_count() - n_l0) // reps} launches), {ctx.K * ctx.Np / ms / 1e6:.2f} GDOF/s, "
      f"algorithmic traffic {bytes_alg / ms / 1e6:.1f} GB/s = {bytes_alg / ms / 1e6 / peak:.3f} of the HBM peak {peak:.0f} GB/s; finite={np.isfinite(ctx.download(sid[0], 0)).all()}")
