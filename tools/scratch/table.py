import json, sys, time
from pathlib import Path
sys.path.insert(0, '/root/repo')
import numpy as np
from hopefoam_b200 import capi
from tests import helpers as H
gold = Path('/root/repo/tests/golden')
T = json.loads((gold/'golden_errors.json').read_text())['slide18_full']
for N in range(1, 7):
    for mi, mesh in enumerate(T['meshes']):
        d = np.load(gold/f'{mesh}.npz')
        ctx = capi.Context(0); ctx.set_order(N)
        ctx.set_mesh_triangles(d['xy'], d['tris'], None, [d['patch_edges']])
        xy, pxy = ctx.node_coords(), ctx.patch_node_coords(0)
        r, u, e = H.vortex_state(xy[..., 0], xy[..., 1], 0.0)
        sid = ctx.state_create(4); ctx.upload(sid, 0, r); ctx.upload(sid, 1, u); ctx.upload(sid, 3, e)
        dt = T['dt'][str(N)][mi]; t = 0.0; t0 = time.time()
        for _ in range(int(round(2.0/dt))):
            br, bu, be = H.vortex_state(pxy[:, 0], pxy[:, 1], t)
            ctx.set_patch_values(sid, 0, 0, br); ctx.set_patch_values(sid, 1, 0, bu); ctx.set_patch_values(sid, 3, 0, be)
            ctx.euler_step_ssprk2(sid, 1.4, dt); t += dt
        ctx.sync()
        rx, ux, _ = H.vortex_state(xy[..., 0], xy[..., 1], t)
        rho, rhoU, _ = H.download_euler(ctx, sid)
        er = np.abs(rho-rx).sum()/rx.size; eu = np.sqrt(((rhoU-ux)**2).sum(-1)).sum()/rx.size
        pr, pu = T['rho'][str(N)][mi], T['rhoU'][str(N)][mi]
        print(f"N={N} {mesh} dt={dt}: rho {er:.4e} (pub {pr:.3e}, rel {abs(er-pr)/pr:.1e})  rhoU {eu:.4e} (pub {pu:.3e}, rel {abs(eu-pu)/pu:.1e})  {time.time()-t0:.1f}s", flush=True)
        ctx.close()
