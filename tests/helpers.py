"""Shared helpers for the parity tests: oracle <-> product glue on identical inputs."""
import ctypes as C

import numpy as np

from hopefoam_b200 import capi, meshgen
from oracle import dg_oracle as o


class HostContext(capi.Context):
    """device = -1: operators + connectivity only (CPU tests of the host logic; compute calls raise)."""

    def __init__(self):
        self.lib = capi.load_library()
        h = C.c_void_p()
        assert self.lib.hdg_create(-1, C.byref(h)) == 0
        self.h = h
        self.N = None


def oracle_mesh(mg):
    """meshgen dict -> oracle DGMesh (same triangles, same patch edge order)."""
    pe = [[(int(c), (int(a), int(b))) for c, a, b in e] for e in mg["patch_edges"]]
    info = [{"name": f"patch{i}", "type": "patch"} for i in range(len(pe))]
    return o.build_connectivity(mg["xy"], mg["tris"], pe, info, point_equiv=mg["point_equiv"])


def vortex_state(x, y, t=0.0, gamma=1.4):
    r, ru, rv, e = o.vortex_exact(x, y, t, gamma)
    return r, np.stack([ru, rv], axis=-1), e


def rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def setup_euler(ctx, case, rho, rhoU, E, bR, bU, bE, kinds):
    """Create a 4-plane state holding (rho, rhoU, E) + patch kinds/values, mirroring the oracle's inputs."""
    sid = ctx.state_create(4)
    ctx.upload(sid, 0, rho)
    ctx.upload(sid, 1, rhoU)
    ctx.upload(sid, 3, E)
    for ip, kind in enumerate(kinds):
        ctx.set_patch_kind(sid, ip, {o.BC_FIXED: capi.BC_FIXED_VALUE, o.BC_ZEROGRAD: capi.BC_ZERO_GRADIENT,
                                     o.BC_REFLECTIVE: capi.BC_REFLECTIVE}[kind])
        if kind == o.BC_FIXED and case.mesh.patches[ip]["faces"].size:
            ctx.set_patch_values(sid, 0, ip, bR[ip])
            ctx.set_patch_values(sid, 1, ip, bU[ip])
            ctx.set_patch_values(sid, 3, ip, bE[ip])
    return sid


def download_euler(ctx, sid):
    rho = ctx.download(sid, 0)
    rhoU = ctx.download(sid, 1, 2)
    E = ctx.download(sid, 3)
    return rho, rhoU, E
