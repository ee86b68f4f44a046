"""Synthetic unstructured triangle meshes of the BASELINE configs.

A square [x0,x1]x[y0,y1] of n x n quads, each split into two triangles with the diagonal direction
alternating by (i+j)&1; interior vertices jittered by U(-0.2h, 0.2h) (numpy default_rng(20240501)).
`periodic=True` glues left/right and bottom/top through a canonical point map (config 2-P); otherwise all four
sides form ONE patch (config 2-F: the reference's exact-solution fixedValue boundary).
"""
from __future__ import annotations

import numpy as np


def jittered_square(n: int, x0=0.0, x1=10.0, y0=-5.0, y1=5.0, periodic=False, jitter=0.2, seed=20240501):
    """Returns dict(xy (P,2) f64, tris (K,3) i32 CCW, point_equiv (P,) i32 | None, patch_edges [ (m,3) i32 ])."""
    hx, hy = (x1 - x0) / n, (y1 - y0) / n
    ii, jj = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="xy")     # jj = row (y), ii = column (x)
    x = x0 + hx * ii.astype(np.float64)
    y = y0 + hy * jj.astype(np.float64)
    rng = np.random.default_rng(seed)
    dx = rng.uniform(-jitter * hx, jitter * hx, size=x.shape)
    dy = rng.uniform(-jitter * hy, jitter * hy, size=y.shape)
    interior = (ii > 0) & (ii < n) & (jj > 0) & (jj < n)
    x = np.where(interior, x + dx, x)
    y = np.where(interior, y + dy, y)
    xy = np.stack([x.reshape(-1), y.reshape(-1)], axis=1)
    pid = lambda i, j: j * (n + 1) + i
    qi, qj = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    qi, qj = qi.reshape(-1), qj.reshape(-1)
    p00, p10, p11, p01 = pid(qi, qj), pid(qi + 1, qj), pid(qi + 1, qj + 1), pid(qi, qj + 1)
    alt = ((qi + qj) & 1) == 1
    # diagonal p00-p11 when alt == 0, p10-p01 when alt == 1 ; both triangles CCW
    t0 = np.where(alt[:, None], np.stack([p00, p10, p01], 1), np.stack([p00, p10, p11], 1))
    t1 = np.where(alt[:, None], np.stack([p10, p11, p01], 1), np.stack([p00, p11, p01], 1))
    tris = np.empty((2 * n * n, 3), dtype=np.int32)
    tris[0::2] = t0
    tris[1::2] = t1
    out = {"xy": xy, "tris": tris, "point_equiv": None, "patch_edges": []}
    if periodic:
        eq = np.arange((n + 1) * (n + 1), dtype=np.int32).reshape(n + 1, n + 1)   # [row j][col i]
        eq[:, n] = eq[:, 0]
        eq[n, :] = eq[0, :]
        out["point_equiv"] = eq.reshape(-1)
    else:
        q = lambda i, j: 2 * (j * n + i)           # first triangle of quad (i,j)
        edges = []
        k = np.arange(n)
        # bottom (j=0): edge p00-p10 belongs to t0 of quad (k,0) in both splittings
        edges.append(np.stack([q(k, 0), pid(k, 0), pid(k + 1, 0)], 1))
        # right (i=n-1): edge p10-p11: alt==0 -> t0 (p00,p10,p11); alt==1 -> t1 (p10,p11,p01)
        a = ((n - 1 + k) & 1)
        edges.append(np.stack([q(n - 1, k) + a, pid(n, k), pid(n, k + 1)], 1))
        # top (j=n-1): edge p11-p01 always in t1
        edges.append(np.stack([q(k, n - 1) + 1, pid(k + 1, n), pid(k, n)], 1))
        # left (i=0): edge p01-p00: alt==0 -> t1 (p00,p11,p01); alt==1 -> t0 (p00,p10,p01)
        a = ((0 + k) & 1)
        edges.append(np.stack([q(0, k) + (1 - a), pid(0, k + 1), pid(0, k)], 1))
        out["patch_edges"] = [np.concatenate(edges).astype(np.int32)]
    return out


def jittered_rect(nx: int, ny: int, x0=0.0, x1=10.0, y0=-5.0, y1=5.0, periodic=True, jitter=0.2, seed=20240501):
    """nx x ny quads of the `jittered_square` pattern on a rectangle (the global mesh of the weak-scaling runs: BASELINE configs[2] is
    `world` squares stacked in y).  periodic=True glues left/right and bottom/top; else all four sides form one patch."""
    uv, tris, sides = structured_triangles(nx, ny, jitter=jitter, seed=seed)
    xy = np.stack([x0 + (x1 - x0) * uv[:, 0] / nx, y0 + (y1 - y0) * uv[:, 1] / ny], axis=1)
    out = {"xy": xy, "tris": tris, "point_equiv": None, "patch_edges": []}
    if periodic:
        eq = np.arange((nx + 1) * (ny + 1), dtype=np.int32).reshape(ny + 1, nx + 1)
        eq[:, nx] = eq[:, 0]
        eq[ny, :] = eq[0, :]
        out["point_equiv"] = eq.reshape(-1)
    else:
        out["patch_edges"] = [np.concatenate([sides["bottom"], sides["right"], sides["top"], sides["left"]]).astype(np.int32)]
    return out


def structured_triangles(nx: int, ny: int, jitter=0.0, seed=20240501):
    """Parameter-space version of `jittered_square` on an nx x ny grid of unit quads: returns (uv (P,2) with u in [0,nx], v in [0,ny],
    tris (K,3) i32 CCW in (u,v), side edge lists).  Element order: quad (i,j) -> triangles 2*(j*nx+i), 2*(j*nx+i)+1 (rows of constant j)."""
    ii, jj = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    u, v = ii.astype(np.float64), jj.astype(np.float64)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        interior = (ii > 0) & (ii < nx) & (jj > 0) & (jj < ny)
        u = np.where(interior, u + rng.uniform(-jitter, jitter, size=u.shape), u)
        v = np.where(interior, v + rng.uniform(-jitter, jitter, size=v.shape), v)
    uv = np.stack([u.reshape(-1), v.reshape(-1)], axis=1)
    pid = lambda i, j: j * (nx + 1) + i
    qi, qj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    qi, qj = qi.reshape(-1), qj.reshape(-1)
    p00, p10, p11, p01 = pid(qi, qj), pid(qi + 1, qj), pid(qi + 1, qj + 1), pid(qi, qj + 1)
    alt = ((qi + qj) & 1) == 1
    t0 = np.where(alt[:, None], np.stack([p00, p10, p01], 1), np.stack([p00, p10, p11], 1))
    t1 = np.where(alt[:, None], np.stack([p10, p11, p01], 1), np.stack([p00, p11, p01], 1))
    tris = np.empty((2 * nx * ny, 3), dtype=np.int32)
    tris[0::2] = t0
    tris[1::2] = t1
    q = lambda i, j: 2 * (j * nx + i)
    kx, ky = np.arange(nx), np.arange(ny)
    sides = {
        "bottom": np.stack([q(kx, 0), pid(kx, 0), pid(kx + 1, 0)], 1),
        "right": np.stack([q(nx - 1, ky) + ((nx - 1 + ky) & 1), pid(nx, ky), pid(nx, ky + 1)], 1),
        "top": np.stack([q(kx, ny - 1) + 1, pid(kx + 1, ny), pid(kx, ny)], 1),
        "left": np.stack([q(0, ky) + (1 - ((0 + ky) & 1)), pid(0, ky + 1), pid(0, ky)], 1),
    }
    return uv, tris, {k: v.astype(np.int32) for k, v in sides.items()}


def ogrid_sector(n_r: int, n_theta: int, theta0: float, theta1: float, r0=0.5, r1=20.0, closed=False):
    """O-grid sector around a cylinder of diameter 2*r0 (BASELINE configs[4] geometry, straight-sided faces): n_r x n_theta quads split
    into triangles, geometric radial stretching r_i = r0 (r1/r0)^(i/n_r).  Element rows are rings of constant theta index.
    Returns xy, tris and the four sides: 'left' = cylinder wall (r0), 'right' = far field (r1), 'bottom'/'top' = the radial cuts at
    theta0/theta1.  closed=True glues theta1 onto theta0 (whole annulus on one GPU) through point_equiv."""
    uv, tris, sides = structured_triangles(n_r, n_theta)
    r = r0 * (r1 / r0) ** (uv[:, 0] / n_r)
    th = theta0 + (theta1 - theta0) * uv[:, 1] / n_theta
    xy = np.stack([r * np.cos(th), r * np.sin(th)], axis=1)
    out = {"xy": xy, "tris": tris, "sides": sides, "point_equiv": None}
    if closed:
        eq = np.arange((n_r + 1) * (n_theta + 1), dtype=np.int32).reshape(n_theta + 1, n_r + 1)
        xy2 = xy.reshape(n_theta + 1, n_r + 1, 2)
        xy2[n_theta] = xy2[0]                      # identical coordinates on the seam
        eq[n_theta, :] = eq[0, :]
        out["xy"] = xy2.reshape(-1, 2)
        out["point_equiv"] = eq.reshape(-1)
    return out
