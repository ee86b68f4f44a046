"""CPU ORACLE (test infrastructure, NOT product code) for HopeFOAM's explicit 2-D nodal-DG stage.

This module is a numpy restatement of the reference algorithm.  It exists only so that tests,
`__graft_entry__.smoke()` and `bench.py --impl reference / cpu_baseline` can check and time
against it.  Nothing under `hopefoam_b200/` imports it.

Parity pinning: the reference cannot be compiled in this container (needs PETSc/SLEPc/MPI/flex,
SURVEY.md §8-c), so the oracle is pinned END-TO-END against the reference's published
isentropic-vortex errors (HopeFOAM-0.1 User Guide §1.8: rhoError 8.807979526797244e-06,
rhoUError 1.865574862711117e-05 at N=4 / vortex1024.msh / dt=0.004 / t=2) and the workshop
convergence table (tests/test_oracle_golden.py).  There is no per-stage golden dump in the
reference, so per-stage parity is "GPU vs this oracle".

All file:line citations are relative to /root/reference/HopeFOAM-0.1/ (DG/ = src/DG/,
TUT/ = tutorials/DG/2D/).

Conventions restated from the reference
  * tensor slots of dxdr: [0]=x_r [1]=y_r [3]=x_s [4]=y_s ; drdx: [0]=rx [1]=sx [3]=ry [4]=sy
    (DG/element/baseFunctions/straightBaseFunctions/triangleBaseFunction/triangleBaseFunction.C:315-343)
  * element nodes: rows of constant s, r increasing (triangleBaseFunction.C:134-142)
  * local faces: f0 = v0->v1, f1 = v1->v2, f2 = v2->v0 (triangleBaseFunction.C:75-94)
  * dgFace order: cell-major, local-face-minor, created by the poly owner (dgPolyMesh.C:346-396)
"""
from __future__ import annotations

import json
import math
import re
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent

# --------------------------------------------------------------------------------------------
# 1. Jacobi polynomials, Gauss / Gauss-Lobatto nodes, Vandermonde matrices
#    DG/element/polynomials/Legendre/Legendre.C
# --------------------------------------------------------------------------------------------


def _gamma_int(x) -> float:
    """Legendre::gamma(int x) = (x-1)!  (Legendre.C:243-247; argument is truncated to int)."""
    x = int(x)
    ans = 1.0
    for i in range(2, x):
        ans *= i
    return ans


def jacobi_gq(alpha: float, beta: float, N: int):
    """Gauss quadrature nodes/weights, Legendre.C:39-166.

    The reference builds the symmetric tridiagonal Jacobi matrix with ZERO diagonal (only valid for
    alpha == beta, which is all it ever uses), solves it with SLEPc EPS (un-vendored; any symmetric
    eigensolver gives the same unique answer), sorts ascending and symmetrises the weights.
    """
    if N == 0:
        return np.array([-(alpha - beta) / (alpha + beta + 2)]), np.array([2.0])
    A = np.zeros((N + 1, N + 1))
    for i in range(N):
        j = i + 1
        v = 2.0 / (2 * i + alpha + beta + 2) * math.sqrt(
            j * (j + alpha + beta) * (j + alpha) * (j + beta) / (2 * i + alpha + beta + 1) / (2 * i + alpha + beta + 3))
        A[i, j] = v
        A[j, i] = v
    lam, vec = np.linalg.eigh(A)
    w = vec[0, :] ** 2 * 2.0 ** (alpha + beta + 1) / (alpha + beta + 1) * _gamma_int(alpha + 1) * _gamma_int(beta + 1) \
        / _gamma_int(alpha + beta + 1)
    order = np.argsort(lam, kind="stable")
    lam, w = lam[order].copy(), w[order].copy()
    for i in range((N - 1) // 2 + 1):                       # Legendre.C:154-157
        t = 0.5 * (w[i] + w[N - i])
        w[i] = w[N - i] = t
    return lam, w


def gauss_jacobi(alpha: float, beta: float, n: int):
    """n-point Gauss-Jacobi rule for the weight (1-x)^alpha (1+x)^beta by Golub-Welsch with the FULL three-term recurrence (diagonal
    included).  NOT a reference function: the reference's JacobiGQ is only valid for alpha == beta (see jacobi_gq); this is used for the
    collapsed cubature of the orders the reference's table does not reach (N = 9, 10; parity unpinned)."""
    A = np.zeros((n, n))
    for k in range(n):
        d = (2 * k + alpha + beta) * (2 * k + alpha + beta + 2)
        A[k, k] = (beta * beta - alpha * alpha) / d if d != 0 else (beta - alpha) / (alpha + beta + 2)
        if k + 1 < n:
            h = 2 * k + alpha + beta
            v = 2.0 / (h + 2) * math.sqrt((k + 1) * (k + 1 + alpha + beta) * (k + 1 + alpha) * (k + 1 + beta) / (h + 1) / (h + 3))
            A[k, k + 1] = A[k + 1, k] = v
    lam, vec = np.linalg.eigh(A)
    mu0 = 2.0 ** (alpha + beta + 1) * math.gamma(alpha + 1) * math.gamma(beta + 1) / math.gamma(alpha + beta + 2)
    return lam, vec[0, :] ** 2 * mu0


def collapsed_cubature(order: int):
    """Cubature of the reference triangle exact to degree >= `order`: the conical (collapsed-coordinate) product of an m-point
    Gauss-Legendre rule in a and an m-point Gauss-Jacobi(1,0) rule in b, m = order // 2 + 1, with r = (1+a)(1-b)/2 - 1, s = b and
    weight w_a w_b / 2 (sum = 2, the area).  Point order: b outer, a inner.  Used for volIntOrder > 28 only (N = 9, 10: BASELINE
    configs[3] asks for the sweep beyond the reference's table, gaussTriangleIntegration.C:50-64)."""
    m = order // 2 + 1
    xa, wa = gauss_jacobi(0.0, 0.0, m)
    xb, wb = gauss_jacobi(1.0, 0.0, m)
    r = np.array([(1 + a) * (1 - b) / 2 - 1 for b in xb for a in xa])
    s_ = np.array([b for b in xb for a in xa])
    w = np.array([u * v / 2 for v in wb for u in wa])
    return r, s_, w


def jacobi_gl(alpha: float, beta: float, N: int):
    """Gauss-Lobatto nodes, Legendre.C:169-183."""
    x = np.zeros(N + 1)
    x[0], x[N] = -1.0, 1.0
    if N < 2:
        return x
    xi, _ = jacobi_gq(alpha + 1, beta + 1, N - 2)
    x[1:N] = xi
    return x


def jacobi_p(x, alpha: float, beta: float, N: int):
    """Orthonormal Jacobi polynomial P_N^(alpha,beta)(x), Legendre.C:185-220."""
    x = np.asarray(x, dtype=float)
    gamma0 = 2.0 ** (alpha + beta + 1) / (alpha + beta + 1) * _gamma_int(alpha + 1) * _gamma_int(beta + 1) / _gamma_int(alpha + beta + 1)
    gamma1 = (alpha + 1) * (beta + 1) / (alpha + beta + 3) * gamma0
    PL = [np.full_like(x, 1.0 / math.sqrt(gamma0))]
    if N == 0:
        return PL[0]
    PL.append(((alpha + beta + 2) / 2 * x + (alpha - beta) / 2) / math.sqrt(gamma1))
    if N == 1:
        return PL[1]
    aold = 2.0 / (2 + alpha + beta) * math.sqrt((alpha + 1) * (beta + 1) / (alpha + beta + 3))
    for i in range(1, N):
        h1 = 2.0 * i + alpha + beta
        anew = 2.0 / (h1 + 2) * math.sqrt((i + 1) * (i + 1 + alpha + beta) * (i + 1 + alpha) * (i + 1 + beta) / (h1 + 1) / (h1 + 3))
        bnew = -(alpha * alpha - beta * beta) / h1 / (h1 + 2)
        PL.append(1.0 / anew * (-aold * PL[i - 1] + (x - bnew) * PL[i]))
        aold = anew
    return PL[N]


def grad_jacobi_p(x, alpha: float, beta: float, N: int):
    """Legendre.C:222-232."""
    x = np.asarray(x, dtype=float)
    if N == 0:
        return np.zeros_like(x)
    return math.sqrt(N * (N + alpha + beta + 1)) * jacobi_p(x, alpha + 1, beta + 1, N - 1)


def vandermonde1d(N: int, r):
    """Legendre.C:250-262."""
    r = np.asarray(r, dtype=float)
    return np.stack([jacobi_p(r, 0, 0, i) for i in range(N + 1)], axis=1)


def _rs_to_ab(r, s):
    a = np.where(s != 1.0, 2 * (1 + r) / np.where(s != 1.0, 1 - s, 1.0) - 1, -1.0)
    return a, s.copy()


def vandermonde2d(N: int, r, s):
    """Legendre.C:273-299."""
    r = np.asarray(r, dtype=float)
    s = np.asarray(s, dtype=float)
    a, b = _rs_to_ab(r, s)
    V = np.zeros((r.size, (N + 1) * (N + 2) // 2))
    sk = 0
    for i in range(N + 1):
        for j in range(N - i + 1):
            V[:, sk] = math.sqrt(2.0) * jacobi_p(a, 0, 0, i) * jacobi_p(b, 2 * i + 1, 0, j) * (1 - b) ** i
            sk += 1
    return V


def grad_vandermonde2d(N: int, r, s):
    """Legendre.C:387-450; returns (Vr, Vs)."""
    r = np.asarray(r, dtype=float)
    s = np.asarray(s, dtype=float)
    a, b = _rs_to_ab(r, s)
    n = (N + 1) * (N + 2) // 2
    Vr = np.zeros((r.size, n))
    Vs = np.zeros((r.size, n))
    sk = 0
    for i in range(N + 1):
        for j in range(N - i + 1):
            fa = jacobi_p(a, 0, 0, i)
            gb = jacobi_p(b, 2 * i + 1, 0, j)
            dfa = grad_jacobi_p(a, 0, 0, i)
            dgb = grad_jacobi_p(b, 2 * i + 1, 0, j)
            c = 2.0 ** (i + 0.5)
            if i > 0:
                Vr[:, sk] = dfa * gb * c * (0.5 * (1 - b)) ** (i - 1)
                tmp = dgb * (0.5 * (1 - b)) ** i - 0.5 * i * gb * (0.5 * (1 - b)) ** (i - 1)
                Vs[:, sk] = (dfa * gb * 0.5 * (1 + a) * (0.5 * (1 - b)) ** (i - 1) + fa * tmp) * c
            else:
                Vr[:, sk] = dfa * gb * c
                tmp = dgb
                Vs[:, sk] = (dfa * gb * 0.5 * (1 + a) + fa * tmp) * c
            sk += 1
    return Vr, Vs


# --------------------------------------------------------------------------------------------
# 2. Reference element: Warp&Blend nodes, face maps, cubature, interpolation matrices
#    triangleBaseFunction.C, lineBaseFunction.C, gaussIntegration.C, gaussTriangleIntegration.C
# --------------------------------------------------------------------------------------------

_ALPOPT = [0.0000, 0.0000, 1.4152, 0.1001, 0.2751, 0.9800, 1.0999,
           1.2832, 1.3648, 1.4773, 1.4959, 1.5743, 1.5770, 1.6223, 1.6258]

_CUB = None


def cubature_table(order: int):
    """gaussTriangleIntegration::dataTable(order) (…DataTable.C:32-316); data in cubature_tri.json."""
    global _CUB
    if _CUB is None:
        _CUB = json.loads((_HERE / "cubature_tri.json").read_text())
    d = _CUB[str(order)]
    return (np.array([float.fromhex(x) for x in d["r"]]), np.array([float.fromhex(x) for x in d["s"]]),
            np.array([float.fromhex(x) for x in d["w"]]))


def _warp_factor(N: int, rout):
    """triangleBaseFunction.C:188-233."""
    LGLr = jacobi_gl(0, 0, N)
    req = 2.0 / N * np.arange(N + 1) - 1
    Veq = vandermonde1d(N, req)
    Pmat = np.stack([jacobi_p(rout, 0, 0, i) for i in range(N + 1)], axis=0)
    Lmat = np.linalg.inv(Veq.T) @ Pmat
    warp = Lmat.T @ (LGLr - req)
    zerof = (np.abs(rout) < 1.0 - 1.0e-10).astype(float)
    sf = 1.0 - (zerof * rout) ** 2
    return warp / sf + warp * (zerof - 1.0)


def warp_blend_nodes(N: int):
    """triangleBaseFunction::initDofLocation, triangleBaseFunction.C:122-186 -> (r, s)."""
    Np = (N + 1) * (N + 2) // 2
    alpha = _ALPOPT[N - 1] if N < 16 else 5.0 / 3.0
    L1 = np.zeros(Np)
    L3 = np.zeros(Np)
    sk = 0
    for n in range(1, N + 2):
        for m in range(1, N + 3 - n):
            L1[sk] = (n - 1) / N
            L3[sk] = (m - 1) / N
            sk += 1
    L2 = 1.0 - L1 - L3
    x = L3 - L2
    y = (2 * L1 - L2 - L3) / math.sqrt(3.0)
    blend1, blend2, blend3 = 4 * L2 * L3, 4 * L1 * L3, 4 * L1 * L2
    w1 = blend1 * _warp_factor(N, L3 - L2) * (1 + (alpha * L1) ** 2)
    w2 = blend2 * _warp_factor(N, L1 - L3) * (1 + (alpha * L2) ** 2)
    w3 = blend3 * _warp_factor(N, L2 - L1) * (1 + (alpha * L3) ** 2)
    x = x + 1 * w1 + math.cos(2 * math.pi / 3) * w2 + math.cos(4 * math.pi / 3) * w3
    y = y + 0 * w1 + math.sin(2 * math.pi / 3) * w2 + math.sin(4 * math.pi / 3) * w3
    l1 = (math.sqrt(3.0) * y + 1.0) / 3.0
    l2 = (-3.0 * x - math.sqrt(3.0) * y + 2.0) / 6.0
    l3 = (3.0 * x - math.sqrt(3.0) * y + 2.0) / 6.0
    return l3 - l2 - l1, l1 - l2 - l3


def face_to_cell_index(N: int):
    """faceToCellIndex_[face][rotate][i], triangleBaseFunction.C:75-94."""
    Nfp = N + 1
    Np = (N + 1) * (N + 2) // 2
    idx = np.zeros((3, 2, Nfp), dtype=np.int32)
    idx[0, 0, :] = np.arange(Nfp)
    idx[1, 0, 0] = N
    idx[2, 0, 0] = Np - 1
    for i in range(1, Nfp):
        idx[1, 0, i] = idx[1, 0, i - 1] + (Nfp - i)
        idx[2, 0, i] = idx[2, 0, i - 1] - i - 1
    for f in range(3):
        idx[f, 1, :] = idx[f, 0, ::-1]
    return idx


@dataclass
class RefElement:
    """stdElement = triangleBaseFunction + gaussTriangleIntegration (DG/element/stdElement/stdElement.H)."""
    N: int
    Np: int = 0
    Nfp: int = 0
    Ng: int = 0
    Nfg: int = 0
    r: np.ndarray = None
    s: np.ndarray = None
    V: np.ndarray = None
    invV: np.ndarray = None
    Dr: np.ndarray = None          # drMatrix_ r-part (Np x Np)
    Ds: np.ndarray = None
    f2c: np.ndarray = None         # faceToCellIndex_
    gr: np.ndarray = None          # cubature points
    gs: np.ndarray = None
    gw: np.ndarray = None
    Vg: np.ndarray = None          # cellVandermonde_  (Ng x Np)
    Dgr: np.ndarray = None         # cellDr_ r-part    (Ng x Np)
    Dgs: np.ndarray = None
    fx: np.ndarray = None          # face Gauss nodes
    fw: np.ndarray = None
    If: np.ndarray = None          # faceInterp_ (Nfg x Nfp)

    def __post_init__(self):
        N = self.N
        self.Np = (N + 1) * (N + 2) // 2
        self.Nfp = N + 1
        self.r, self.s = warp_blend_nodes(N)
        self.V = vandermonde2d(N, self.r, self.s)
        self.invV = np.linalg.inv(self.V)                      # Legendre::matrixInv (PETSc LU) Legendre.C:540-618
        Vr, Vs = grad_vandermonde2d(N, self.r, self.s)
        self.Dr, self.Ds = Vr @ self.invV, Vs @ self.invV      # triangleBaseFunction.C:235-246
        self.f2c = face_to_cell_index(N)
        vol_order = 3 * (N + 1)                                # gaussIntegration.C:66
        face_order = 2 * (N + 1)                               # gaussIntegration.C:68
        if vol_order > 33:
            raise ValueError(f"volIntOrder_ = {vol_order} is not implemented")   # gaussTriangleIntegration.C:59-64 stops at 28 (N = 8)
        if vol_order > 28:      # N = 9, 10: beyond the reference's table (it aborts there) - own collapsed Gauss-Jacobi rule, parity unpinned
            self.gr, self.gs, self.gw = collapsed_cubature(vol_order)
        else:
            self.gr, self.gs, self.gw = cubature_table(vol_order)
        self.Ng = self.gr.size
        self.Vg = vandermonde2d(N, self.gr, self.gs) @ self.invV
        Vgr, Vgs = grad_vandermonde2d(N, self.gr, self.gs)
        self.Dgr, self.Dgs = Vgr @ self.invV, Vgs @ self.invV
        self.fx, self.fw = jacobi_gq(0, 0, face_order // 2)    # gaussTriangleIntegration.C:80-86
        self.Nfg = self.fx.size
        lgl = jacobi_gl(0, 0, N)                               # lineBaseFunction.C:52-63
        invV1 = np.linalg.inv(vandermonde1d(N, lgl))
        self.If = vandermonde1d(N, self.fx) @ invV1            # gaussTriangleIntegration.C:88-95


# --------------------------------------------------------------------------------------------
# 3. Mesh input: Fluent .msh (tutorial fixtures) and OpenFOAM polyMesh -> 2-D triangles
# --------------------------------------------------------------------------------------------


def read_fluent_msh(path):
    """Minimal Fluent ASCII reader for the tutorial fixtures.

    Returns points (P,2), faces list of (n0, n1, c0, c1, zone) with 1-based->0-based node ids and
    cell ids (c == -1 for none), and zone bc types {zone: bcType}.
    """
    text = Path(path).read_text()
    pts = None
    faces = []
    zones = {}
    pos = 0
    # nodes
    for m in re.finditer(r"\(10\s*\(([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s*([0-9a-fA-F]*)\)\s*\(", text):
        zone = int(m.group(1), 16)
        if zone == 0:
            continue
        first, last = int(m.group(2), 16), int(m.group(3), 16)
        nd = int(m.group(5), 16) if m.group(5) else 2
        end = text.index(")", m.end())
        vals = np.array(text[m.end():end].split(), dtype=float).reshape(-1, nd)
        if pts is None:
            pts = np.zeros((last, 2))
        pts[first - 1:last, :] = vals[:, :2]
    for m in re.finditer(r"\(13\s*\(([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\)\s*\(", text):
        zone = int(m.group(1), 16)
        if zone == 0:
            continue
        bctype = int(m.group(4), 16)
        ftype = int(m.group(5), 16)
        end = text.index(")", m.end())
        toks = text[m.end():end].split()
        zones[zone] = bctype
        stride = 4 if ftype != 0 else None
        i = 0
        while i < len(toks):
            if ftype == 0:
                nn = int(toks[i], 16)
                i += 1
            else:
                nn = ftype
            assert nn == 2, "2-D line faces expected"
            n0, n1 = int(toks[i], 16) - 1, int(toks[i + 1], 16) - 1
            c0, c1 = int(toks[i + 2], 16) - 1, int(toks[i + 3], 16) - 1
            faces.append((n0, n1, c0, c1, zone))
            i += 4
    return pts, faces, zones


def triangles_from_fluent(pts, faces):
    """Assemble CCW triangles + boundary edge zones from Fluent face records.

    Vertex order inside a cell is NOT the one fluentMeshToFoam would produce (that tool needs flex
    and cannot be built here); the end-of-run error norms are invariant under it.
    Returns tris (K,3) int32 CCW, boundary dict {(min(n0,n1), max(n0,n1)): zone}.
    """
    ncell = max(max(f[2], f[3]) for f in faces) + 1
    cell_edges = [[] for _ in range(ncell)]
    bnd = {}
    for n0, n1, c0, c1, zone in faces:
        for c in (c0, c1):
            if c >= 0:
                cell_edges[c].append((n0, n1))
        if c0 < 0 or c1 < 0:
            bnd[(min(n0, n1), max(n0, n1))] = zone
    tris = np.zeros((ncell, 3), dtype=np.int32)
    for c, edges in enumerate(cell_edges):
        assert len(edges) == 3, "triangles expected"
        a, b = edges[0]
        others = set(edges[1]) | set(edges[2])
        cpt = (others - {a, b}).pop()
        v = [a, b, cpt]
        A = pts[v[1]] - pts[v[0]]
        B = pts[v[2]] - pts[v[0]]
        if A[0] * B[1] - A[1] * B[0] < 0:
            v[1], v[2] = v[2], v[1]
        tris[c] = v
    return tris, bnd


# ---- OpenFOAM polyMesh (ASCII) ---------------------------------------------------------------

def _strip_foam(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    m = re.search(r"FoamFile\s*\{.*?\}", text, flags=re.S)
    if m:
        text = text[:m.start()] + text[m.end():]
    return text


def read_polymesh(dirpath):
    """Read constant/polyMesh/{points,faces,owner,neighbour,boundary} (ASCII)."""
    d = Path(dirpath)
    t = _strip_foam((d / "points").read_text())
    n = int(re.search(r"(\d+)\s*\(", t).group(1))
    nums = re.findall(r"\(\s*([-+0-9.eE]+)\s+([-+0-9.eE]+)\s+([-+0-9.eE]+)\s*\)", t)
    points = np.array(nums, dtype=float)
    assert points.shape[0] == n
    t = _strip_foam((d / "faces").read_text())
    faces = [np.array(g.split(), dtype=np.int64) for g in re.findall(r"\d+\s*\(([0-9\s]+)\)", t[t.index("(") + 1:])]
    def _labels(name):
        tt = _strip_foam((d / name).read_text())
        body = tt[tt.index("(") + 1: tt.rindex(")")]
        return np.array(body.split(), dtype=np.int64)
    owner = _labels("owner")
    neighbour = _labels("neighbour")
    t = _strip_foam((d / "boundary").read_text())
    t = re.sub(r"#\{.*?#\}", "", t, flags=re.S)
    patches = []
    for m in re.finditer(r"(\w+)\s*\{([^{}]*)\}", t):
        body = m.group(2)
        typ = re.search(r"type\s+(\w+)\s*;", body)
        nf = re.search(r"nFaces\s+(\d+)\s*;", body)
        sf = re.search(r"startFace\s+(\d+)\s*;", body)
        if typ and nf and sf:
            patches.append({"name": m.group(1), "type": typ.group(1), "nFaces": int(nf.group(1)), "startFace": int(sf.group(1))})
    assert len(faces) == owner.size
    return {"points": points, "faces": faces, "owner": owner, "neighbour": neighbour, "patches": patches}


def triangles_from_polymesh(pm):
    """dgPolyMesh 2-D rules (dgPolyMesh.C:154-190 z==0 face, :490-509 CCW swap).

    Returns tris (K,3) int64 of poly point labels (CCW, v0 = first point of the z==0 face as stored),
    xy (npoints,2), and per-patch list of (owner cell, (pA,pB)) boundary edges in polyPatch face order
    (dgPatch.C:70-100; `empty` patches carry no dgFaces).
    """
    P, faces, owner, neighbour = pm["points"], pm["faces"], pm["owner"], pm["neighbour"]
    ncell = int(owner.max()) + 1
    tris = -np.ones((ncell, 3), dtype=np.int64)
    cells_faces = [[] for _ in range(ncell)]
    for f, o in enumerate(owner):
        cells_faces[o].append(f)
    for f, nb in enumerate(neighbour):
        cells_faces[nb].append(f)
    for c in range(ncell):
        for f in cells_faces[c]:
            if np.all(P[faces[f], 2] == 0.0):
                assert faces[f].size == 3, "prism cells expected (tri base)"
                v = list(faces[f])
                A = P[v[1]] - P[v[0]]
                B = P[v[2]] - P[v[0]]
                if A[0] * B[1] - A[1] * B[0] < 0:
                    v[1], v[2] = v[2], v[1]
                tris[c] = v
    assert (tris >= 0).all()
    patch_edges = []
    for p in pm["patches"]:
        lst = []
        if p["type"] != "empty":
            for f in range(p["startFace"], p["startFace"] + p["nFaces"]):
                pts0 = [q for q in faces[f] if P[q, 2] == 0.0]
                assert len(pts0) == 2
                lst.append((int(owner[f]), (int(pts0[0]), int(pts0[1]))))
        patch_edges.append(lst)
    return tris, P[:, :2].copy(), patch_edges


# --------------------------------------------------------------------------------------------
# 4. DG connectivity (dgPolyMesh.C / physicalElementData.C / dgPatch.C)
# --------------------------------------------------------------------------------------------


@dataclass
class DGMesh:
    K: int
    xy: np.ndarray                 # (P,2)
    tris: np.ndarray               # (K,3) CCW vertex ids
    F: int = 0
    face_owner: np.ndarray = None  # (F,) int32
    face_nbr: np.ndarray = None    # (F,) int32, -1 on patches
    face_loc_o: np.ndarray = None  # faceIndexInOwner_
    face_loc_n: np.ndarray = None  # faceIndexInNeighbour_ (-1 on patches)
    face_rot: np.ndarray = None    # faceRotate_ = firstPointIndex_ (-1 on patches)
    cell_face: np.ndarray = None   # (K,3) dgCellFaceNewID_
    patches: list = field(default_factory=list)   # [{'name','type','faces': int32 array of dgFace ids}]


def build_connectivity(xy, tris, patch_edges=None, patch_info=None, point_equiv=None) -> DGMesh:
    """Restates dgPolyMesh::calNeighbourFace/calFirstPointIndex/initialDgFace (dgPolyMesh.C:346-410,
    835-896, 999-1038) for a conforming triangle mesh given as CCW vertex triples.

    patch_edges: list (per patch) of [(ownerCell, (pA,pB)), ...] in polyPatch face order.
    point_equiv: optional canonical point ids used for edge matching = periodic gluing (extension: the
    reference has no compiled cyclic patch).
    """
    tris = np.asarray(tris)
    K = tris.shape[0]
    ctris = tris if point_equiv is None else np.asarray(point_equiv)[tris]
    edge_map = {}
    for c in range(K):
        for f in range(3):
            a, b = int(ctris[c, f]), int(ctris[c, (f + 1) % 3])
            edge_map.setdefault((min(a, b), max(a, b)), []).append((c, f))
    nbr = -np.ones((K, 3), dtype=np.int64)
    nbr_face = -np.ones((K, 3), dtype=np.int64)
    for key, lst in edge_map.items():
        if len(lst) == 2:
            (c0, f0), (c1, f1) = lst
            nbr[c0, f0], nbr_face[c0, f0] = c1, f1
            nbr[c1, f1], nbr_face[c1, f1] = c0, f0
        elif len(lst) > 2:
            raise ValueError("non-manifold edge")
    fo, fn, flo, fln, frot = [], [], [], [], []
    cell_face = -np.ones((K, 3), dtype=np.int32)
    for c in range(K):
        for f in range(3):
            nb = nbr[c, f]
            if nb >= 0 and nb < c:
                continue                       # created by the (lower-numbered) poly owner
            if nb == c:
                raise ValueError("self-neighbour")
            fid = len(fo)
            fo.append(c)
            fn.append(nb)
            flo.append(f)
            cell_face[c, f] = fid
            if nb >= 0:
                nf = int(nbr_face[c, f])
                fln.append(nf)
                cell_face[nb, nf] = fid
                first = ctris[c, f]
                frot.append(0 if ctris[nb, nf] == first else 1)  # calFirstPointIndex, dgPolyMesh.C:868-896
            else:
                fln.append(-1)
                frot.append(-1)
    mesh = DGMesh(K=K, xy=np.asarray(xy, dtype=float), tris=tris.astype(np.int64))
    mesh.F = len(fo)
    mesh.face_owner = np.array(fo, dtype=np.int32)
    mesh.face_nbr = np.array(fn, dtype=np.int32)
    mesh.face_loc_o = np.array(flo, dtype=np.int32)
    mesh.face_loc_n = np.array(fln, dtype=np.int32)
    mesh.face_rot = np.array(frot, dtype=np.int32)
    mesh.cell_face = cell_face
    if patch_edges is not None:
        for ip, lst in enumerate(patch_edges):
            ids = []
            for (c, (pa, pb)) in lst:
                found = -1
                for f in range(3):
                    a, b = int(tris[c, f]), int(tris[c, (f + 1) % 3])
                    if {a, b} == {pa, pb}:
                        found = cell_face[c, f]
                        break
                assert found >= 0
                ids.append(found)
            info = patch_info[ip] if patch_info else {"name": f"patch{ip}", "type": "patch"}
            mesh.patches.append({"name": info["name"], "type": info["type"], "faces": np.array(ids, dtype=np.int32)})
    return mesh


def mesh_from_fluent(path, bc_names=None) -> DGMesh:
    pts, faces, zones = read_fluent_msh(path)
    tris, bnd = triangles_from_fluent(pts, faces)
    # group boundary edges by zone, ordered by (owner cell, local face) for determinism
    by_zone = {}
    for c in range(tris.shape[0]):
        for f in range(3):
            a, b = int(tris[c, f]), int(tris[c, (f + 1) % 3])
            key = (min(a, b), max(a, b))
            if key in bnd:
                by_zone.setdefault(bnd[key], []).append((c, (a, b)))
    zl = sorted(by_zone)
    pe = [by_zone[z] for z in zl]
    info = [{"name": (bc_names or {}).get(z, f"zone{z}"), "type": "patch"} for z in zl]
    return build_connectivity(pts, tris, pe, info)


def mesh_from_polymesh(dirpath) -> DGMesh:
    pm = read_polymesh(dirpath)
    tris, xy, pe = triangles_from_polymesh(pm)
    return build_connectivity(xy, tris, pe, [{"name": p["name"], "type": p["type"]} for p in pm["patches"]])


# --------------------------------------------------------------------------------------------
# 5. Physical element data (physicalCellElement.C:72-110, physicalElementData.C:228-263)
# --------------------------------------------------------------------------------------------


@dataclass
class Geometry:
    x: np.ndarray        # dofLocation_ (K,Np,2)
    WJ: np.ndarray       # jacobianWeights_ (K,Ng)
    M: np.ndarray        # massMatrix_ (K,Np,Np)
    D1x: np.ndarray      # cellD1dx_ x-part (K,Ng,Np)
    D1y: np.ndarray
    fnx: np.ndarray      # faceNx_ per dgFace (F,Nfg,2) in the owner's orientation
    fWJ: np.ndarray      # faceWJ_ (F,Nfg)
    Minv: np.ndarray = None


def build_geometry(mesh: DGMesh, ref: RefElement) -> Geometry:
    v = mesh.xy[mesh.tris]                                             # (K,3,2)
    r, s = ref.r[None, :, None], ref.s[None, :, None]
    x = -(r + s) * 0.5 * v[:, 0:1, :] + (r + 1) * 0.5 * v[:, 1:2, :] + (s + 1) * 0.5 * v[:, 2:3, :]   # physicalNodesLoc
    # dxdr at nodes = drMatrix (x) nodes  (triangleBaseFunction.C:315-328)
    xr = np.einsum("ij,kjc->kic", ref.Dr, x)       # (K,Np,2): x_r, y_r
    xs = np.einsum("ij,kjc->kic", ref.Ds, x)
    # cellDxdr at cubature points = cellDr_ (x) nodes (gaussIntegration.C:72-79)
    gxr = np.einsum("gj,kjc->kgc", ref.Dgr, x)
    gxs = np.einsum("gj,kjc->kgc", ref.Dgs, x)
    J = gxr[..., 0] * gxs[..., 1] - gxr[..., 1] * gxs[..., 0]          # xr*ys - yr*xs
    WJ = J * ref.gw[None, :]
    rx, sx = gxs[..., 1] / J, -gxr[..., 1] / J                          # drdx slots [0],[1]
    ry, sy = -gxs[..., 0] / J, gxr[..., 0] / J                          # slots [3],[4]
    D1x = ref.Dgr[None] * rx[..., None] + ref.Dgs[None] * sx[..., None]  # physicalCellElement.C:88-96
    D1y = ref.Dgr[None] * ry[..., None] + ref.Dgs[None] * sy[..., None]
    M = np.einsum("gi,kg,gj->kij", ref.Vg, WJ, ref.Vg)                  # physicalCellElement.C:101-110
    # faces: dxdr gathered at the owner's face nodes, interpolated to face Gauss points
    F = mesh.F
    fnx = np.zeros((F, ref.Nfg, 2))
    fWJ = np.zeros((F, ref.Nfg))
    o, lf = mesh.face_owner, mesh.face_loc_o
    nodes = ref.f2c[lf, 0, :]                                          # (F,Nfp)
    fxr = np.einsum("gi,fic->fgc", ref.If, xr[o[:, None], nodes])      # (F,Nfg,2)
    fxs = np.einsum("gi,fic->fgc", ref.If, xs[o[:, None], nodes])
    nx = np.where(lf[:, None] == 0, fxr[..., 1], np.where(lf[:, None] == 1, fxs[..., 1] - fxr[..., 1], -fxs[..., 1]))
    ny = np.where(lf[:, None] == 0, -fxr[..., 0], np.where(lf[:, None] == 1, fxr[..., 0] - fxs[..., 0], fxs[..., 0]))
    Js = np.sqrt(nx * nx + ny * ny)                                    # gaussTriangleIntegration.C:153-192
    fnx[..., 0], fnx[..., 1] = nx / Js, ny / Js
    fWJ = Js * ref.fw[None, :]
    return Geometry(x=x, WJ=WJ, M=M, D1x=D1x, D1y=D1y, fnx=fnx, fWJ=fWJ, Minv=np.linalg.inv(M))


# --------------------------------------------------------------------------------------------
# 6. Fields on quadrature points, boundary conditions (dgGaussField.C:188-269, dgPatchFields)
# --------------------------------------------------------------------------------------------

BC_FIXED, BC_ZEROGRAD, BC_REFLECTIVE = 0, 1, 2


class Case:
    """Holds mesh + operators + per-patch BC kinds; the analogue of dgMesh + boundary field types."""

    def __init__(self, mesh: DGMesh, N: int, bc_kinds=None):
        self.mesh = mesh
        self.ref = RefElement(N)
        self.geo = build_geometry(mesh, self.ref)
        self.bc_kinds = list(bc_kinds) if bc_kinds is not None else [BC_FIXED] * len(mesh.patches)
        ref = self.ref
        m = mesh
        # owner / neighbour dof mappings per dgFace ("vmapM"/"vmapP", physicalFaceElement.C:80-93)
        self.map_o = m.face_owner[:, None].astype(np.int64) * ref.Np + ref.f2c[m.face_loc_o, 0, :]
        interior = m.face_nbr >= 0
        self.interior = interior
        mp = np.zeros_like(self.map_o)
        mp[interior] = m.face_nbr[interior, None].astype(np.int64) * ref.Np + \
            ref.f2c[m.face_loc_n[interior], m.face_rot[interior], :]
        self.map_n = mp
        # patch dof offsets (physicalElementData.C:163-183): per patch, faces in dgFaceIndex order, Nfp each
        self.patch_of_face = -np.ones(m.F, dtype=np.int64)
        self.patch_off = -np.ones(m.F, dtype=np.int64)
        for ip, p in enumerate(m.patches):
            self.patch_of_face[p["faces"]] = ip
            self.patch_off[p["faces"]] = np.arange(p["faces"].size) * ref.Np * 0 + np.arange(p["faces"].size) * ref.Nfp

    # patchInternalField (dgPatchField.C:268-290)
    def patch_internal(self, q, ip):
        faces = self.mesh.patches[ip]["faces"]
        flat = q.reshape(-1, *q.shape[2:])
        return flat[self.map_o[faces].reshape(-1)]

    def patch_normals(self, ip):
        """reflective: nHat[i] = faceNx_[i] for i < Nfp (reflectiveDgPatchField.C:128-136)."""
        faces = self.mesh.patches[ip]["faces"]
        return self.geo.fnx[faces][:, :self.ref.Nfp, :].reshape(-1, 2)

    def evaluate_bc(self, q, bvals, is_vector=False):
        """boundaryField.evaluate(): fixedValue keeps values (fixedValueDgPatchField.C:108-119),
        zeroGradient copies the interior trace (zeroGradientDgPatchField.C:101-110), reflective applies
        transform(I-2nn) to the interior trace (reflectiveDgPatchField.C:111-154; scalars unchanged)."""
        for ip, kind in enumerate(self.bc_kinds):
            if self.mesh.patches[ip]["faces"].size == 0:
                continue
            if kind == BC_FIXED:
                continue
            tr = self.patch_internal(q, ip)
            if kind == BC_ZEROGRAD or not is_vector:
                bvals[ip] = tr.copy()
            else:
                n = self.patch_normals(ip)
                dot = tr[:, 0] * n[:, 0] + tr[:, 1] * n[:, 1]
                bvals[ip] = tr - 2.0 * dot[:, None] * n
        return bvals

    def gauss_field(self, q, bvals):
        """dgGaussField::initField: returns (cell (K,Ng,..), ownerFace (F,Nfg,..), neighborFace (F,Nfg,..))."""
        ref, m = self.ref, self.mesh
        cell = np.einsum("gj,kj...->kg...", ref.Vg, q)
        flat = q.reshape(-1, *q.shape[2:])
        own = np.einsum("gi,fi...->fg...", ref.If, flat[self.map_o])
        nbr = np.einsum("gi,fi...->fg...", ref.If, flat[self.map_n])
        for ip, p in enumerate(m.patches):
            if p["faces"].size == 0:
                continue
            bv = bvals[ip].reshape(p["faces"].size, ref.Nfp, *q.shape[2:])
            nbr[p["faces"]] = np.einsum("gi,fi...->fg...", ref.If, bv)
        return cell, own, nbr


# --------------------------------------------------------------------------------------------
# 7. Fluxes: Roe (RoeFlux.C:46-191), LF nodal (LFFlux.C:105-211)
# --------------------------------------------------------------------------------------------


def roe_flux(nx, ny, rhoM, ruM, rvM, EM, rhoP, ruP, rvP, EP, gamma):
    """Pointwise Roe flux F*.n in the owner's outward direction; M = owner, P = neighbour."""
    QM2 = nx * ruM + ny * rvM
    QP2 = nx * ruP + ny * rvP
    QM3 = nx * rvM - ny * ruM
    QP3 = nx * rvP - ny * ruP
    uM, uP = QM2 / rhoM, QP2 / rhoP
    vM, vP = QM3 / rhoM, QP3 / rhoP
    pM = (gamma - 1) * (EM - 0.5 * (QM2 * uM + QM3 * vM))
    pP = (gamma - 1) * (EP - 0.5 * (QP2 * uP + QP3 * vP))
    HM, HP = (EM + pM) / rhoM, (EP + pP) / rhoP
    fR = (QM2 + QP2) / 2
    fU = (QM2 * uM + pM + QP2 * uP + pP) / 2
    fV = (QM3 * uM + QP3 * uP) / 2
    fE = (uM * (EM + pM) + uP * (EP + pP)) / 2
    rMs, rPs = np.sqrt(rhoM), np.sqrt(rhoP)
    rhob = rMs * rPs
    u = (rMs * uM + rPs * uP) / (rMs + rPs)
    v = (rMs * vM + rPs * vP) / (rMs + rPs)
    H = (rMs * HM + rPs * HP) / (rMs + rPs)
    c2 = (gamma - 1) * (H - 0.5 * (u * u + v * v))
    c = np.sqrt(np.abs(c2) + 0.0)
    dw1 = (-0.5 * rhob * (uP - uM) / c + 0.5 * (pP - pM) / c2) * np.abs(u - c)
    dw2 = ((rhoP - rhoM) - (pP - pM) / c2) * np.abs(u)
    dw3 = (rhob * (vP - vM)) * np.abs(u)
    dw4 = (0.5 * rhob * (uP - uM) / c + 0.5 * (pP - pM) / c2) * np.abs(u + c)
    fR = fR - (dw1 + dw2 + dw4) / 2
    fU = fU - (dw1 * (u - c) + dw2 * u + dw4 * (u + c)) / 2
    fV = fV - (dw1 * v + dw2 * v + dw3 + dw4 * v) / 2
    fE = fE - (dw1 * (H - u * c) + dw2 * (u * u + v * v) / 2 + dw3 * v + dw4 * (H + u * c)) / 2
    return fR, nx * fU - ny * fV, ny * fU + nx * fV, fE


def rusanov_flux(nx, ny, rhoM, ruM, rvM, EM, rhoP, ruP, rvP, EP, gamma):
    """Point-wise local Lax-Friedrichs (Rusanov) flux F*.n, M = owner, P = neighbour.  NOT a restatement of reference code: the reference's
    godunovScheme knows the Roe flux only (godunovFlux/fluxSchemes/scheme/); this is the Euler counterpart of its scalar LFFlux
    (simpleFlux/schemes/LFFlux/LFFlux.C:105-211) that BASELINE's north_star names, stated here as the checker of the product's HDG_FLUX_LF
    option on the Euler entry points.  Parity unpinned (nothing published)."""
    unM = (nx * ruM + ny * rvM) / rhoM
    unP = (nx * ruP + ny * rvP) / rhoP
    pM = (gamma - 1) * (EM - 0.5 * (ruM * ruM + rvM * rvM) / rhoM)
    pP = (gamma - 1) * (EP - 0.5 * (ruP * ruP + rvP * rvP) / rhoP)
    lam = np.maximum(np.abs(unM) + np.sqrt(np.abs(gamma * pM / rhoM)), np.abs(unP) + np.sqrt(np.abs(gamma * pP / rhoP)))
    fR = 0.5 * ((rhoM * unM + rhoP * unP) - lam * (rhoP - rhoM))
    fU = 0.5 * ((ruM * unM + pM * nx + ruP * unP + pP * nx) - lam * (ruP - ruM))
    fV = 0.5 * ((rvM * unM + pM * ny + rvP * unP + pP * ny) - lam * (rvP - rvM))
    fE = 0.5 * (((EM + pM) * unM + (EP + pP) * unP) - lam * (EP - EM))
    return fR, fU, fV, fE


# --------------------------------------------------------------------------------------------
# 8. Equation assembly + mass solve (defaultConvectionScheme.C:48-129, defaultGrad.C:87-166,
#    EulerDdtScheme.C:118-144, dgLduMatrix.C:316-321, Equation.C:42-79, dgMesh.C:129-172)
# --------------------------------------------------------------------------------------------


def _surface_term(case: Case, flux):
    """b[ownerMap[j]] -= sum_i faceWJ_i flux_i If[i,j];  b[neighborMap[j]] += ... (rotated map)."""
    ref, m, geo = case.ref, case.mesh, case.geo
    t = geo.fWJ.reshape(geo.fWJ.shape + (1,) * (flux.ndim - 2)) * flux          # (F,Nfg,...)
    contrib = np.einsum("ij,fi...->fj...", ref.If, t)                           # (F,Nfp,...)
    b = np.zeros((m.K * ref.Np,) + flux.shape[2:])
    np.subtract.at(b, case.map_o.reshape(-1), contrib.reshape((-1,) + flux.shape[2:]))
    it = case.interior
    np.add.at(b, case.map_n[it].reshape(-1), contrib[it].reshape((-1,) + flux.shape[2:]))
    return b.reshape((m.K, ref.Np) + flux.shape[2:])


def _volume_div(case: Case, Ux, Uy, qg):
    """b[j] += sum_g cellD1dx[g,j] . (U_g q_g WJ_g)"""
    geo = case.geo
    shp = (1,) * (qg.ndim - 2)
    tx = (Ux * geo.WJ).reshape(Ux.shape + shp) * qg
    ty = (Uy * geo.WJ).reshape(Uy.shape + shp) * qg
    return np.einsum("kgj,kg...->kj...", geo.D1x, tx) + np.einsum("kgj,kg...->kj...", geo.D1y, ty)


def _solve(case: Case, q_old, b, dt):
    """b += M source (source = q_old/dt); b *= dt; M q = b  =>  q = q_old + dt M^-1 b_weak.
    Restated with the same operation structure: form M.(q_old/dt), scale, block solve."""
    geo = case.geo
    src = q_old / dt
    btot = b + np.einsum("kij,kj...->ki...", geo.M, src)
    btot = btot * dt
    return np.einsum("kij,kj...->ki...", geo.Minv, btot)


def euler_stage(case: Case, rho, rhoU, E, bR, bU, bE, gamma, dt, flux="Roe"):
    """One forward-Euler sub-step exactly as TUT/isentropicVortex/dgEulerFoam/dgEulerFoam.C:77-90.

    rho (K,Np), rhoU (K,Np,2), E (K,Np); b* = per-patch boundary value lists (in place: evaluated
    afterwards by correctBoundaryConditions, dgMatrixSolve.C:209).
    """
    rc, ro, rn = case.gauss_field(rho, bR)
    uc, uo, un = case.gauss_field(rhoU, bU)
    ec, eo, en = case.gauss_field(E, bE)
    Ux, Uy = uc[..., 0] / rc, uc[..., 1] / rc                                   # gther_U
    p = (gamma - 1.0) * (ec - 0.5 * (rc * (Ux * Ux + Uy * Uy)))                 # gther_p
    nx, ny = case.geo.fnx[..., 0], case.geo.fnx[..., 1]
    fR, fUx, fUy, fE = (roe_flux if flux == "Roe" else rusanov_flux)(nx, ny, ro, uo[..., 0], uo[..., 1], eo, rn, un[..., 0], un[..., 1], en, gamma)
    # rho:  ddt(rho) + div(U, rho, fluxRho)
    b = _volume_div(case, Ux, Uy, rc) + _surface_term(case, fR)
    rho_new = _solve(case, rho, b, dt)
    # rhoU: + div(U, rhoU, fluxRhoU) + grad(p)  [grad flux "none" -> volume only]
    b = _volume_div(case, Ux, Uy, uc) + _surface_term(case, np.stack([fUx, fUy], axis=-1))
    pw = p * case.geo.WJ
    b[..., 0] += np.einsum("kgj,kg->kj", case.geo.D1x, pw)
    b[..., 1] += np.einsum("kgj,kg->kj", case.geo.D1y, pw)
    rhoU_new = _solve(case, rhoU, b, dt)
    # E:   + div(U, E, fluxEner) + div(U, p) [flux "none"]
    b = _volume_div(case, Ux, Uy, ec) + _surface_term(case, fE) + _volume_div(case, Ux, Uy, p)
    E_new = _solve(case, E, b, dt)
    case.evaluate_bc(rho_new, bR)
    case.evaluate_bc(rhoU_new, bU, is_vector=True)
    case.evaluate_bc(E_new, bE)
    return rho_new, rhoU_new, E_new


def lf_flux_nodal(case: Case, Ux, Uy, T, bUx, bUy, bT, kind="LF"):
    """LFFlux::fluxCalculateWeak nodal variant (LFFlux.C:105-211): one maxV per face, normals taken at
    Gauss indices < Nfp, flux formed at the Nfp nodes then interpolated to the Nfg points."""
    ref, m = case.ref, case.mesh
    fl = lambda a: a.reshape(-1)
    To, Uxo, Uyo = fl(T)[case.map_o], fl(Ux)[case.map_o], fl(Uy)[case.map_o]
    Tn, Uxn, Uyn = fl(T)[case.map_n], fl(Ux)[case.map_n], fl(Uy)[case.map_n]
    for ip, p in enumerate(m.patches):
        if p["faces"].size == 0:
            continue
        Tn[p["faces"]] = bT[ip].reshape(-1, ref.Nfp)
        Uxn[p["faces"]] = bUx[ip].reshape(-1, ref.Nfp)
        Uyn[p["faces"]] = bUy[ip].reshape(-1, ref.Nfp)
    nx, ny = case.geo.fnx[:, :ref.Nfp, 0], case.geo.fnx[:, :ref.Nfp, 1]
    vO = nx * Uxo + ny * Uyo
    vN = nx * Uxn + ny * Uyn
    maxV = np.maximum(np.abs(vO), np.abs(vN)).max(axis=1, keepdims=True)
    maxV = np.maximum(maxV, 0.0)
    if kind == "average":                      # averageFlux.C:95-190: the central part only
        maxV = 0.0 * maxV
    elif kind == "none":                       # noneFlux.C:45-97
        return np.zeros((m.F, ref.Nfg))
    f = (vO * To + vN * Tn) * 0.5 + maxV * (To - Tn) * 0.5
    return np.einsum("gi,fi->fg", ref.If, f)


def advect_stage(case: Case, T, Ux, Uy, bT, bUx, bUy, dt, kind="LF"):
    """dg::solveEquation(dgm::ddt(T) + dgc::div(U, T)) with `div(U,T) default LF`:
    EquationConvectionScheme Type3 -> defaultConvectionScheme.C:216-303 (nodal U*T interpolated)."""
    ref, geo = case.ref, case.geo
    flux = lf_flux_nodal(case, Ux, Uy, T, bUx, bUy, bT, kind)
    gx = np.einsum("gj,kj->kg", ref.Vg, Ux * T) * geo.WJ
    gy = np.einsum("gj,kj->kg", ref.Vg, Uy * T) * geo.WJ
    b = np.einsum("kgj,kg->kj", geo.D1x, gx) + np.einsum("kgj,kg->kj", geo.D1y, gy) + _surface_term(case, flux)
    T_new = _solve(case, T, b, dt)
    case.evaluate_bc(T_new, bT)
    return T_new


# --------------------------------------------------------------------------------------------
# 8b. Slope limiter `Triangle` (DG/DG/godunovFlux/limiteSchemes/scheme/Trianglelimite/Trianglelimite.C:61-864)
# --------------------------------------------------------------------------------------------
# Restated for the next row of SURVEY.md §8-f (rank 4); the reference publishes no numbers for a limited run (TUT/doubleMach has no
# tabulated result), so this function is PARITY-UNPINNED: it follows the source line by line and is checked only through the
# properties the algorithm guarantees (tests/test_oracle_limiter.py).  No product kernel uses it yet.


def triangle_limit(case: Case, rho, rhoU, E, bR, bU, bE, gamma=1.4, eps=1e-10, tol=1e-2):
    """Godunov.limite(rho, rhoU, Ener): area-weighted gradient limiter on the primitive variables, P1 reconstruction about the cell
    averages.  Inputs as in euler_stage (nodal fields (K,Np[,2]) and per-patch boundary lists); returns the limited (rho, rhoU, E).
    gamma is hard-wired to 1.4 in the reference (:74); eps = epse (:716), tol (:803)."""
    ref, m, geo = case.ref, case.mesh, case.geo
    K, Np, Nfp = m.K, ref.Np, ref.Nfp
    rhou, rhov = rhoU[..., 0], rhoU[..., 1]
    # 1. cell averages with the column sums of the reference mass matrix / 2 (:109-137); massMatrix = (V V^T)^-1 (baseFunction.C:63-77)
    Mref = np.linalg.inv(ref.V @ ref.V.T)
    mpp = Mref.sum(axis=0) / 2.0
    xr = np.einsum("ij,kjc->kic", ref.Dr, geo.x)
    xs = np.einsum("ij,kjc->kic", ref.Ds, geo.x)
    Jn = xr[..., 0] * xs[..., 1] - xr[..., 1] * xs[..., 0]            # nodal jacobian (triangleBaseFunction.C:330-343)
    A0 = (mpp[None, :] * Jn).sum(1) * 2.0 / 3.0                       # :126-128
    nb = sum(p["faces"].size for p in m.patches if p["faces"].size > 0)
    tot = K + nb                                                     # one virtual cell per boundary face (:88)
    ave = np.zeros((4, tot))
    cx, cy = np.zeros(tot), np.zeros(tot)
    for q, f in enumerate((rho, rhou, rhov, E)):
        ave[q, :K] = f @ mpp
    cx[:K], cy[:K] = geo.x[..., 0] @ mpp, geo.x[..., 1] @ mpp
    # ghost cells, patch by patch in dgFaceIndex order (:153-258)
    ghost_of_face = -np.ones(m.F, dtype=np.int64)
    g = K
    for ip, p in enumerate(m.patches):
        if p["faces"].size == 0:
            continue
        kind = case.bc_kinds[ip]
        for f in p["faces"]:
            o = m.face_owner[f]
            A, B = geo.fnx[f, 0, 0], geo.fnx[f, 0, 1]                # faceNx_[0]
            pt = geo.x[o, ref.f2c[m.face_loc_o[f], 0, 0]]            # first owner face node
            C = -pt[0] * A - pt[1] * B
            cx[g] = (B * B - A * A) * cx[o] - 2 * A * B * cy[o] - 2 * A * C
            cy[g] = (-B * B + A * A) * cy[o] - 2 * A * B * cx[o] - 2 * B * C
            if kind == BC_REFLECTIVE:                                # :176-205: the normal momentum is removed once
                un = A * ave[1, o] + B * ave[2, o]
                ave[:, g] = (ave[0, o], ave[1, o] - A * un, ave[2, o] - B * un, ave[3, o])
            elif kind == BC_FIXED:                                   # :206-233: the FIRST value of the patch field, for every face
                ave[:, g] = (bR[ip][0], bU[ip][0, 0], bU[ip][0, 1], bE[ip][0])
            else:
                ave[:, g] = ave[:, o]
            ghost_of_face[f] = g
            g += 1
    # 2. primitive averages (:294-304)
    prim = np.zeros((4, tot))
    prim[0] = ave[0]
    prim[1], prim[2] = ave[1] / ave[0], ave[2] / ave[0]
    prim[3] = (gamma - 1) * (ave[3] - 0.5 * (ave[1] ** 2 + ave[2] ** 2) / ave[0])
    # 3./4. owner faces in cell order, local face order (:341-452): end-point states, diamond areas, face gradients (:497-560)
    flat = [f.reshape(-1) for f in (rho, rhou, rhov, E)]
    faces = [(k, int(m.cell_face[k, lf])) for k in range(K) for lf in range(3)
             if m.face_owner[m.cell_face[k, lf]] == k and m.face_loc_o[m.cell_face[k, lf]] == lf]
    nF = len(faces)
    V = np.zeros((4, 2, nF))
    A2 = np.zeros(nF)
    nbr_of = np.zeros(nF, dtype=np.int64)
    cellA2 = np.zeros(tot)
    for i, (o, f) in enumerate(faces):
        oS, oE = case.map_o[f, 0], case.map_o[f, Nfp - 1]
        if m.face_nbr[f] >= 0:
            nS, nE = case.map_n[f, 0], case.map_n[f, Nfp - 1]
            nbS = [q[nS] for q in flat]
            nbE = [q[nE] for q in flat]
            n = int(m.face_nbr[f])
            A2[i] = A0[o] + A0[n]
        else:
            ip, off = case.patch_of_face[f], case.patch_off[f]
            nbS = [bR[ip][off], bU[ip][off, 0], bU[ip][off, 1], bE[ip][off]]
            nbE = [bR[ip][off + Nfp - 1], bU[ip][off + Nfp - 1, 0], bU[ip][off + Nfp - 1, 1], bE[ip][off + Nfp - 1]]
            n = int(ghost_of_face[f])
            A2[i] = A0[o] + A0[o]
        nbr_of[i] = n
        S = np.array([0.5 * flat[q][oS] + 0.5 * nbS[q] for q in range(4)])
        Ee = np.array([0.5 * flat[q][oE] + 0.5 * nbE[q] for q in range(4)])
        for st in (S, Ee):                                           # conserved -> (rho, u, v, p) (:432-447)
            st[1], st[2] = st[1] / st[0], st[2] / st[0]
            st[3] = (gamma - 1) * (st[3] - 0.5 * st[0] * (st[1] ** 2 + st[2] ** 2))
        p0, p1 = geo.x.reshape(-1, 2)[oS], geo.x.reshape(-1, 2)[oE]
        Ad = ((cx[n] - cx[o]) * (p1[1] - p0[1]) - (p1[0] - p0[0]) * (cy[n] - cy[o])) * 0.5      # :428
        for q in range(4):
            dc, df = prim[q, n] - prim[q, o], S[q] - Ee[q]
            V[q, 0, i] = 0.5 * (dc * (p1[1] - p0[1]) + df * (cy[n] - cy[o])) / Ad
            V[q, 1, i] = -0.5 * (dc * (p1[0] - p0[0]) + df * (cx[n] - cx[o])) / Ad
        cellA2[o] += A2[i]
        if n < K:
            cellA2[n] += A2[i]
    # 5. cell gradients: A_2-weighted means; a ghost cell takes the gradient of its face (:570-640)
    CV = np.zeros((4, 2, tot))
    for i, (o, f) in enumerate(faces):
        n = nbr_of[i]
        CV[:, :, o] += A2[i] * V[:, :, i] / cellA2[o]
        if n < K:
            CV[:, :, n] += A2[i] * V[:, :, i] / cellA2[n]
        else:
            CV[:, :, n] = V[:, :, i]
    # 6. limited gradient per cell and variable: each neighbour weighted by the squared gradient magnitudes of the other two (:727-798)
    out = [np.empty_like(rho) for _ in range(4)]
    for k in range(K):
        c = []
        for lf in range(3):
            f = m.cell_face[k, lf]
            if m.face_owner[f] == k and m.face_loc_o[f] == lf:
                c.append(int(m.face_nbr[f]) if m.face_nbr[f] >= 0 else int(ghost_of_face[f]))
            else:
                c.append(int(m.face_owner[f]))
        L = np.zeros((4, 2))
        for q in range(4):
            g1, g2, g3 = [CV[q, 0, ci] ** 2 + CV[q, 1, ci] ** 2 for ci in c]
            fac = g1 * g1 + g2 * g2 + g3 * g3
            w = np.array([g2 * g3 + eps, g1 * g3 + eps, g2 * g1 + eps]) / (fac + 3 * eps)
            L[q] = sum(w[i] * CV[q, :, c[i]] for i in range(3))
        # 7. P1 reconstruction about the averages, back to conserved variables (:803-850)
        ub, vb = prim[1, k], prim[2, k]
        for i in range(Np):
            dx, dy = geo.x[k, i, 0] - cx[k], geo.x[k, i, 1] - cy[k]
            du, du1, du2, du3 = (dx * L[q, 0] + dy * L[q, 1] for q in range(4))
            while ave[0, k] + du < tol:                              # "crroect negative density" (:823-827)
                du *= 0.5
            out[0][k, i] = ave[0, k] + du
            out[1][k, i] = ave[1, k] + ave[0, k] * du1 + du * ub
            out[2][k, i] = ave[2, k] + ave[0, k] * du2 + du * vb
            out[3][k, i] = ave[3, k] + du3 / (gamma - 1) + 0.5 * du * (ub * ub + vb * vb) + ave[0, k] * (ub * du1 + vb * du2)
    return out[0], np.stack([out[1], out[2]], axis=-1), out[3]


# --------------------------------------------------------------------------------------------
# 9. Isentropic vortex driver (TUT/isentropicVortex/dgEulerFoam/*)
# --------------------------------------------------------------------------------------------


def vortex_exact(x, y, t, gamma=1.4, beta=5.0):
    """setNonUniformInlet.H:19-27 / setBoundaryValues.H:39-46 / eulerError.H:21-26 (note the y-offset is 0
    and the x-offset 5+t, exactly as written there)."""
    r = (x - 5.0 - t) ** 2 + y ** 2
    rho = np.power(1.0 - (gamma - 1.0) * (beta * beta) * np.exp(2.0 * (1.0 - r)) / (16.0 * gamma * math.pi * math.pi), 1.0 / (gamma - 1.0))
    ru = (1 - beta * np.exp(1 - r) * (y - 0) / (2.0 * math.pi)) * rho
    rv = (beta * np.exp(1 - r) * (x - 5 - t) / (2.0 * math.pi)) * rho
    E = np.power(rho, gamma) / (gamma - 1.0) + 0.5 * (ru * ru + rv * rv) / rho
    return rho, ru, rv, E


class VortexRun:
    """The main loop of dgEulerFoam.C:64-131 (SSP-RK2 from two forward-Euler solves; stage-2 boundary
    data stale at t_n)."""

    def __init__(self, case: Case, dt, gamma=1.4):
        self.case, self.dt, self.gamma, self.t = case, dt, gamma, 0.0
        x, y = case.geo.x[..., 0], case.geo.x[..., 1]
        self.rho, ru, rv, self.E = vortex_exact(x, y, 0.0, gamma)
        self.rhoU = np.stack([ru, rv], axis=-1)
        npatch = len(case.mesh.patches)
        self.bR = [case.patch_internal(self.rho, ip) for ip in range(npatch)]
        self.bU = [case.patch_internal(self.rhoU, ip) for ip in range(npatch)]
        self.bE = [case.patch_internal(self.E, ip) for ip in range(npatch)]

    def set_boundary_values(self, t):
        case = self.case
        for ip, kind in enumerate(case.bc_kinds):
            if kind != BC_FIXED or case.mesh.patches[ip]["faces"].size == 0:
                continue
            xy = case.patch_internal(case.geo.x, ip)
            r, ru, rv, e = vortex_exact(xy[:, 0], xy[:, 1], t, self.gamma)
            self.bR[ip], self.bU[ip], self.bE[ip] = r, np.stack([ru, rv], axis=-1), e

    def step(self):
        c, dt, g = self.case, self.dt, self.gamma
        self.set_boundary_values(self.t)                      # runTime - deltaT after runTime++
        r1, u1, e1 = euler_stage(c, self.rho, self.rhoU, self.E, self.bR, self.bU, self.bE, g, dt)
        r2, u2, e2 = euler_stage(c, r1, u1, e1, self.bR, self.bU, self.bE, g, dt)
        self.rho = 0.5 * self.rho + 0.5 * r2
        self.rhoU = 0.5 * self.rhoU + 0.5 * u2
        self.E = 0.5 * self.E + 0.5 * e2
        c.evaluate_bc(self.rho, self.bR)
        c.evaluate_bc(self.rhoU, self.bU, is_vector=True)
        c.evaluate_bc(self.E, self.bE)
        self.t += dt

    def errors(self):
        """eulerError.H:32-38: sum|rho-rho_ex|/nDof ; sum|rhoU-rhoU_ex| (vector magnitude)/nDof."""
        x, y = self.case.geo.x[..., 0], self.case.geo.x[..., 1]
        r, ru, rv, _ = vortex_exact(x, y, self.t, self.gamma)
        ndof = self.rho.size
        e_r = np.abs(r - self.rho).sum() / ndof
        e_u = np.sqrt((ru - self.rhoU[..., 0]) ** 2 + (rv - self.rhoU[..., 1]) ** 2).sum() / ndof
        return e_r, e_u


def doublemach_exact(x, y, t, gamma=1.4):
    """Post-/pre-shock states of the double Mach reflection (TUT/doubleMach/dgEulerFoam/setNonUniformInlet.H:19-40,
    setBoundaryValues.H:17-58): a Mach-10 shock through (1/6, 0) at 60 degrees, moving with 20 t."""
    g = (1.0 + 20.0 * t) / math.sqrt(3.0)
    left = (x - 1.0 / 6.0) / g - y < 0
    rho = np.where(left, 8.0, 1.4)
    ru = np.where(left, 8.25 * math.cos(math.pi / 6.0) * 8.0, 0.0)
    rv = np.where(left, -8.25 * math.sin(math.pi / 6.0) * 8.0, 0.0)
    E = np.where(left, 116.5, 1.0) / (gamma - 1.0) + (ru * ru + rv * rv) / (2.0 * rho)
    return rho, ru, rv, E


class DoubleMachRun:
    """The main loop of TUT/doubleMach/dgEulerFoam/dgEulerFoam.C:60-123: the vortex loop with Godunov.limite after each stage pair.
    Two sets of boundary data exist, as in the reference: the work fields rho1/rhoU1/Ener1 get the moving-shock state on every
    patch not named `wall` (setBoundaryValues.H:27) at t_n; the fields rho/rhoU/Ener keep the fixedValue data they were given at
    start-up (assignment to a fixedValue patch field is a no-op, fixedValueDgPatchField.H:180-194) - `b0`, by default the interior
    trace of the initial state (setNonUniformInlet.H:43-50).  Patch fields that are evaluated from the interior (reflective,
    zeroGradient) are copied by `rho1 = rho` and re-evaluated only by correctBoundaryConditions (after each solve, dgMatrixSolve.C:209,
    and at :121-123); Godunov.limite changes the interior WITHOUT re-evaluating them, so in the reference the second stage sees the wall
    data of the UNLIMITED stage-1 field.  refresh_after_limit=True evaluates them from the limited field instead, which is what a
    kernel that mirrors the current trace does."""

    def __init__(self, case: Case, dt, gamma=1.4, b0=None, refresh_after_limit=False):
        self.case, self.dt, self.gamma, self.t, self.refresh = case, dt, gamma, 0.0, refresh_after_limit
        x, y = case.geo.x[..., 0], case.geo.x[..., 1]
        self.rho, ru, rv, self.E = doublemach_exact(x, y, 0.0, gamma)
        self.rhoU = np.stack([ru, rv], axis=-1)
        npatch = len(case.mesh.patches)
        self.b0 = b0 if b0 is not None else [[case.patch_internal(f, ip) for ip in range(npatch)] for f in (self.rho, self.rhoU, self.E)]
        self._evaluate(self.b0, self.rho, self.rhoU, self.E)                                   # setNonUniformInlet.H:52-54
        self.b1 = [[v.copy() for v in q] for q in self.b0]                                     # dgScalarField rho1("rho1", rho)

    def boundary_state(self, t):
        case = self.case
        out = [[], [], []]
        for ip in range(len(case.mesh.patches)):
            xy = case.patch_internal(case.geo.x, ip)
            r, ru, rv, e = doublemach_exact(xy[:, 0], xy[:, 1], t, self.gamma)
            out[0].append(r)
            out[1].append(np.stack([ru, rv], axis=-1))
            out[2].append(e)
        return out

    def _evaluate(self, b, rho, rhoU, E):
        c = self.case
        c.evaluate_bc(rho, b[0])
        c.evaluate_bc(rhoU, b[1], is_vector=True)
        c.evaluate_bc(E, b[2])

    def step(self):
        c, dt, g = self.case, self.dt, self.gamma
        ex = self.boundary_state(self.t)                       # setBoundaryValues(rho1, rhoU1, Ener1, gamma, runTime - deltaT)
        for ip, kind in enumerate(c.bc_kinds):
            for q in range(3):
                if kind != BC_FIXED:
                    self.b1[q][ip] = self.b0[q][ip].copy()     # rho1 = rho copies the patch fields that do not fix their value
                elif c.mesh.patches[ip]["name"] != "wall":
                    self.b1[q][ip] = ex[q][ip]
        r1, u1, e1 = euler_stage(c, self.rho, self.rhoU, self.E, *self.b1, g, dt)              # evaluates b1 from the new fields
        r1, u1, e1 = triangle_limit(c, r1, u1, e1, *self.b1, gamma=1.4)                        # :92 (gamma hard-wired in the limiter)
        if self.refresh:
            self._evaluate(self.b1, r1, u1, e1)
        r2, u2, e2 = euler_stage(c, r1, u1, e1, *self.b1, g, dt)
        for ip, kind in enumerate(c.bc_kinds):                 # rho = 0.5*rho + 0.5*rho1, boundary field included (fixedValue: no-op)
            if kind != BC_FIXED:
                for q in range(3):
                    self.b0[q][ip] = 0.5 * self.b0[q][ip] + 0.5 * self.b1[q][ip]
        self.rho = 0.5 * self.rho + 0.5 * r2
        self.rhoU = 0.5 * self.rhoU + 0.5 * u2
        self.E = 0.5 * self.E + 0.5 * e2
        self.rho, self.rhoU, self.E = triangle_limit(c, self.rho, self.rhoU, self.E, *self.b0, gamma=1.4)      # :119
        self._evaluate(self.b0, self.rho, self.rhoU, self.E)                                   # correctBoundaryConditions, :121-123
        self.t += dt


# --------------------------------------------------------------------------------------------
# 10. Domain decomposition (dgDecomposePar): `simple` method + processor-mesh maps
#     src/parallel/decompose/decompositionMethods/simpleGeomDecomp/simpleGeomDecomp.C:55-84,129-197
#     src/parallel/decompose/decompositionMethods/geomDecomp/geomDecomp.C:53-64
#     applications/utilities/DG/dgDecomposePar/domainDecompositionMesh.C:102-511
# --------------------------------------------------------------------------------------------


def simple_decomp(mesh: DGMesh, n, delta=0.001):
    """cellToProc of `method simple; simpleCoeffs{n (nx ny nz); delta}` on the cell centres (triangle centroids)."""
    d = 1 - 0.5 * delta * delta
    a = delta
    R = np.array([[d * d, -a * d, a], [a * d - a * a * d, a * a * a + d * d, -2 * a * d], [a * d * d + a * a, a * d - a * a * d, d * d - a * a]])
    v = mesh.xy[mesh.tris]
    cx = (v[:, 0, 0] + v[:, 1, 0] + v[:, 2, 0]) / 3.0
    cy = (v[:, 0, 1] + v[:, 1, 1] + v[:, 2, 1]) / 3.0
    K = mesh.K
    final = np.zeros(K, dtype=np.int64)
    mult = 1
    for direction in range(3):
        coord = R[direction, 0] * cx + R[direction, 1] * cy + R[direction, 2] * 0.0
        idx = np.argsort(coord, kind="stable")
        ng = int(n[direction])
        jump = K // ng
        fst = K - jump * ng
        group = np.concatenate([np.repeat(np.arange(fst), jump + 1), np.repeat(np.arange(fst, ng), jump)])
        final[idx] += mult * group
        mult *= ng
    return final.astype(np.int32)


def decompose(mesh: DGMesh, cell_to_proc, nprocs, rank, poly_face=None):
    """Processor mesh maps of one rank.  Returns dict(cell, point, tris (local ids), patches=[(name, nbrProc, [global dgFace ids])]).
    poly_face: optional polyMesh id of every dgFace (orders the cut faces); default = upper-triangular (owner, neighbour) rank."""
    c2p = np.asarray(cell_to_proc)
    cell = np.nonzero(c2p == rank)[0].astype(np.int32)                        # ascending (invertOneToMany, :124)
    used = np.zeros(mesh.xy.shape[0], dtype=bool)
    used[mesh.tris[cell].reshape(-1)] = True
    point = np.nonzero(used)[0].astype(np.int32)                              # ascending (:463-511)
    g2l = -np.ones(mesh.xy.shape[0], dtype=np.int64)
    g2l[point] = np.arange(point.size)
    patches = []
    for p in mesh.patches:                                                    # original patches, faces where the cell lives (:160-185)
        faces = [int(f) for f in p["faces"] if c2p[mesh.face_owner[f]] == rank]
        patches.append((p["name"], -1, faces))
    cuts = {}
    for f in range(mesh.F):
        nb = mesh.face_nbr[f]
        if nb < 0:
            continue
        po, pn = c2p[mesh.face_owner[f]], c2p[nb]
        if po == pn or (po != rank and pn != rank):
            continue
        key = int(poly_face[f]) if poly_face is not None else int(mesh.face_owner[f]) * mesh.K + int(nb)
        cuts.setdefault(int(pn if po == rank else po), []).append((key, f))
    for q in sorted(cuts):                                                    # ascending neighbour processor (:355-372)
        patches.append((f"procBoundary{rank}to{q}", q, [f for _, f in sorted(cuts[q])]))    # ascending global face id (:215-240)
    return {"cell": cell, "point": point, "tris": g2l[mesh.tris[cell]].astype(np.int32), "patches": patches}


def decompose_polymesh(pm, cell_to_proc, rank):
    """The processor polyMesh of one rank as dgDecomposePar cuts it out of the global polyMesh (pm = read_polymesh(...)):
    domainDecompositionMesh.C:124 (cells ascending), :132-143 (internal faces, no turning index), :160-185 (patch faces where the cell
    lives), :53-99 + :355-411 (cut faces per neighbour in ascending rank, +(f+1) on the owner side, -(f+1) on the neighbour side),
    :463-511 (points ascending), domainDecomposition.C:270-330 (faces copied, reversed when the index is negative).
    Returns dict(cell, point, face (= faceProcAddressing), faces (local point lists), owner, neighbour, patches [(name, n, start)])."""
    c2p = np.asarray(cell_to_proc)
    owner, neigh, faces = pm["owner"], pm["neighbour"], pm["faces"]
    n_int = neigh.size
    cell = np.nonzero(c2p == rank)[0]
    face_addr = [f + 1 for f in range(n_int) if c2p[owner[f]] == rank and c2p[neigh[f]] == rank]
    n_proc_int = len(face_addr)
    patches = []
    for p in pm["patches"]:
        start = len(face_addr)
        face_addr += [f + 1 for f in range(p["startFace"], p["startFace"] + p["nFaces"]) if c2p[owner[f]] == rank]
        patches.append((p["name"], len(face_addr) - start, start))
    cuts = {}
    for f in range(n_int):
        po, pn = int(c2p[owner[f]]), int(c2p[neigh[f]])
        if po != pn:
            if po == rank:
                cuts.setdefault(pn, []).append(f + 1)
            elif pn == rank:
                cuts.setdefault(po, []).append(-(f + 1))
    for q in sorted(cuts):
        patches.append((f"procBoundary{rank}to{q}", len(cuts[q]), len(face_addr)))
        face_addr += cuts[q]
    used = np.zeros(pm["points"].shape[0], dtype=bool)
    for fa in face_addr:
        used[faces[abs(fa) - 1]] = True
    point = np.nonzero(used)[0]
    lookup = -np.ones(used.size, dtype=np.int64)
    lookup[point] = np.arange(point.size)
    clook = -np.ones(c2p.size, dtype=np.int64)
    clook[cell] = np.arange(cell.size)
    lfaces, lown, lnei = [], [], []
    for i, fa in enumerate(face_addr):
        f = abs(fa) - 1
        pts = list(faces[f])
        if fa < 0:
            pts = [pts[0]] + pts[:0:-1]                      # face::reverseFace keeps the first point
        lfaces.append([int(lookup[q]) for q in pts])
        lown.append(int(clook[owner[f] if fa > 0 else neigh[f]]))
        if i < n_proc_int:
            lnei.append(int(clook[neigh[f]]))
    return {"cell": cell, "point": point, "face": np.array(face_addr), "faces": lfaces, "owner": np.array(lown), "neighbour": np.array(lnei),
            "patches": patches}


# --------------------------------------------------------------------------------------------
# 11. Curved (`arc`) boundary patches
#     dgMesh/dgPatches/constraint/arc/arcDgPatch.C:361-540 (closest point of the parametric curve, end points kept)
#     element/physicalElementData/physicalElementData.C:185-224 (displacement of the patch-face nodes blended into the owner cell)
#     element/baseFunctions/straightBaseFunctions/triangleBaseFunction/triangleBaseFunction.C:421-466 (addFaceShiftToCell)
#     NOTE dgMesh.C:110-113: initElements (all metrics, mass matrices, cellD1dx, face normals) runs BEFORE the displacement and nothing
#     recomputes them: in the reference a curved patch moves dofLocation (where fields / boundary values are sampled), not the operators.
# --------------------------------------------------------------------------------------------


def add_face_shift_to_cell(ref: RefElement, face: int, shift):
    """Displacement of all Np cell nodes caused by the displacement `shift` (Nfp,2) of the nodes of local face `face` (in the face's own
    traversal order): triangleBaseFunction::addFaceShiftToCell."""
    shift = np.asarray(shift, dtype=float)
    if face == 2:
        shift = shift[::-1]                                            # :424-431
    vr = ref.r if face == 0 else ref.s                                 # :434-440
    lgl = jacobi_gl(0, 0, ref.N)
    inv_v1 = np.linalg.inv(vandermonde1d(ref.N, lgl))                  # invFaceMatrix_ = base_1D->invV_ (:61-63)
    d = vandermonde1d(ref.N, vr) @ (inv_v1 @ shift)                    # :443-446
    for p in range(ref.Np):
        if abs(1.0 - vr[p]) < 1e-7:                                    # :451-452
            continue
        blend = (ref.r[p] + 1) / (1 - vr[p]) if face == 1 else -(ref.r[p] + ref.s[p]) / (1 - vr[p])
        d[p] *= blend
    return d


def arc_closest_point(curve, u_range, p, tol=1e-14):
    """arcDgPatch::position: the point of the curve u -> curve(u) = (x, y) closest to p (orthogonality (c - p).c' = 0).  The reference
    iterates a damped secant method with finite-difference derivatives to 1e-12 (getShortestPoint, :361-470); any root finder gives the
    same point - here: coarse scan + bisection/secant on the orthogonality function."""
    us = np.linspace(u_range[0], u_range[1], 2001)
    pts = np.array([curve(u) for u in us])
    i = int(np.argmin(((pts - np.asarray(p)) ** 2).sum(1)))
    h = 1e-6 * abs(u_range[1] - u_range[0])
    g = lambda u: float(np.dot(np.asarray(curve(u)) - p, (np.asarray(curve(u + h)) - np.asarray(curve(u - h))) / (2 * h)))
    a, b = us[max(i - 1, 0)], us[min(i + 1, us.size - 1)]
    ga, gb = g(a), g(b)
    if ga * gb > 0:
        return np.asarray(curve(us[i]))
    for _ in range(200):
        m = 0.5 * (a + b)
        gm = g(m)
        if ga * gm <= 0:
            b, gb = m, gm
        else:
            a, ga = m, gm
        if abs(b - a) < tol * max(1.0, abs(a)):
            break
    return np.asarray(curve(0.5 * (a + b)))


def apply_arc_patch(case, ip, curve, u_range):
    """physicalElementData::updatePatchDofIndexMapping for one curved patch: returns (positions of the patch-face nodes on the curve,
    displaced dofLocation of all cells).  case.geo.x itself is left alone (the metrics stay straight-sided, see the note above)."""
    ref, m = case.ref, case.mesh
    x = case.geo.x.copy()
    faces = m.patches[ip]["faces"]
    std = case.patch_internal(case.geo.x, ip).reshape(faces.size, ref.Nfp, 2)
    pos = std.copy()
    for f in range(faces.size):
        for i in range(1, ref.Nfp - 1):                                # end points stay (isEnd, arcDgPatch.C:527-535)
            pos[f, i] = arc_closest_point(curve, u_range, std[f, i])
    for f, fid in enumerate(faces):
        x[m.face_owner[fid]] += add_face_shift_to_cell(ref, int(m.face_loc_o[fid]), pos[f] - std[f])
    return pos.reshape(-1, 2), x

