"""tools/fluentMeshToCase.py: Fluent 2-D triangle mesh -> constant/polyMesh (the step the reference leaves to fluentMeshToFoam; the
doubleMach tutorial ships only doubleMach.msh).  Checked on a generated mesh (round trip through the product's polyMesh reader) and, in
this container, on the reference's doubleMach.msh - where the product's limiter core is also held to the oracle on the tutorial's own
mesh, initial state and boundary set."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from hopefoam_b200 import meshgen
from oracle import dg_oracle as o
from tests import helpers as H
from tests import test_limiter_core_host as T
from tests.test_limiter_core_host import harness  # noqa: F401  (fixture)

ROOT = Path(__file__).resolve().parent.parent
TOOL = ROOT / "tools" / "fluentMeshToCase.py"
REF_MSH = Path("/root/reference/HopeFOAM-0.1/tutorials/DG/2D/doubleMach/doubleMach.msh")


def _write_fluent(path, xy, tris, sides, names):
    """GAMBIT-style writer for the test: one face zone per boundary side + one interior zone."""
    edge = {}
    for c, t in enumerate(tris):
        for f in range(3):
            a, b = int(t[f]), int(t[(f + 1) % 3])
            edge.setdefault((min(a, b), max(a, b)), []).append((c, a, b))
    bzone = {}
    for z, sd in enumerate(sides):
        for c, a, b in sd:
            bzone[(min(a, b), max(a, b))] = z
    zones = [[] for _ in sides]
    interior = []
    for key, lst in edge.items():
        if len(lst) == 2:
            (c0, a, b), (c1, _, _) = lst
            interior.append(f"2 {a + 1:x} {b + 1:x} {c0 + 1:x} {c1 + 1:x}")
        else:
            c0, a, b = lst[0]
            zones[bzone[key]].append(f"2 {a + 1:x} {b + 1:x} {c0 + 1:x} 0")
    out = ['(0 "GAMBIT to Fluent File")', "", '(0 "Dimension:")', "(2 2)", "", f"(10 (0 1 {len(xy):x} 1 2))", f"(10 (1 1 {len(xy):x} 1 2)("]
    out += [f"  {x:.16e}  {y:.16e}" for x, y in xy] + ["))", "", '(0 "Faces:")']
    nf = len(interior) + sum(len(z) for z in zones)
    out.append(f"(13(0 1 {nf:x} 0))")
    start = 1
    for z, rows in enumerate(zones):
        out += [f"(13({z + 3:x} {start:x} {start + len(rows) - 1:x}  3 0)("] + rows + ["))"]
        start += len(rows)
    out += [f"(13({len(zones) + 4:x} {start:x} {start + len(interior) - 1:x} 2 0)("] + interior + ["))", "", '(0 "Cells:")']
    out += [f"(12 (0 1 {len(tris):x} 0))", f"(12 (2 1 {len(tris):x} 1 1))", "", '(0 "Zones:")', "(45 (2 fluid fluid)())"]
    out += [f"(45 ({z + 3:x} wall {n})())" for z, n in enumerate(names)] + [f"(45 ({len(zones) + 4:x} interior default-interior)())", ""]
    Path(path).write_text("\n".join(out))


def _areas(xy, tris):
    a, b, c = xy[tris[:, 0]], xy[tris[:, 1]], xy[tris[:, 2]]
    return 0.5 * ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]))


def test_generated_mesh_round_trip(built_library, tmp_path):
    n = 6
    mg, _ = T.multi_patch_mesh(n)
    names = ["wall", "outlet", "far", "inlet"]
    _write_fluent(tmp_path / "m.msh", mg["xy"], mg["tris"], mg["patch_edges"], names)
    r = subprocess.run([sys.executable, str(TOOL), str(tmp_path / "m.msh"), str(tmp_path / "case")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert f"{(n + 1) ** 2} points, {2 * n * n} triangles" in r.stdout
    c = H.HostContext()
    c.set_order(2)
    c.set_mesh_polymesh(str(tmp_path / "case" / "constant" / "polyMesh"))
    assert c.K == 2 * n * n and c.n_ghost == 4 * n
    info = [c.patch_info(p) for p in range(c.n_patches)]
    assert [(i[0], i[1], i[2]) for i in info][:4] == [("wall", "wall", n), ("outlet", "patch", n), ("far", "patch", n), ("inlet", "patch", n)]
    # same triangles (as vertex-coordinate sets) in the same cell order, all counter-clockwise
    pts = T.ctx_points(c)
    tv = c.cell_vertices()
    assert (_areas(pts, tv) > 0).all()
    for k in range(c.K):
        got = sorted(map(tuple, np.round(pts[tv[k]], 8)))
        want = sorted(map(tuple, np.round(mg["xy"][mg["tris"][k]], 8)))
        assert got == want
    # every patch face lies on its side of the square
    xyp = [c.patch_node_coords(p) for p in range(4)]
    assert np.abs(xyp[0][:, 1] + 5).max() < 1e-12 and np.abs(xyp[1][:, 0] - 10).max() < 1e-12
    assert np.abs(xyp[2][:, 1] - 5).max() < 1e-12 and np.abs(xyp[3][:, 0]).max() < 1e-12


def test_rejects_non_triangle_input(tmp_path):
    (tmp_path / "bad.msh").write_text('(0 "x")\n(2 3)\n')
    r = subprocess.run([sys.executable, str(TOOL), str(tmp_path / "bad.msh"), str(tmp_path / "case")], capture_output=True, text=True)
    assert r.returncode != 0 and "2-D" in r.stderr


@pytest.mark.skipif(not REF_MSH.exists(), reason="reference tutorial mesh not present (GPU box)")
def test_doublemach_tutorial_mesh_and_limiter(built_library, harness, tmp_path):  # noqa: F811
    r = subprocess.run([sys.executable, str(TOOL), str(REF_MSH), str(tmp_path / "case")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "2795 points, 5390 triangles, patches: far (41), wall (101), outlet (32), inlet (24)" in r.stdout
    N = 1
    c = H.HostContext()
    c.set_order(N)
    c.set_mesh_polymesh(str(tmp_path / "case" / "constant" / "polyMesh"))
    pts, tv = T.ctx_points(c), c.cell_vertices()
    assert c.K == 5390 and abs(_areas(pts, tv).sum() - 3.2) < 1e-4          # [0, 3.2] x [0, 1]
    wall = c.patch_node_coords(1)
    assert np.abs(wall[:, 1]).max() < 1e-12 and wall[:, 0].min() > 1 / 6 - 1e-4   # the reflecting wall starts at x = 1/6
    # the tutorial's boundary set (0/rho: wall reflective, far/inlet/outlet fixedValue) and initial state (setNonUniformInlet.H:19-40)
    om = o.mesh_from_polymesh(tmp_path / "case" / "constant" / "polyMesh")
    names = [p["name"] for p in om.patches]
    kinds = [o.BC_REFLECTIVE if nm == "wall" else o.BC_FIXED for nm in names if nm != "frontAndBackPlanes"]
    kinds += [o.BC_ZEROGRAD] * (len(om.patches) - len(kinds))
    case = o.Case(om, N, bc_kinds=kinds)

    def state(x, y, t=0.0):
        g = (1 + 20 * t) / np.sqrt(3.0)
        left = (x - 1.0 / 6.0) / g - y < 0
        rho = np.where(left, 8.0, 1.4)
        ru = np.where(left, 8.25 * np.cos(np.pi / 6) * 8.0, 0.0)
        rv = np.where(left, -8.25 * np.sin(np.pi / 6) * 8.0, 0.0)
        E = np.where(left, 116.5, 1.0) / 0.4 + (ru ** 2 + rv ** 2) / (2 * rho)
        return rho, np.stack([ru, rv], -1), E

    rho, U, E = state(case.geo.x[..., 0], case.geo.x[..., 1])
    bR, bU, bE = [], [], []
    for ip in range(len(om.patches)):
        if om.patches[ip]["faces"].size == 0:
            bR.append(np.zeros(0)); bU.append(np.zeros((0, 2))); bE.append(np.zeros(0))
            continue
        xp = case.patch_xy(ip) if hasattr(case, "patch_xy") else c.patch_node_coords(ip)
        r_, u_, e_ = state(xp[:, 0], xp[:, 1])
        bR.append(r_); bU.append(u_); bE.append(e_)
    case.evaluate_bc(rho, bR)
    case.evaluate_bc(U, bU, is_vector=True)
    case.evaluate_bc(E, bE)
    want = o.triangle_limit(case, rho, U, E, bR, bU, bE)
    got = T.run_product_core(harness, c, (rho, U, E), (bR, bU, bE), kinds)
    for g_, w_ in zip(got, want):
        assert np.isfinite(g_).all()
        assert np.abs(g_ - w_).max() <= 1e-10 * np.abs(w_).max()
    assert got[0].min() > 1.0 and got[0].max() < 9.0
