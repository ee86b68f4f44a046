"""Curved `arc` boundary patches: arcDgPatch::positions + physicalElementData::updatePatchDofIndexMapping +
triangleBaseFunction::addFaceShiftToCell, product (C ABI + facade expression interpreter) against the oracle's restatement.

What the reference does with a curved patch (and therefore what is reproduced): the interior nodes of the patch faces move to the
closest point of the parametric curve, the displacement is blended into the owner cell's dofLocation - AFTER initElements has built every
metric, mass matrix and face normal from the straight-sided nodes (dgMesh.C:110-113), and nothing rebuilds them.  So `arc` changes
where fields and boundary values are sampled and written, not the operators.  Parity unpinned (no published number, reference not
runnable here)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from hopefoam_b200 import meshgen
from oracle import dg_oracle as o
from tests import helpers as H
from tests.case_writer import write_euler_case

ROOT = Path(__file__).resolve().parent.parent
APP = ROOT / "hopefoam_b200" / "apps" / "bin" / "hopeDgToVTK"
R0 = 0.5


def _annulus(n_r=3, n_th=12):
    mg = meshgen.ogrid_sector(n_r, n_th, 0.0, 2 * np.pi, r0=R0, r1=3.0, closed=True)
    mg["patch_edges"] = [mg["sides"]["left"], mg["sides"]["right"]]       # 0 = cylinder wall (chords of the circle r = R0), 1 = far field
    return mg


@pytest.mark.parametrize("N", [2, 4, 6])
def test_curved_wall_matches_oracle(built_library, N):
    mg = _annulus()
    om = H.oracle_mesh(mg)
    case = o.Case(om, N)
    circle = lambda u: (R0 * np.cos(u), R0 * np.sin(u))
    pos, want = o.apply_arc_patch(case, 0, circle, (-0.1, 2 * np.pi + 0.1))
    c = H.HostContext()
    c.set_order(N)
    c.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    straight = c.node_coords()
    assert np.abs(straight - case.geo.x).max() < 1e-13
    c.set_curved_patch(0, pos)
    got = c.node_coords()
    assert np.abs(got - want).max() <= 1e-13
    # the wall nodes now lie on the circle; the chords did not (sagitta of a 30-degree chord: R0 (1 - cos 15 deg) = 0.017)
    wall = c.patch_node_coords(0)
    assert np.abs(np.hypot(wall[:, 0], wall[:, 1]) - R0).max() < 1e-12
    sw = case.patch_internal(case.geo.x, 0)
    assert (R0 - np.hypot(sw[:, 0], sw[:, 1])).max() > 0.01
    # only the cells on the wall moved, their far vertex and their two other faces did not; the far-field patch is untouched
    moved = np.abs(got - straight).reshape(c.K, -1).max(1) > 0
    owners = set(int(k) for k in c.faces()["owner"][c.patch_faces(0)])
    assert set(np.nonzero(moved)[0].tolist()) == owners
    assert np.abs(c.patch_node_coords(1) - case.patch_internal(case.geo.x, 1)).max() < 1e-13
    f2c, fc = c.face_to_cell_index(), c.faces()
    for fid in c.patch_faces(0):
        k, lf = int(fc["owner"][fid]), int(fc["loc_o"][fid])
        for other in range(3):
            if other != lf:
                nodes = f2c[other, 0]
                assert np.abs(got[k, nodes] - straight[k, nodes]).max() < 1e-14      # the blend vanishes on the other faces


def test_face_shift_operator_matches_oracle(built_library):
    for N in (1, 3, 5, 8):
        c = H.HostContext()
        c.set_order(N)
        ref = o.RefElement(N)
        B = c.operator("faceShift", (3, ref.Np, ref.Nfp))
        rng = np.random.default_rng(N)
        for f in range(3):
            sh = rng.standard_normal((ref.Nfp, 2))
            assert np.abs(B[f] @ sh - o.add_face_shift_to_cell(ref, f, sh)).max() < 1e-12


def test_facade_reads_arc_patch_code(tmp_path, built_library):
    """polyMesh/boundary `type arc` with the tutorial-style run-time code entry (TUT/cylinder/constant/polyMesh/boundary:45-80): the
    facade interprets the expression, projects the wall nodes and hands them to the library; hopeDgToVTK writes the displaced nodes."""
    if not APP.exists():
        subprocess.run(["make", "-C", str(ROOT / "hopefoam_b200" / "csrc"), "apps"], check=True)
    N = 4
    mg = _annulus()
    wall, far = mg["patch_edges"]
    # the O-grid is closed through point_equiv, which a polyMesh cannot express: cut it open into a sector mesh with two straight sides
    mg = meshgen.ogrid_sector(3, 8, 0.0, 1.5 * np.pi, r0=R0, r1=3.0)
    patches = [("cylinder", "arc", mg["sides"]["left"]), ("far", "patch", np.concatenate([mg["sides"]["right"], mg["sides"]["bottom"], mg["sides"]["top"]]))]
    case = write_euler_case(tmp_path / "acase", mg, N, 1e-3, 1e-3, patches=patches)
    b = case / "constant" / "polyMesh" / "boundary"
    code = ("        name            codecyl;\n        u_Range         (-0.1 4.9);\n        v_Range         (0 0);\n        code\n        #{\n"
            f"            {R0}*Foam::cos(u),\n            {R0}*Foam::sin(Foam::constant::mathematical::pi*u/Foam::constant::mathematical::pi),\n            v\n        #}};\n")
    t = b.read_text()
    i = t.index("type            arc;") + len("type            arc;\n")
    b.write_text(t[:i] + code + t[i:])
    out = subprocess.run([str(APP), "-case", str(case), "-time", "0", "rho"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    vtk = (case / "VTK" / "acase_0.vtk").read_text().split("\n")
    K, Np = mg["tris"].shape[0], (N + 1) * (N + 2) // 2
    pts = np.array([l.split() for l in vtk[5:5 + K * Np]], dtype=float)[:, :2].reshape(K, Np, 2)
    om = o.mesh_from_polymesh(case / "constant" / "polyMesh")      # the same file the facade reads (the writer rotates the cell vertices)
    assert [p["name"] for p in om.patches][:2] == ["cylinder", "far"]
    ocase = o.Case(om, N)
    _, want = o.apply_arc_patch(ocase, 0, lambda u: (R0 * np.cos(u), R0 * np.sin(u)), (-0.1, 4.9))
    assert np.abs(pts - want).max() < 1e-10                      # 12 printed digits
    assert np.abs(want - ocase.geo.x).max() > 1e-3


def test_arc_expression_errors(tmp_path, built_library):
    """An expression outside the interpreted subset fails loudly (the reference would fail in its run-time compilation)."""
    mg = meshgen.ogrid_sector(2, 4, 0.0, np.pi, r0=R0, r1=2.0)
    patches = [("cylinder", "arc", mg["sides"]["left"]), ("far", "patch", np.concatenate([mg["sides"]["right"], mg["sides"]["bottom"], mg["sides"]["top"]]))]
    case = write_euler_case(tmp_path / "bcase", mg, 2, 1e-3, 1e-3, patches=patches)
    b = case / "constant" / "polyMesh" / "boundary"
    t = b.read_text()
    i = t.index("type            arc;") + len("type            arc;\n")
    b.write_text(t[:i] + "        u_Range (0 3.2);\n        code\n        #{\n            myFunction(u), 0.5*Foam::sin(u), v\n        #};\n" + t[i:])
    out = subprocess.run([str(APP), "-case", str(case), "-time", "0", "rho"], capture_output=True, text=True, timeout=120)
    assert out.returncode != 0 and "unknown function myFunction" in (out.stdout + out.stderr)
