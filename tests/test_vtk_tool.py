"""hopeDgToVTK (host-only post-processing tool on the facade): sub-triangulated legacy VTK of a time directory - runs without a GPU."""
import subprocess
from pathlib import Path

import numpy as np

from hopefoam_b200 import meshgen
from tests.case_writer import write_euler_case

ROOT = Path(__file__).resolve().parent.parent
APP = ROOT / "hopefoam_b200" / "apps" / "bin" / "hopeDgToVTK"


def test_vtk_export_of_initial_fields(tmp_path, built_library):
    if not APP.exists():
        subprocess.run(["make", "-C", str(ROOT / "hopefoam_b200" / "csrc"), "apps"], check=True)
    N = 3
    mg = meshgen.jittered_square(4)
    case = write_euler_case(tmp_path / "vcase", mg, N, 1e-3, 1e-3)
    out = subprocess.run([str(APP), "-case", str(case), "-time", "0", "rho", "rhoU"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    vtk = (case / "VTK" / "vcase_0.vtk").read_text().split("\n")
    K, Np = 32, 10
    assert vtk[4] == f"POINTS {K * Np} double"
    cells = vtk.index(f"CELLS {K * N * N} {K * N * N * 4}")
    tri = np.array([l.split() for l in vtk[cells + 1:cells + 1 + K * N * N]], dtype=int)
    assert (tri[:, 0] == 3).all() and tri[:, 1:].max() == K * Np - 1
    pts = np.array([l.split() for l in vtk[5:5 + K * Np]], dtype=float)
    # every sub-triangle is counter-clockwise and the sub-triangles of an element tile it exactly
    a, b, c = pts[tri[:, 1]], pts[tri[:, 2]], pts[tri[:, 3]]
    area = 0.5 * ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]))
    assert (area > 0).all()
    assert abs(area.sum() - 100.0) < 1e-9                      # the [0,10]x[-5,5] square
    assert "SCALARS rho double 1" in vtk and "VECTORS rhoU double" in vtk
    i = vtk.index("VECTORS rhoU double")
    assert vtk[i + 1].split() == ["1", "0", "0"]
