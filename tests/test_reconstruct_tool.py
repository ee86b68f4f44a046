"""hopeDgReconstructPar (host-only, no GPU): fields written per processor are merged back through cellProcAddressing."""
import re
import subprocess
from pathlib import Path

import numpy as np

from hopefoam_b200 import meshgen
from tests.case_writer import HDR, read_field, write_euler_case
from tests.polymesh_writer import write_processor_polymeshes

ROOT = Path(__file__).resolve().parent.parent
TOOL = ROOT / "hopefoam_b200" / "apps" / "bin" / "hopeDgReconstructPar"


def _write(path, name, cls, vals, procs, bvals=None):
    n = vals.shape[0]
    if vals.ndim == 1:
        body = "\n".join(repr(float(v)) for v in vals)
        typ = "scalar"
    else:
        body = "\n".join("(" + " ".join(repr(float(c)) for c in v) + ")" for v in vals)
        typ = "vector"
    t = HDR.format(cls=cls, obj=name) + f"\ndimensions      [1 -3 0 0 0 0 0];\n\ninternalField   nonuniform List<{typ}> \n{n}\n(\n{body}\n)\n;\n\nboundaryField\n{{\n"
    if bvals is None:
        bv = "uniform " + ("0" if vals.ndim == 1 else "(0 0 0)")
    elif bvals.ndim == 1:
        bv = f"nonuniform List<scalar> {bvals.shape[0]}(" + " ".join(repr(float(v)) for v in bvals) + ")"
    else:
        bv = f"nonuniform List<vector> {bvals.shape[0]}(" + " ".join("(" + " ".join(repr(float(c)) for c in v) + ")" for v in bvals) + ")"
    t += "    boundary\n    {\n        type            fixedValue;\n        value           " + bv + ";\n    }\n"
    t += "    frontAndBackPlanes\n    {\n        type            empty;\n    }\n"
    for pn in procs:
        t += f"    {pn}\n    {{\n        type            processor;\n    }}\n"
    path.parent.mkdir(parents=True, exist_ok=True)
    path.write_text(t + "}\n")


def test_reconstruct_fields_from_three_processors(tmp_path, built_library):
    if not TOOL.exists():
        subprocess.run(["make", "-C", str(ROOT / "hopefoam_b200" / "csrc"), "apps"], check=True)
    N, nprocs = 3, 3
    Np = (N + 1) * (N + 2) // 2
    mg = meshgen.jittered_square(6)
    patches = [("boundary", "patch", mg["patch_edges"][0])]
    case = write_euler_case(tmp_path / "case", mg, N, 1e-3, 1e-2)
    K = mg["tris"].shape[0]
    rng = np.random.default_rng(7)
    c2p = rng.integers(0, nprocs, K)                      # an arbitrary (non-contiguous) decomposition
    write_processor_polymeshes(case, mg["xy"], mg["tris"], patches, c2p, nprocs)
    rho = rng.standard_normal((K, Np))
    rhoU = rng.standard_normal((K, Np, 3))
    from tests.helpers import HostContext
    fb = lambda xy: np.sin(3 * xy[:, 0]) + 2 * xy[:, 1]                                   # boundary data as a function of the node position
    fbU = lambda xy: np.stack([xy[:, 0] * xy[:, 1], np.cos(xy[:, 0]), 0 * xy[:, 0]], -1)
    for r in range(nprocs):
        pdir = case / f"processor{r}"
        procs = re.findall(r"(procBoundary\d+to\d+)", (pdir / "constant" / "polyMesh" / "boundary").read_text())
        cells = np.nonzero(c2p == r)[0]
        hc = HostContext()
        hc.set_order(N)
        hc.set_mesh_polymesh(pdir / "constant" / "polyMesh")
        pxy = hc.patch_node_coords(0).reshape(-1, 2)                                     # patch `boundary`, this processor's faces, patch-dof order
        _write(pdir / "0.01" / "rho", "rho", "dgScalarField", rho[cells].reshape(-1), procs, fb(pxy))
        _write(pdir / "0.01" / "rhoU", "rhoU", "dgVectorField", rhoU[cells].reshape(-1, 3), procs, fbU(pxy))
        _write(pdir / "0.005" / "rho", "rho", "dgScalarField", np.zeros(cells.size * Np), procs)      # an older time: must not be picked
    out = subprocess.run([str(TOOL), "-case", str(case)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "time 0.01 from 3 processors" in out.stdout
    # every value lands in its cell and node; written with the case's writePrecision (16 digits: the last bit may differ)
    assert np.abs(read_field(case / "0.01" / "rho", 1).reshape(K, Np) - rho).max() <= 1e-15
    assert np.abs(read_field(case / "0.01" / "rhoU", 3).reshape(K, Np, 3) - rhoU).max() <= 1e-15
    txt = (case / "0.01" / "rhoU").read_text()
    assert "class       dgVectorField;" in txt and "type            fixedValue;" in txt and "procBoundary" not in txt
    # the fixedValue data of the original patch come back in the undecomposed patch-dof order (a restart keeps its boundary data)
    gc = HostContext()
    gc.set_order(N)
    gc.set_mesh_polymesh(case / "constant" / "polyMesh")
    gxy = gc.patch_node_coords(0).reshape(-1, 2)

    def patch_values(path, nc):
        m = re.search(r"boundary\s*\{[^}]*?value\s+nonuniform List<\w+>\s*(\d+)\s*\((.*?)\);\s*\}", path.read_text(), re.S)
        assert m, "no nonuniform value list on patch `boundary`"
        v = np.array([float(x) for x in re.findall(r"[-+0-9.eE]+", m.group(2))])
        assert int(m.group(1)) * nc == v.size
        return v.reshape(-1, nc)

    assert np.abs(patch_values(case / "0.01" / "rho", 1)[:, 0] - fb(gxy)).max() <= 1e-15
    assert np.abs(patch_values(case / "0.01" / "rhoU", 3) - fbU(gxy)).max() <= 1e-15
    # no processor directories: the reference-style fatal error
    bad = subprocess.run([str(TOOL), "-case", str(tmp_path)], capture_output=True, text=True)
    assert bad.returncode != 0
