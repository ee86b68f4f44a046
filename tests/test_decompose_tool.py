"""hopeDgDecomposePar (host-only, no GPU): processorN/constant/polyMesh written by the tool load as the processor meshes of the
dgDecomposePar rules, fields are split through the cell / patch-face addressing, and hopeDgReconstructPar inverts it."""
import re
import subprocess
from pathlib import Path

import numpy as np

from hopefoam_b200 import meshgen
from tests import helpers as H
from tests.case_writer import HDR, read_field, write_euler_case

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "hopefoam_b200" / "apps" / "bin"


def _labels(path):
    t = Path(path).read_text()
    return np.array(t[t.index("(", t.index("// *")) + 1: t.rindex(")")].split(), dtype=int)


def _nonuniform(name, cls, vals, bvals):
    typ = "scalar" if vals.ndim == 1 else "vector"
    fmt = (lambda v: repr(float(v))) if vals.ndim == 1 else (lambda v: "(" + " ".join(repr(float(c)) for c in v) + ")")
    t = HDR.format(cls=cls, obj=name) + f"\ndimensions      [1 -3 0 0 0 0 0];\n\ninternalField   nonuniform List<{typ}> \n{vals.shape[0]}\n(\n"
    t += "\n".join(fmt(v) for v in vals) + "\n)\n;\n\nboundaryField\n{\n    boundary\n    {\n        type            fixedValue;\n"
    t += f"        value           nonuniform List<{typ}> {bvals.shape[0]}(" + " ".join(fmt(v) for v in bvals) + ");\n    }\n"
    return t + "    frontAndBackPlanes\n    {\n        type            empty;\n    }\n}\n"


def test_decompose_then_reconstruct(tmp_path, built_library):
    if not (BIN / "hopeDgDecomposePar").exists() or not (BIN / "hopeDgReconstructPar").exists():
        subprocess.run(["make", "-C", str(ROOT / "hopefoam_b200" / "csrc"), "apps"], check=True)
    N, nprocs = 3, 3
    Np, Nfp = (N + 1) * (N + 2) // 2, N + 1
    mg = meshgen.jittered_square(6)
    case = write_euler_case(tmp_path / "case", mg, N, 1e-3, 1e-2)
    K = mg["tris"].shape[0]
    nb = mg["patch_edges"][0].shape[0]
    (case / "system" / "decomposeParDict").write_text(HDR.format(cls="dictionary", obj="decomposeParDict") +
                                                      "\nnumberOfSubdomains 3;\nmethod simple;\nsimpleCoeffs\n{\n    n (3 1 1);\n    delta 0.001;\n}\n")
    rng = np.random.default_rng(11)
    rho, rhoU = rng.standard_normal((K, Np)), rng.standard_normal((K, Np, 3))
    brho, brhoU = rng.standard_normal((nb, Nfp)), rng.standard_normal((nb, Nfp, 3))
    (case / "0" / "rho").write_text(_nonuniform("rho", "dgScalarField", rho.reshape(-1), brho.reshape(-1)))
    (case / "0" / "rhoU").write_text(_nonuniform("rhoU", "dgVectorField", rhoU.reshape(-1, 3), brhoU.reshape(-1, 3)))
    out = subprocess.run([str(BIN / "hopeDgDecomposePar"), "-case", str(case)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Number of processors: 3" in out.stdout

    # the global mesh and the processor meshes through the library's polyMesh reader (host-only contexts)
    g = H.HostContext()
    g.set_order(N)
    g.set_mesh_polymesh(str(case / "constant" / "polyMesh"))
    gx = g.node_coords()
    gfaces = g.patch_faces(0)
    seen = np.zeros(K, dtype=int)
    shared = {}
    for r in range(nprocs):
        pdir = case / f"processor{r}"
        addr = _labels(pdir / "constant" / "polyMesh" / "cellProcAddressing")
        assert np.all(np.diff(addr) > 0)                                  # cells in ascending global id (domainDecompositionMesh.C:124)
        seen[addr] += 1
        c = H.HostContext()
        c.set_order(N)
        c.set_mesh_polymesh(str(pdir / "constant" / "polyMesh"))
        assert c.K == addr.size
        assert np.array_equal(c.node_coords(), gx[addr])                   # same cells, same vertex order: bit-identical node coordinates
        btxt = (pdir / "constant" / "polyMesh" / "boundary").read_text()
        for q, n in re.findall(r"procBoundary%dto(\d+)\s*\{\s*type\s+processor;\s*nFaces\s+(\d+);" % r, btxt):
            shared[(r, int(q))] = int(n)
        assert f"myProcNo        {r};" in btxt or not re.search("procBoundary", btxt)
        # fields: internal values through cellProcAddressing, fixedValue values through the patch-face addressing
        assert np.array_equal(read_field(pdir / "0" / "rho", 1).reshape(-1, Np), rho[addr])
        assert np.array_equal(read_field(pdir / "0" / "rhoU", 3).reshape(-1, Np, 3), rhoU[addr])
        lf = c.patch_faces(0)
        if lf.size:
            fl, fg = c.faces(), g.faces()
            own_l, own_g = fl["owner"][lf], fg["owner"][gfaces]             # owner cells of the local / global boundary faces
            loc_l, loc_g = fl["loc_o"][lf], fg["loc_o"][gfaces]
            pos = [int(np.nonzero((own_g == addr[o]) & (loc_g == l))[0][0]) for o, l in zip(own_l, loc_l)]
            t = (pdir / "0" / "rho").read_text()
            vals = np.array(t[t.index("(", t.index("value")) + 1: t.index(")", t.index("value"))].split(), dtype=float).reshape(-1, Nfp)
            assert np.array_equal(vals, brho[pos])
        c.close()
    assert np.all(seen == 1)
    for (a, b), n in shared.items():
        assert shared[(b, a)] == n                                          # both sides of a processor patch list the same number of faces
    g.close()

    # round trip: the decomposed time-0 fields reconstructed over the originals
    (case / "0" / "rho").unlink()
    (case / "0" / "rhoU").unlink()
    rec = subprocess.run([str(BIN / "hopeDgReconstructPar"), "-case", str(case), "-time", "0", "rho", "rhoU"], capture_output=True, text=True, timeout=120)
    assert rec.returncode == 0, rec.stdout + rec.stderr
    assert np.abs(read_field(case / "0" / "rho", 1).reshape(K, Np) - rho).max() <= 1e-15
    assert np.abs(read_field(case / "0" / "rhoU", 3).reshape(K, Np, 3) - rhoU).max() <= 1e-15
