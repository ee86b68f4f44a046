"""hopeDgDecomposePar (host-only, no GPU): processorN/constant/polyMesh written by the tool load as the processor meshes of the
dgDecomposePar rules, fields are split through the cell / patch-face addressing, and hopeDgReconstructPar inverts it."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

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
    # the processor polyMesh files are the reference's cut of the global polyMesh, bit for bit: faceProcAddressing with the turning index
    # (domainDecomposition.C:1014-1027), points / faces / owner / neighbour / patch ranges against the oracle's restatement
    from oracle import dg_oracle as o
    pm = o.read_polymesh(case / "constant" / "polyMesh")
    c2p = np.empty(K, dtype=int)
    for r in range(nprocs):
        c2p[_labels(case / f"processor{r}" / "constant" / "polyMesh" / "cellProcAddressing")] = r
    for r in range(nprocs):
        pd = case / f"processor{r}" / "constant" / "polyMesh"
        want = o.decompose_polymesh(pm, c2p, r)
        assert np.array_equal(_labels(pd / "faceProcAddressing"), want["face"])
        assert np.array_equal(_labels(pd / "pointProcAddressing"), want["point"])
        assert np.array_equal(_labels(pd / "cellProcAddressing"), want["cell"])
        lp = o.read_polymesh(pd)
        assert np.array_equal(lp["owner"], want["owner"]) and np.array_equal(lp["neighbour"], want["neighbour"])
        assert [list(map(int, f)) for f in lp["faces"]] == want["faces"]
        assert np.array_equal(lp["points"], pm["points"][want["point"]])
        assert [(p["name"], p["nFaces"], p["startFace"]) for p in lp["patches"]] == want["patches"]
        assert (want["face"] < 0).any() or r == 0                           # some cut faces are held from the neighbour side (reversed)
        nb = len(pm["patches"])
        assert _labels(pd / "boundaryProcAddressing").tolist() == list(range(nb)) + [-1] * (len(want["patches"]) - nb)
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


REF_CYL = Path("/root/reference/HopeFOAM-0.1/tutorials/DG/2D/cylinder/constant/polyMesh")


@pytest.mark.skipif(not REF_CYL.exists(), reason="reference tutorial polyMesh not present (GPU box)")
def test_reference_cylinder_polymesh_is_cut_as_the_reference_rules_say(tmp_path, built_library):
    """The only polyMesh the reference ships (TUT/cylinder: wall / patch / arc patches with #{ code #} entries, 1840 prisms): the tool's
    processor meshes equal the oracle's restatement of dgDecomposePar bit for bit, the arc entries travel verbatim, and every processor
    mesh loads through the library's reader with the cells / cut faces of the in-memory decomposition."""
    import shutil
    from oracle import dg_oracle as o
    mg = meshgen.jittered_square(2)
    case = write_euler_case(tmp_path / "case", mg, 2, 1e-3, 1e-2)
    shutil.rmtree(case / "constant" / "polyMesh")
    shutil.copytree(REF_CYL, case / "constant" / "polyMesh")
    for f in (case / "0").iterdir():
        f.unlink()
    (case / "system" / "decomposeParDict").write_text(HDR.format(cls="dictionary", obj="decomposeParDict") +
                                                      "\nnumberOfSubdomains 4;\nmethod simple;\nsimpleCoeffs\n{\n    n (2 2 1);\n    delta 0.001;\n}\n")
    out = subprocess.run([str(BIN / "hopeDgDecomposePar"), "-case", str(case)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    pm = o.read_polymesh(REF_CYL)
    g = H.HostContext()
    g.set_order(2)
    g.set_mesh_polymesh(str(REF_CYL))
    c2p = g.decompose_simple(2, 2, 1, 0.001)
    for r in range(4):
        pd = case / f"processor{r}" / "constant" / "polyMesh"
        want = o.decompose_polymesh(pm, c2p, r)
        assert np.array_equal(_labels(pd / "faceProcAddressing"), want["face"])
        assert np.array_equal(_labels(pd / "pointProcAddressing"), want["point"])
        lp = o.read_polymesh(pd)
        assert [list(map(int, f)) for f in lp["faces"]] == want["faces"]
        assert np.array_equal(lp["owner"], want["owner"]) and np.array_equal(lp["neighbour"], want["neighbour"])
        btxt = (pd / "boundary").read_text()
        assert "0.05*Foam::sin(Foam::constant::mathematical::pi*u)" in btxt and btxt.count("type            arc;") == 2
        loc, mem = H.HostContext(), H.HostContext()
        loc.set_order(2); mem.set_order(2)
        loc.set_mesh_polymesh(str(pd))
        mem.set_mesh_from_decomposition(g, c2p, 4, r)
        assert loc.K == mem.K and np.array_equal(loc.node_coords(), mem.node_coords())
        lf, mf = loc.faces(), mem.faces()
        for p in range(mem.n_patches):                       # same dgFaces (owner cell, local face) in the same order, patch by patch
            assert loc.patch_info(p)[0] == mem.patch_info(p)[0]
            a, b = loc.patch_faces(p), mem.patch_faces(p)
            assert np.array_equal(lf["owner"][a], mf["owner"][b]) and np.array_equal(lf["loc_o"][a], mf["loc_o"][b])
