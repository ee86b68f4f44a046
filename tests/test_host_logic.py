"""CPU tests of the product's host logic (host-only context, no GPU): reference-element operators and DG connectivity
against the oracle; polyMesh reader; analytic identities the reference relies on."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from hopefoam_b200 import meshgen, partition
from oracle import dg_oracle as o
from tests import helpers as H
from tests.polymesh_writer import write_polymesh, write_processor_polymeshes

GOLD = Path(__file__).resolve().parent / "golden"
REF_CYL = Path("/root/reference/HopeFOAM-0.1/tutorials/DG/2D/cylinder/constant/polyMesh")


@pytest.mark.parametrize("N", range(1, 11))      # 9, 10: own collapsed cubature (the reference stops at N = 8)
def test_operators_match_oracle(built_library, N):
    c = H.HostContext()
    c.set_order(N)
    ref = o.RefElement(N)
    assert (c.Np, c.Nfp, c.Ng, c.Nfg) == (ref.Np, ref.Nfp, ref.Ng, ref.Nfg)
    Mref = ref.Vg.T @ np.diag(ref.gw) @ ref.Vg
    Mi = np.linalg.inv(Mref)
    want = dict(r=ref.r, s=ref.s, V=ref.V, invV=ref.invV, Dr=ref.Dr, Ds=ref.Ds, gr=ref.gr, gs=ref.gs, gw=ref.gw, Vg=ref.Vg,
                Dgr=ref.Dgr, Dgs=ref.Dgs, fx=ref.fx, fw=ref.fw, If=ref.If, Mref=Mref,
                Pr=Mi @ ref.Dgr.T @ np.diag(ref.gw), Ps=Mi @ ref.Dgs.T @ np.diag(ref.gw))
    for name, val in want.items():
        got = c.operator(name)
        assert got.size == val.size
        assert np.abs(got - val.reshape(-1)).max() <= 1e-12 * max(1.0, np.abs(val).max()), name
    assert (c.face_to_cell_index() == ref.f2c).all()          # integer maps: bit-exact


@pytest.mark.parametrize("N", [1, 4, 8, 9, 10])
def test_operator_identities(built_library, N):
    c = H.HostContext()
    c.set_order(N)
    Np, Ng, Nfg, Nfp = c.Np, c.Ng, c.Nfg, c.Nfp
    assert abs(c.operator("gw").sum() - 2.0) < 1e-12          # cubature weights sum to the triangle area (…DataTable.C:37-38)
    assert abs(c.operator("fw").sum() - 2.0) < 1e-13
    assert np.abs(c.operator("Dr", (Np, Np)).sum(1)).max() < 1e-11     # derivative of a constant
    assert np.abs(c.operator("Vg", (Ng, Np)).sum(1) - 1).max() < 1e-12  # interpolation reproduces constants
    V = c.operator("V", (Np, Np))
    Mref = c.operator("Mref", (Np, Np))
    assert np.abs(Mref - np.linalg.inv(V @ V.T)).max() < 1e-11          # Mref = (V V^T)^-1 (baseFunction.C:63-77)
    # cubature exactness: integrates x^a y^b exactly for a+b <= 3(N+1)
    gr, gs, gw = c.operator("gr"), c.operator("gs"), c.operator("gw")
    from math import factorial
    deg = 3 * (N + 1)
    for a, b in [(deg, 0), (deg // 2, deg - deg // 2), (1, deg - 1)]:
        # integral over the reference triangle of ((1+r)/2)^a ((1+s)/2)^b dr ds = 4 a! b! / (a+b+2)!
        want = 4.0 * factorial(a) * factorial(b) / factorial(a + b + 2)
        got = (gw * ((1 + gr) / 2) ** a * ((1 + gs) / 2) ** b).sum()
        assert abs(got - want) < 5e-12 * max(1.0, want) + 1e-14
    # weak-form consistency: Pr*Vg*1-vector relation  sum_j Mref (Pr Vg)[:, j] = Dgr^T w Vg -> check Dw = Pr Vg
    Pr, Vg = c.operator("Pr", (Np, Ng)), c.operator("Vg", (Ng, Np))
    assert np.abs(c.operator("Dwr", (Np, Np)) - Pr @ Vg).max() < 1e-12


@pytest.mark.parametrize("periodic", [False, True])
def test_connectivity_matches_oracle_bit_exact(built_library, periodic):
    mg = meshgen.jittered_square(9, periodic=periodic)
    om = H.oracle_mesh(mg)
    c = H.HostContext()
    c.set_order(3)
    c.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
    f = c.faces()
    assert (c.K, c.F) == (om.K, om.F)
    for k, a in (("owner", "face_owner"), ("nbr", "face_nbr"), ("loc_o", "face_loc_o"), ("loc_n", "face_loc_n"), ("rot", "face_rot")):
        assert (f[k] == getattr(om, a)).all(), k
    assert (c.cell_vertices() == om.tris).all()
    for ip, p in enumerate(om.patches):
        assert (c.patch_faces(ip) == p["faces"]).all()
    if periodic:
        assert (f["nbr"] >= 0).all() and set(f["rot"].tolist()) == {1}      # conforming CCW mesh: every face rotated (App. A.12)
    case = o.Case(om, 3)
    assert np.abs(c.node_coords() - case.geo.x).max() < 1e-13
    for ip in range(len(om.patches)):
        assert np.abs(c.patch_node_coords(ip) - case.patch_internal(case.geo.x, ip)).max() < 1e-13


def test_polymesh_reader_matches_oracle(built_library, tmp_path):
    mg = meshgen.jittered_square(5)
    e = mg["patch_edges"][0]
    patches = [("inlet", "patch", e[:7]), ("walls", "wall", e[7:])]
    write_polymesh(tmp_path, mg["xy"], mg["tris"], patches)
    om = o.mesh_from_polymesh(tmp_path)
    c = H.HostContext()
    c.set_order(2)
    c.set_mesh_polymesh(tmp_path)
    assert (c.K, c.F, c.n_patches) == (om.K, om.F, 3)
    f = c.faces()
    for k, a in (("owner", "face_owner"), ("nbr", "face_nbr"), ("loc_o", "face_loc_o"), ("loc_n", "face_loc_n"), ("rot", "face_rot")):
        assert (f[k] == getattr(om, a)).all(), k
    assert (c.cell_vertices() == om.tris).all()
    # v0 is the first point of the z==0 face AS STORED (dgPolyMesh.C:154-190): the writer rotated it per cell
    assert not (c.cell_vertices()[:, 0] == mg["tris"][:, 0]).all()
    assert [c.patch_info(i)[:2] for i in range(3)] == [("inlet", "patch"), ("walls", "wall"), ("frontAndBackPlanes", "empty")]
    assert c.patch_info(2)[2] == 0 and c.patch_info(0)[2] == 7
    for ip in range(2):
        assert (c.patch_faces(ip) == om.patches[ip]["faces"]).all()


def _h(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int32).tobytes()).hexdigest()


@pytest.mark.skipif(not REF_CYL.exists(), reason="reference polyMesh fixture only exists in the build container")
def test_cylinder_polymesh_connectivity(built_library):
    """The only polyMesh the reference ships: 1840 prisms -> 2666 interior dgFaces + 128+12+16+16+16 patch faces
    (TUT/cylinder/constant/polyMesh/boundary:21-80); product connectivity == committed checksums of the oracle's."""
    gold = json.loads((GOLD / "cylinder_connectivity.json").read_text())
    c = H.HostContext()
    c.set_order(6)
    c.set_mesh_polymesh(REF_CYL)
    f = c.faces()
    assert (c.K, c.F, int((f["nbr"] >= 0).sum())) == (1840, 2854, 2666)
    assert [list(c.patch_info(i)) for i in range(c.n_patches)] == gold["patches"]
    s = gold["sha256"]
    assert _h(c.cell_vertices()) == s["tris"]
    assert (_h(f["owner"]), _h(f["nbr"]), _h(f["loc_o"]), _h(f["loc_n"]), _h(f["rot"])) == \
        (s["face_owner"], s["face_nbr"], s["face_loc_o"], s["face_loc_n"], s["face_rot"])
    assert [_h(c.patch_faces(i)) for i in range(c.n_patches)] == s["patch_faces"]


def test_mesh_error_paths(built_library):
    from hopefoam_b200 import capi
    c = H.HostContext()
    with pytest.raises(capi.HdgError):
        c.set_mesh_triangles(np.zeros((3, 2)), np.array([[0, 1, 2]]))          # order not set
    c.set_order(2)
    with pytest.raises(capi.HdgError):
        c.set_order(11)                                                          # orders 1..10 exist (9, 10 with own cubature); 11 does not
    with pytest.raises(capi.HdgError):
        c.set_mesh_triangles(np.array([[0., 0], [1, 0], [2, 0]]), np.array([[0, 1, 2]]))   # degenerate triangle
    mg = meshgen.jittered_square(3)
    with pytest.raises(capi.HdgError):                                           # boundary faces without a patch
        c.set_mesh_triangles(mg["xy"], mg["tris"], None, [])
    with pytest.raises(capi.HdgError):
        c.set_mesh_polymesh("/nonexistent/polyMesh")


def test_strip_partition_faces_pair_up(built_library):
    """Host-side halo logic: face k of my bottom patch and face k of the lower strip's top patch are the same edge,
    traversed in opposite directions (so the sender-side reversal of processorDgPatchField.C:253-260 aligns the nodes)."""
    n, world = 6, 3
    ctxs, parts = [], []
    for r in range(world):
        mg = partition.strip_partition(n, world, r)
        c = H.HostContext()
        c.set_order(3)
        c.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], mg["patch_edges"])
        ctxs.append(c)
        parts.append(mg)
    for r in range(world):
        down = parts[r]["peers"][0]
        mine = ctxs[r].patch_node_coords(0).reshape(n, -1, 2)          # my bottom faces, my traversal order
        theirs = ctxs[down].patch_node_coords(1).reshape(n, -1, 2)     # their top faces, their traversal order
        shift = np.array([0.0, 10.0 * world if down > r else 0.0])
        assert np.abs(mine - (theirs[:, ::-1, :] - shift)).max() < 1e-12


@pytest.mark.parametrize("n_div", [(2, 1, 1), (2, 2, 1), (3, 2, 1)])
def test_decomposition_maps_match_oracle_bit_exact(built_library, n_div):
    """`method simple` + dgDecomposePar rules: cellToProc, cellProcAddressing, pointProcAddressing, patch order, neighbour processor per
    patch and the cut-face order are INTEGER maps -> bit-exact against the oracle's restatement, for every rank."""
    mg = meshgen.jittered_square(9)
    e = mg["patch_edges"][0]
    mg["patch_edges"] = [e[:9], e[9:]]
    om = H.oracle_mesh(mg)
    g = H.HostContext()
    g.set_order(3)
    g.set_mesh_triangles(mg["xy"], mg["tris"], None, mg["patch_edges"])
    nprocs = n_div[0] * n_div[1] * n_div[2]
    c2p = g.decompose_simple(*n_div, 0.001)
    want = o.simple_decomp(om, n_div, 0.001)
    assert (c2p == want).all()
    counts = np.bincount(c2p, minlength=nprocs)
    assert counts.max() - counts.min() <= nprocs          # banded assignment: near-perfect balance
    seen_faces = {}
    for r in range(nprocs):
        loc = H.HostContext()
        loc.set_order(3)
        loc.set_mesh_from_decomposition(g, c2p, nprocs, r)
        addr = loc.proc_addressing()
        od = o.decompose(om, want, nprocs, r)
        assert (addr["cell"] == od["cell"]).all() and (addr["point"] == od["point"]).all()
        assert (loc.cell_vertices() == od["tris"]).all()
        assert loc.n_patches == len(od["patches"])
        assert addr["patch_nbr_proc"].tolist() == [q for _, q, _ in od["patches"]]
        assert addr["patch_face_global"].tolist() == [f for _, _, fs in od["patches"] for f in fs]
        assert [loc.patch_info(p)[0] for p in range(loc.n_patches)] == [nm for nm, _, _ in od["patches"]]
        # node coordinates of the processor mesh are those of the global cells it owns
        assert np.abs(loc.node_coords() - g.node_coords()[addr["cell"]]).max() == 0.0
        for p, (_, q, fs) in enumerate(od["patches"]):
            if q >= 0:
                seen_faces[(r, q)] = (fs, loc.patch_node_coords(p).reshape(len(fs), -1, 2))
    for (r, q), (fs, xy) in seen_faces.items():          # both sides list the cut faces in the same order, traversed oppositely
        fs2, xy2 = seen_faces[(q, r)]
        assert fs == fs2
        assert np.abs(xy - xy2[:, ::-1, :]).max() < 1e-13        # same edge seen from the two cells (round-off only)


@pytest.mark.parametrize("n_div", [(1, 2, 1), (2, 2, 1), (3, 1, 1)])
def test_decomposition_of_a_periodic_mesh(built_library, n_div):
    """Periodic gluing (pointEquiv; an extension - the reference has no compiled cyclic patch) survives decomposition: glued faces inside
    a processor stay interior faces, glued faces between processors become processor faces, every global face is accounted for once,
    both sides list a cut in the same order, and the par plan splits the octets into those with / without a processor face."""
    mg = meshgen.jittered_square(8, periodic=True)
    g = H.HostContext()
    g.set_order(2)
    g.set_mesh_triangles(mg["xy"], mg["tris"], mg["point_equiv"], [])
    gf = g.faces()
    assert (gf["nbr"] >= 0).all()
    nprocs = int(np.prod(n_div))
    c2p = g.decompose_simple(*n_div, 0.001)
    interior, cut = 0, {}
    for r in range(nprocs):
        loc = H.HostContext()
        loc.set_order(2)
        loc.set_mesh_from_decomposition(g, c2p, nprocs, r)
        addr = loc.proc_addressing()
        assert (c2p[addr["cell"]] == r).all() and (np.diff(addr["cell"]) > 0).all()
        lf = loc.faces()
        interior += int((lf["nbr"] >= 0).sum())
        # an interior face of the processor mesh joins the same two global cells as in the global mesh
        for f in np.nonzero(lf["nbr"] >= 0)[0]:
            a, b = int(addr["cell"][lf["owner"][f]]), int(addr["cell"][lf["nbr"][f]])
            hit = ((gf["owner"] == min(a, b)) & (gf["nbr"] == max(a, b))).sum()
            assert hit == 1
        off = 0
        for p in range(loc.n_patches):
            nf = loc.patch_info(p)[2]
            q = int(addr["patch_nbr_proc"][p])
            assert q >= 0 and q != r and loc.patch_info(p)[1] == "processor"
            cut[(r, q)] = addr["patch_face_global"][off:off + nf].tolist()
            off += nf
    for (r, q), fs in cut.items():
        assert cut[(q, r)] == fs and len(set(fs)) == len(fs)
        assert all({int(c2p[gf["owner"][f]]), int(c2p[gf["nbr"][f]])} == {r, q} for f in fs)
    assert interior + sum(len(v) for v in cut.values()) // 2 == g.F


@pytest.mark.parametrize("nprocs", [2, 3, 5, 8])
def test_graph_partitioner_contract(built_library, nprocs, tmp_path):
    """`method scotch | metis` -> the native graph partitioner (hdg_decompose_graph): parts balanced to one cell, every part connected, cut
    no larger than the geometric strips of `simple`, deterministic, and selected by the decomposeParDict keyword like the reference's
    decompositionMethod::New.  (The cellToProc itself cannot equal scotch's: the library is not buildable here.)"""
    mg = meshgen.jittered_square(24)
    g = H.HostContext()
    g.set_order(2)
    g.set_mesh_triangles(mg["xy"], mg["tris"], None, mg["patch_edges"])
    c2p = g.decompose_graph(nprocs)
    sizes = np.bincount(c2p, minlength=nprocs)
    assert sizes.max() - sizes.min() <= 1 and sizes.sum() == g.K
    assert np.array_equal(c2p, g.decompose_graph(nprocs))
    f = g.faces()
    inner = f["nbr"] >= 0
    own, nbr = f["owner"][inner], f["nbr"][inner]
    cut = int((c2p[own] != c2p[nbr]).sum())
    strips = g.decompose_simple(nprocs, 1, 1)
    assert cut <= int((strips[own] != strips[nbr]).sum())
    for r in range(nprocs):                                 # connected parts: flood fill over the faces inside the part
        cells = np.nonzero(c2p == r)[0]
        adj = {int(c): [] for c in cells}
        for a, b in zip(own, nbr):
            if c2p[a] == r and c2p[b] == r:
                adj[int(a)].append(int(b)); adj[int(b)].append(int(a))
        seen, todo = {int(cells[0])}, [int(cells[0])]
        while todo:
            for n in adj[todo.pop()]:
                if n not in seen:
                    seen.add(n); todo.append(n)
        assert len(seen) == cells.size
    (tmp_path / "system").mkdir()
    for method in ("scotch", "metis"):
        (tmp_path / "system" / "decomposeParDict").write_text("FoamFile\n{\n    version 2.0;\n    format ascii;\n    class dictionary;\n    object decomposeParDict;\n}\n"
                                                              f"numberOfSubdomains {nprocs};\nmethod {method};\n")
        n, got = g.decompose_from_dict(tmp_path)
        assert n == nprocs and np.array_equal(got, c2p)


def test_decomposition_uses_polymesh_face_order(built_library, tmp_path):
    """When the mesh comes from a polyMesh directory the cut faces follow the polyMesh face ids (the reference's ascending
    global face id), whatever order the writer chose for the internal faces."""
    mg = meshgen.jittered_square(6)
    write_polymesh(tmp_path, mg["xy"], mg["tris"], [("boundary", "patch", mg["patch_edges"][0])])
    g = H.HostContext()
    g.set_order(2)
    g.set_mesh_polymesh(tmp_path)
    om = o.mesh_from_polymesh(tmp_path)
    pm = o.read_polymesh(tmp_path)
    # polyMesh id of the lateral face behind each dgFace, restated independently
    lat = {}
    for fid, pts in enumerate(pm["faces"]):
        z0 = [int(q) for q in pts if pm["points"][q, 2] == 0.0]
        if len(z0) == 2 and len(pts) == 4:
            lat[(min(z0), max(z0))] = fid
    poly_face = []
    for f in range(om.F):
        c, lf = om.face_owner[f], om.face_loc_o[f]
        a, b = int(om.tris[c, lf]), int(om.tris[c, (lf + 1) % 3])
        poly_face.append(lat[(min(a, b), max(a, b))])
    c2p = g.decompose_simple(2, 2, 1, 0.001)
    for r in range(4):
        loc = H.HostContext()
        loc.set_order(2)
        loc.set_mesh_from_decomposition(g, c2p, 4, r)
        od = o.decompose(om, c2p, 4, r, poly_face=np.array(poly_face))
        assert loc.proc_addressing()["patch_face_global"].tolist() == [f for _, _, fs in od["patches"] for f in fs]


def test_two_triangle_hand_mesh(built_library):
    """Hand-checkable connectivity: unit square split along the diagonal 0-2.
       cell 0 = (0,1,2), cell 1 = (0,2,3); dgFaces in cell-major/local-face-minor order created by the lower cell."""
    xy = np.array([[0.0, 0], [1, 0], [1, 1], [0, 1]])
    tris = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int32)
    edges = np.array([[0, 0, 1], [0, 1, 2], [1, 2, 3], [1, 3, 0]], dtype=np.int32)
    c = H.HostContext()
    c.set_order(1)
    c.set_mesh_triangles(xy, tris, None, [edges])
    f = c.faces()
    assert f["owner"].tolist() == [0, 0, 0, 1, 1]
    assert f["nbr"].tolist() == [-1, -1, 1, -1, -1]
    assert f["loc_o"].tolist() == [0, 1, 2, 1, 2]
    assert f["loc_n"].tolist() == [-1, -1, 0, -1, -1]          # the diagonal is local face 0 (v0->v1 = 0->2) of cell 1
    assert f["rot"].tolist() == [-1, -1, 1, -1, -1]            # cell 0 walks it 2->0, cell 1 walks it 0->2: rotated
    assert c.patch_faces(0).tolist() == [0, 1, 3, 4]
    # a clockwise input triangle is turned counter-clockwise by swapping v1, v2 (dgPolyMesh.C:490-509)
    c.set_mesh_triangles(xy, np.array([[0, 2, 1], [0, 2, 3]], dtype=np.int32), None, [edges])
    assert c.cell_vertices().tolist() == [[0, 1, 2], [0, 2, 3]]


def test_processor_directories_drop_in(built_library, tmp_path):
    """A case decomposed into processorN/constant/polyMesh directories (dgDecomposePar layout) loads rank by rank and gives the same
    processor meshes - cells, vertices, patches, neighbour ranks, cut-face order - as decomposing the global mesh in memory."""
    mg = meshgen.jittered_square(7)
    e = mg["patch_edges"][0]
    patches = [("inlet", "patch", e[:7]), ("walls", "wall", e[7:])]
    write_polymesh(tmp_path / "constant" / "polyMesh", mg["xy"], mg["tris"], patches, rotate_vertices=False)
    g = H.HostContext()
    g.set_order(2)
    g.set_mesh_polymesh(tmp_path / "constant" / "polyMesh")
    nprocs = 4
    c2p = g.decompose_simple(2, 2, 1, 0.001)
    # the writer rebuilds the global polyMesh itself (no vertex rotation) - cells/points keep their ids
    import tests.polymesh_writer as pw
    orig = pw.write_polymesh
    pw.write_polymesh = lambda d, xy, tris, pats, thickness=1.0: orig(d, xy, tris, pats, thickness, rotate_vertices=False)
    try:
        write_processor_polymeshes(tmp_path, mg["xy"], mg["tris"], patches, c2p, nprocs)
    finally:
        pw.write_polymesh = orig
    for r in range(nprocs):
        mem = H.HostContext(); mem.set_order(2)
        mem.set_mesh_from_decomposition(g, c2p, nprocs, r)
        disk = H.HostContext(); disk.set_order(2)
        disk.set_mesh_polymesh(tmp_path / f"processor{r}" / "constant" / "polyMesh")
        assert (disk.K, disk.F) == (mem.K, mem.F)
        fm, fd = mem.faces(), disk.faces()
        for k in fm:
            assert (fm[k] == fd[k]).all(), k
        assert (mem.cell_vertices() == disk.cell_vertices()).all()
        assert np.abs(mem.node_coords() - disk.node_coords()).max() == 0.0
        am, ad = mem.proc_addressing(), disk.proc_addressing()
        # on disk the empty front/back patch sits between the original and the processor patches; it owns no dgFaces
        keep = [p for p in range(disk.n_patches) if disk.patch_info(p)[1] != "empty"]
        keep_m = [p for p in range(mem.n_patches) if mem.patch_info(p)[1] != "empty"]
        assert [disk.patch_info(p)[0] for p in keep] == [mem.patch_info(p)[0] for p in keep_m]
        assert ad["patch_nbr_proc"][keep].tolist() == am["patch_nbr_proc"][keep_m].tolist()
        assert max(ad["patch_nbr_proc"][keep]) >= 0
        for pm_, pd_ in zip(keep_m, keep):
            assert (mem.patch_faces(pm_) == disk.patch_faces(pd_)).all()


def test_decompose_from_dict_simple_and_manual(built_library, tmp_path):
    """system/decomposeParDict drives the decomposition like dgDecomposePar: `method simple` (the tutorial's file) and
    `method manual` with a cellDecomposition labelList (how a scotch decomposition made elsewhere drops in)."""
    from hopefoam_b200 import capi
    mg = meshgen.jittered_square(6)
    g = H.HostContext()
    g.set_order(2)
    g.set_mesh_triangles(mg["xy"], mg["tris"], None, mg["patch_edges"])
    (tmp_path / "system").mkdir()
    (tmp_path / "constant").mkdir()
    hdr = "FoamFile\n{\n    version 2.0;\n    format ascii;\n    class dictionary;\n    object decompositionDict;\n}\n"
    (tmp_path / "system" / "decomposeParDict").write_text(hdr + "numberOfSubdomains 6;\nmethod simple;\nsimpleCoeffs\n{\n    n        (3 2 1);\n    delta    0.001;\n}\n")
    n, c2p = g.decompose_from_dict(tmp_path)
    assert n == 6 and (c2p == g.decompose_simple(3, 2, 1, 0.001)).all()
    # manual: any labelList, e.g. one produced by scotch on another machine
    rng = np.random.default_rng(3)
    manual = rng.integers(0, 3, size=g.K)
    (tmp_path / "constant" / "cellDecomposition").write_text(
        "FoamFile\n{\n    version 2.0;\n    format ascii;\n    class labelList;\n    object cellDecomposition;\n}\n// comment\n"
        f"{g.K}\n(\n" + "\n".join(map(str, manual)) + "\n)\n")
    (tmp_path / "system" / "decomposeParDict").write_text(hdr + 'numberOfSubdomains 3;\nmethod manual;\nmanualCoeffs\n{\n    dataFile "cellDecomposition";\n}\n')
    n, c2p = g.decompose_from_dict(tmp_path)
    assert n == 3 and (c2p == manual).all()
    # every rank of the (scattered) manual decomposition still yields a consistent processor mesh
    om = H.oracle_mesh(mg)
    for r in range(3):
        loc = H.HostContext(); loc.set_order(2)
        loc.set_mesh_from_decomposition(g, c2p, 3, r)
        od = o.decompose(om, manual, 3, r)
        a = loc.proc_addressing()
        assert (a["cell"] == od["cell"]).all() and a["patch_face_global"].tolist() == [f for _, _, fs in od["patches"] for f in fs]
    # error conventions
    (tmp_path / "system" / "decomposeParDict").write_text(hdr + "numberOfSubdomains 4;\nmethod simple;\nsimpleCoeffs\n{\n    n (3 2 1);\n    delta 0.001;\n}\n")
    with pytest.raises(capi.HdgError, match="Wrong number of processor divisions"):
        g.decompose_from_dict(tmp_path)
    (tmp_path / "system" / "decomposeParDict").write_text(hdr + "numberOfSubdomains 4;\nmethod hierarchical;\n")
    with pytest.raises(capi.HdgError, match="Unknown decompositionMethod hierarchical"):
        g.decompose_from_dict(tmp_path)
