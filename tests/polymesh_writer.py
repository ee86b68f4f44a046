"""Test utility: write a one-layer prism polyMesh (ASCII, OpenFOAM format) from a 2-D triangle mesh, so that the polyMesh
readers of the product (hdg_set_mesh_polymesh) and of the oracle can be fed the same case directory."""
from pathlib import Path

import numpy as np

HEADER = """/*--------------------------------*- C++ -*----------------------------------*\\
| test polyMesh                                                               |
\\*---------------------------------------------------------------------------*/
FoamFile
{{
    version     2.0;
    format      ascii;
    class       {cls};
    location    "constant/polyMesh";
    object      {obj};
}}
// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //

"""


def write_polymesh(dirpath, xy, tris, patches, thickness=1.0, rotate_vertices=True):
    """patches: list of (name, type, edges (m,3) int = (cell, pa, pb)).  Returns the poly face list for inspection."""
    d = Path(dirpath)
    d.mkdir(parents=True, exist_ok=True)
    P = xy.shape[0]
    K = tris.shape[0]
    pts = np.concatenate([np.c_[xy, np.zeros(P)], np.c_[xy, np.full(P, thickness)]])
    edge_cells = {}
    for c in range(K):
        for f in range(3):
            a, b = int(tris[c, f]), int(tris[c, (f + 1) % 3])
            edge_cells.setdefault((min(a, b), max(a, b)), []).append((c, a, b))
    faces, owner, neigh = [], [], []
    internal = []
    for key, lst in edge_cells.items():
        if len(lst) == 2:
            (c0, a0, b0), (c1, a1, b1) = sorted(lst)
            internal.append((c0, c1, a0, b0))
    internal.sort()
    for c0, c1, a, b in internal:                      # owner = lower cell; normal points out of the owner
        faces.append([a, b, b + P, a + P])
        owner.append(c0)
        neigh.append(c1)
    blocks = []
    for name, typ, edges in patches:
        start = len(faces)
        for c, pa, pb in np.asarray(edges).reshape(-1, 3):
            # orient along the owner's CCW traversal so that the normal is outward
            t = list(tris[c])
            i = t.index(pa)
            a, b = (pa, pb) if t[(i + 1) % 3] == pb else (pb, pa)
            faces.append([int(a), int(b), int(b) + P, int(a) + P])
            owner.append(int(c))
        blocks.append((name, typ, len(faces) - start, start))
    start = len(faces)
    for c in range(K):                                  # base plane z == 0 (outward normal -z) then top plane
        a, b, cc = [int(v) for v in tris[c]]
        if rotate_vertices:                            # vary which vertex comes first, as a mesher would
            r = c % 3
            a, b, cc = [a, b, cc][r:] + [a, b, cc][:r]
        faces.append([a, cc, b])
        owner.append(c)
    for c in range(K):
        a, b, cc = [int(v) + P for v in tris[c]]
        faces.append([a, b, cc])
        owner.append(c)
    blocks.append(("frontAndBackPlanes", "empty", len(faces) - start, start))

    def w(name, cls, body):
        (d / name).write_text(HEADER.format(cls=cls, obj=name) + body)

    w("points", "vectorField", f"{len(pts)}\n(\n" + "\n".join(f"({float(p[0])!r} {float(p[1])!r} {float(p[2])!r})" for p in pts) + "\n)\n")
    w("faces", "faceList", f"{len(faces)}\n(\n" + "\n".join(f"{len(f)}({' '.join(map(str, f))})" for f in faces) + "\n)\n")
    w("owner", "labelList", f"{len(owner)}\n(\n" + "\n".join(map(str, owner)) + "\n)\n")
    w("neighbour", "labelList", f"{len(neigh)}\n(\n" + "\n".join(map(str, neigh)) + "\n)\n")
    body = f"{len(blocks)}\n(\n"
    for name, typ, n, s in blocks:
        body += f"    {name}\n    {{\n        type            {typ};\n        nFaces          {n};\n        startFace       {s};\n    }}\n"
    body += ")\n"
    w("boundary", "polyBoundaryMesh", body)
    return faces


def write_processor_polymeshes(case_dir, xy, tris, patches, cell_to_proc, nprocs, thickness=1.0):
    """Test stand-in for dgDecomposePar: writes <case>/processorN/constant/polyMesh for every rank following
    applications/utilities/DG/dgDecomposePar/domainDecompositionMesh.C:102-511 (cells ascending; internal faces ascending; original
    patches in order; processor patches by ascending neighbour with faces in ascending GLOBAL face id, flipped on the neighbour side;
    points ascending), from the same global polyMesh `write_polymesh` writes (whose face list it rebuilds identically)."""
    import tempfile
    from pathlib import Path as _P
    P = xy.shape[0]
    K = tris.shape[0]
    with tempfile.TemporaryDirectory() as tmp:
        faces = write_polymesh(tmp, xy, tris, patches, thickness)
        t = (_P(tmp) / "owner").read_text()
        owner = np.array(t[t.index("(", t.index("// *")) + 1: t.rindex(")")].split(), dtype=int)
        t = (_P(tmp) / "neighbour").read_text()
        neigh = np.array(t[t.index("(", t.index("// *")) + 1: t.rindex(")")].split(), dtype=int)
    nint = neigh.size
    pts = np.concatenate([np.c_[xy, np.zeros(P)], np.c_[xy, np.full(P, thickness)]])
    # global patch table incl. the empty front/back patch
    starts, blocks = nint, []
    for name, typ, edges in patches:
        n = np.asarray(edges).reshape(-1, 3).shape[0]
        blocks.append((name, typ, starts, n))
        starts += n
    blocks.append(("frontAndBackPlanes", "empty", starts, 2 * K))
    c2p = np.asarray(cell_to_proc)
    for r in range(nprocs):
        cells = np.nonzero(c2p == r)[0]
        g2l_cell = -np.ones(K, dtype=int)
        g2l_cell[cells] = np.arange(cells.size)
        lf, lown, lnei, lblocks = [], [], [], []
        for f in range(nint):
            if c2p[owner[f]] == r and c2p[neigh[f]] == r:
                lf.append((f, False)); lown.append(g2l_cell[owner[f]]); lnei.append(g2l_cell[neigh[f]])
        for name, typ, st, n in blocks:
            s0 = len(lf)
            for f in range(st, st + n):
                if c2p[owner[f]] == r:
                    lf.append((f, False)); lown.append(g2l_cell[owner[f]])
            lblocks.append((name, typ, len(lf) - s0, s0, None))
        cut = {}
        for f in range(nint):
            po, pn = c2p[owner[f]], c2p[neigh[f]]
            if po != pn and r in (po, pn):
                cut.setdefault(int(pn if po == r else po), []).append(f)
        for q in sorted(cut):
            s0 = len(lf)
            for f in cut[q]:
                if c2p[owner[f]] == r:
                    lf.append((f, False)); lown.append(g2l_cell[owner[f]])
                else:
                    lf.append((f, True)); lown.append(g2l_cell[neigh[f]])
            lblocks.append((f"procBoundary{r}to{q}", "processor", len(lf) - s0, s0, q))
        used = np.zeros(2 * P, dtype=bool)
        for f, _ in lf:
            used[faces[f]] = True
        lp = np.nonzero(used)[0]
        g2l_pt = -np.ones(2 * P, dtype=int)
        g2l_pt[lp] = np.arange(lp.size)
        d = _P(case_dir) / f"processor{r}" / "constant" / "polyMesh"
        d.mkdir(parents=True, exist_ok=True)

        def w(name, cls, body):
            (d / name).write_text(HEADER.format(cls=cls, obj=name) + body)
        w("points", "vectorField", f"{lp.size}\n(\n" + "\n".join(f"({float(pts[q, 0])!r} {float(pts[q, 1])!r} {float(pts[q, 2])!r})" for q in lp) + "\n)\n")
        rows = []
        for f, flip in lf:
            ids = [int(g2l_pt[q]) for q in (faces[f][::-1] if flip else faces[f])]
            rows.append(f"{len(ids)}({' '.join(map(str, ids))})")
        w("faces", "faceList", f"{len(rows)}\n(\n" + "\n".join(rows) + "\n)\n")
        w("owner", "labelList", f"{len(lown)}\n(\n" + "\n".join(map(str, lown)) + "\n)\n")
        w("neighbour", "labelList", f"{len(lnei)}\n(\n" + "\n".join(map(str, lnei)) + "\n)\n")
        body = f"{len(lblocks)}\n(\n"
        for name, typ, n, s0, q in lblocks:
            body += f"    {name}\n    {{\n        type            {typ};\n        nFaces          {n};\n        startFace       {s0};\n"
            if q is not None:
                body += f"        myProcNo        {r};\n        neighbProcNo    {q};\n"
            body += "    }\n"
        w("boundary", "polyBoundaryMesh", body + ")\n")
        w("cellProcAddressing", "labelList", f"{cells.size}\n(\n" + "\n".join(map(str, cells)) + "\n)\n")
