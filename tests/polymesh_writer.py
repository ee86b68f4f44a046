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
