#!/usr/bin/env python
"""Fluent/GAMBIT 2-D triangle mesh (.msh) -> constant/polyMesh of a one-layer prism case, the form the DG mesh reader consumes
(dgPolyMesh.C:154-190 takes the z == 0 face of each prism).  The reference's doubleMach tutorial ships only `doubleMach.msh` and relies
on OpenFOAM's fluentMeshToFoam for this step; this is the small stand-in for the 2-D triangle case.

    tools/fluentMeshToCase.py <mesh.msh> <caseDir> [--wall-type wall]

Sections read: (10 nodes) (13 faces: `2 n0 n1 c0 c1`, hex, one zone per boundary) (45 zone names).  A face's right-hand cell c0 lies to
the LEFT of n0 -> n1 in Fluent's convention; orientation is recomputed from the coordinates anyway.  Boundary zones become patches named
after the zone (type `wall` for a zone NAMED wall, else `patch`), plus the empty frontAndBackPlanes patch."""
import argparse
import re
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def parse_fluent_2d(path):
    """-> xy (P,2), tris (K,3) CCW int32, zones: list of (name, edges (m,3) int32 = (cell, pa, pb)) for the boundary zones."""
    text = Path(path).read_text()
    dim = re.search(r"\(2\s+(\d)\)", text)
    if not dim or dim.group(1) != "2":
        raise ValueError("not a 2-D Fluent mesh")
    xy = None
    for m in re.finditer(r"\(10\s*\(\s*([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s+\d+(?:\s+\d+)?\)\s*\(([^()]*)\)", text):
        zone, first, last = int(m.group(1), 16), int(m.group(2), 16), int(m.group(3), 16)
        if zone == 0:
            continue
        vals = np.array(m.group(4).split(), dtype=np.float64).reshape(-1, 2)
        if xy is None:
            xy = np.zeros((0, 2))
        if first != xy.shape[0] + 1 or vals.shape[0] != last - first + 1:
            raise ValueError("node sections must be consecutive")
        xy = np.concatenate([xy, vals])
    if xy is None:
        raise ValueError("no node section")
    names = {int(m.group(1), 16): m.group(3) for m in re.finditer(r"\(45\s*\(\s*([0-9a-fA-F]+)\s+(\S+)\s+(\S+)\s*\)\s*\(\s*\)\s*\)", text)}
    ncell = 0
    for m in re.finditer(r"\(12\s*\(\s*([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s+[0-9a-fA-F]+(?:\s+[0-9a-fA-F]+)?\)\)", text):
        if int(m.group(1), 16) == 0:
            ncell = int(m.group(3), 16)
    if ncell == 0:
        raise ValueError("no cell declaration")
    cell_pts = [[] for _ in range(ncell)]
    zones = []
    for m in re.finditer(r"\(13\s*\(\s*([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\)\s*\(([^()]*)\)", text):
        zone, ftype = int(m.group(1), 16), int(m.group(5), 16)
        rows = [ln.split() for ln in m.group(6).strip().splitlines() if ln.strip()]
        edges = []
        for r in rows:
            v = [int(t, 16) for t in r]
            if ftype == 0:                      # mixed: the node count leads every line
                if v[0] != 2:
                    raise ValueError("only 2-node faces (2-D mesh) are supported")
                v = v[1:]
            n0, n1, c0, c1 = v
            for c in (c0, c1):
                if c:
                    cell_pts[c - 1].append((n0 - 1, n1 - 1))
            if c0 == 0 or c1 == 0:
                edges.append((max(c0, c1) - 1, n0 - 1, n1 - 1))
        if edges:
            if len(edges) != len(rows):
                raise ValueError(f"zone {zone} mixes boundary and interior faces")
            zones.append((names.get(zone, f"zone{zone}"), np.array(edges, dtype=np.int32)))
    tris = np.empty((ncell, 3), dtype=np.int32)
    for c, ed in enumerate(cell_pts):
        pts = sorted({p for e in ed for p in e})
        if len(ed) != 3 or len(pts) != 3:
            raise ValueError(f"cell {c + 1} is not a triangle")
        a, b, cc = pts
        area2 = (xy[b, 0] - xy[a, 0]) * (xy[cc, 1] - xy[a, 1]) - (xy[b, 1] - xy[a, 1]) * (xy[cc, 0] - xy[a, 0])
        tris[c] = (a, b, cc) if area2 > 0 else (a, cc, b)
    return xy, tris, zones


def convert(msh, case_dir, wall_names=("wall",)):
    from tests.polymesh_writer import write_polymesh        # the ASCII polyMesh writer shared with the test cases
    xy, tris, zones = parse_fluent_2d(msh)
    patches = [(name, "wall" if name in wall_names else "patch", edges) for name, edges in zones]
    write_polymesh(Path(case_dir) / "constant" / "polyMesh", xy, tris, patches)
    return xy, tris, zones


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("msh")
    ap.add_argument("case")
    ap.add_argument("--wall-type", nargs="*", default=["wall"], help="zone names written with patch type `wall`")
    a = ap.parse_args()
    xy, tris, zones = convert(a.msh, a.case, tuple(a.wall_type))
    print(f"{xy.shape[0]} points, {tris.shape[0]} triangles, patches: " + ", ".join(f"{n} ({e.shape[0]})" for n, e in zones))
