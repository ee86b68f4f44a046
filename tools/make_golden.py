#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ from the reference's own fixtures (run in THIS container,
/root/reference is not available on the GPU box):

  * vortex{0256,1024}.npz   - the tutorial Fluent meshes TUT/isentropicVortex/vortex*.msh as (xy, tris, boundary edges)
  * doubleMach.npz          - the tutorial Fluent mesh TUT/doubleMach/doubleMach.msh (xy, tris, boundary zones far/wall/outlet/inlet)
  * cylinder_connectivity.json - dgFace counts / checksums of the only shipped polyMesh (TUT/cylinder/constant/polyMesh),
                                 computed by the oracle's restatement of the dgPolyMesh rules
  * golden_errors.json      - the reference's PUBLISHED numbers (User Guide §1.8, workshop slide 18) the oracle is pinned to
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import dg_oracle as o  # noqa: E402

TUT = Path("/root/reference/HopeFOAM-0.1/tutorials/DG/2D")
OUT = ROOT / "tests" / "golden"


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    for name in ("vortex0256", "vortex1024", "vortex4096"):
        pts, faces, zones = o.read_fluent_msh(TUT / "isentropicVortex" / f"{name}.msh")
        tris, bnd = o.triangles_from_fluent(pts, faces)
        edges = []
        for c in range(tris.shape[0]):
            for f in range(3):
                a, b = int(tris[c, f]), int(tris[c, (f + 1) % 3])
                if (min(a, b), max(a, b)) in bnd:
                    edges.append((c, a, b))
        np.savez_compressed(OUT / f"{name}.npz", xy=pts, tris=tris.astype(np.int32), patch_edges=np.array(edges, dtype=np.int32))
        print(name, tris.shape[0], "triangles", len(edges), "boundary edges")
    from tools.fluentMeshToCase import parse_fluent_2d
    xy, tris, zones = parse_fluent_2d(TUT / "doubleMach" / "doubleMach.msh")
    np.savez_compressed(OUT / "doubleMach.npz", xy=xy, tris=tris, zone_names=np.array([z[0] for z in zones]),
                        **{f"zone_{z[0]}": z[1] for z in zones})
    print("doubleMach", tris.shape[0], "triangles", [(z[0], z[1].shape[0]) for z in zones])
    m = o.mesh_from_polymesh(TUT / "cylinder" / "constant" / "polyMesh")
    h = lambda a: hashlib.sha256(np.ascontiguousarray(a, dtype=np.int32).tobytes()).hexdigest()
    cyl = {"K": int(m.K), "F": int(m.F), "interior": int((m.face_nbr >= 0).sum()),
           "patches": [[p["name"], p["type"], int(p["faces"].size)] for p in m.patches],
           "sha256": {"tris": h(m.tris), "face_owner": h(m.face_owner), "face_nbr": h(m.face_nbr), "face_loc_o": h(m.face_loc_o),
                      "face_loc_n": h(m.face_loc_n), "face_rot": h(m.face_rot),
                      "patch_faces": [h(p["faces"]) for p in m.patches]}}
    (OUT / "cylinder_connectivity.json").write_text(json.dumps(cyl, indent=1))
    print("cylinder", cyl["K"], cyl["F"], cyl["interior"], cyl["patches"])
    golden = {
        "source": "HopeFOAM-0.1_User_Guide.pdf §1.8 (p.16) and '2017-07-24-OpenFOAM workshop.pptx' slide 18; dt from User Guide Table 1.1",
        "user_guide": {"mesh": "vortex1024", "N": 4, "dt": 0.004, "endTime": 2.0,
                       "rhoError": 8.807979526797244e-06, "rhoUError": 1.865574862711117e-05},
        "slide18": [
            {"mesh": "vortex0256", "N": 1, "dt": 0.04, "rho": 1.101e-02, "rhoU": 2.361e-02},
            {"mesh": "vortex0256", "N": 2, "dt": 0.02, "rho": 3.481e-03, "rhoU": 6.384e-03},
            {"mesh": "vortex0256", "N": 3, "dt": 0.008, "rho": 6.523e-04, "rhoU": 1.610e-03},
            {"mesh": "vortex0256", "N": 4, "dt": 0.008, "rho": 2.352e-04, "rhoU": 5.187e-04},
            {"mesh": "vortex0256", "N": 5, "dt": 0.004, "rho": 6.597e-05, "rhoU": 1.388e-04},
            {"mesh": "vortex0256", "N": 6, "dt": 0.002, "rho": 2.530e-05, "rhoU": 4.675e-05},
            {"mesh": "vortex1024", "N": 1, "dt": 0.02, "rho": 3.344e-03, "rhoU": 7.182e-03},
            {"mesh": "vortex1024", "N": 2, "dt": 0.01, "rho": 3.621e-04, "rhoU": 8.060e-04},
        ],
        # the complete published table (slide 18: rho / rhoU error at t=2 for N=1..6 on the three tutorial meshes; dt = User Guide Table 1.1)
        "slide18_full": {
            "dt": {"1": [0.04, 0.02, 0.01], "2": [0.02, 0.01, 0.005], "3": [0.008, 0.004, 0.002], "4": [0.008, 0.004, 0.002],
                   "5": [0.004, 0.002, 0.001], "6": [0.002, 0.001, 0.0005]},
            "meshes": ["vortex0256", "vortex1024", "vortex4096"],
            "rho": {"1": [1.101e-02, 3.344e-03, 8.759e-04], "2": [3.481e-03, 3.621e-04, 3.972e-05], "3": [6.523e-04, 5.471e-05, 3.173e-06],
                    "4": [2.352e-04, 8.808e-06, 3.690e-07], "5": [6.597e-05, 1.477e-06, 5.505e-08], "6": [2.530e-05, 2.675e-07, 1.232e-08]},
            "rhoU": {"1": [2.361e-02, 7.182e-03, 1.793e-03], "2": [6.384e-03, 8.060e-04, 9.689e-05], "3": [1.610e-03, 1.226e-04, 7.141e-06],
                     "4": [5.187e-04, 1.866e-05, 8.533e-07], "5": [1.388e-04, 2.695e-06, 1.591e-07], "6": [4.675e-05, 4.760e-07, 3.837e-08]}}}
    (OUT / "golden_errors.json").write_text(json.dumps(golden, indent=1))


if __name__ == "__main__":
    main()
