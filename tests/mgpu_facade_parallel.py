#!/usr/bin/env python3
"""`-parallel` runs of the facade solvers, one process per GPU (tools/hoperun; NCCL halo exchange inside libhopedg.so), on a case
decomposed by the dgDecomposePar rules, against the serial run of the same solver on the undecomposed case:

    /usr/local/graft/bin/gpurun --gpus 2 -- python tests/mgpu_facade_parallel.py [nprocs]

Checked: (i) the fields written to processorN/<time>/ equal the serial fields cell by cell (through cellProcAddressing; nodes matched by
their coordinates, the vertex order of a cell may differ between the two polyMeshes) to 1e-13; (ii) the master's printed
rhoError/rhoUError times its ownership-range end equals the serial value times the number of dofs (the reference divides the global sum
by dgMesh::localRange().second(), dgMesh.C:194-219); both for hopeEulerFoam and for the reference's unmodified tutorial solver
(oracle/_ref/dgEulerFoam) when it was built."""
import re
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hopefoam_b200 import meshgen  # noqa: E402
from oracle import dg_oracle as o  # noqa: E402
from tests.case_writer import HDR, read_field, write_euler_case  # noqa: E402
from tests.polymesh_writer import write_processor_polymeshes  # noqa: E402

N, DT, STEPS = 4, 2e-3, 10
FIELDS = {"T": ("dgScalarField", "0"), "U": ("dgVectorField", "(1 0.5 0)"), "p": ("dgScalarField", "1"), "rho": ("dgScalarField", "1"),
          "rhoU": ("dgVectorField", "(1 0 0)"), "Ener": ("dgScalarField", "3")}


def write_processor_fields(case, nprocs):
    for r in range(nprocs):
        pdir = case / f"processor{r}"
        (pdir / "0").mkdir(parents=True, exist_ok=True)
        procs = re.findall(r"(procBoundary\d+to\d+)", (pdir / "constant" / "polyMesh" / "boundary").read_text())
        for name, (cls, uni) in FIELDS.items():
            body = HDR.format(cls=cls, obj=name) + f"\ndimensions      [0 0 0 0 0 0 0];\n\ninternalField   uniform {uni};\n\nboundaryField\n{{\n"
            body += f"    boundary\n    {{\n        type            fixedValue;\n        value           uniform {uni};\n    }}\n"
            body += "    frontAndBackPlanes\n    {\n        type            empty;\n    }\n"
            for pn in procs:
                body += f"    {pn}\n    {{\n        type            processor;\n        value           uniform {uni};\n    }}\n"
            (pdir / "0" / name).write_text(body + "}\n")


def node_coords(polymesh_dir):
    om = o.mesh_from_polymesh(polymesh_dir)
    om.patches = [p for p in om.patches if p["type"] != "empty"]
    return o.Case(om, N).geo.x          # (K, Np, 2)


def errors(stdout, names=("rhoError", "rhoUError")):
    return tuple(float(re.search(name + r":\s*([0-9.eE+-]+)", stdout).group(1)) for name in names)


def check(app, nprocs, tmp, fields=(("rho", 1), ("rhoU", 3), ("Ener", 1)), err_names=("rhoError", "rhoUError"), box=None):
    mg = meshgen.jittered_square(12, **(box or {}))
    patches = [("boundary", "patch", mg["patch_edges"][0])]
    case = write_euler_case(tmp / f"case_{Path(app).name}", mg, N, DT, DT * STEPS, write_interval=STEPS)
    K = mg["tris"].shape[0]
    ser = subprocess.run([app, "-case", str(case)], capture_output=True, text=True, timeout=600)
    assert ser.returncode == 0, ser.stdout[-2000:] + ser.stderr[-2000:]
    tname = f"{DT * STEPS:.6g}"
    xg = node_coords(case / "constant" / "polyMesh")
    Np = xg.shape[1]
    glob = {f: read_field(case / tname / f, c).reshape((K, Np) + ((c,) if c > 1 else ())) for f, c in fields}
    c2p = (np.arange(K) * nprocs) // K
    write_processor_polymeshes(case, mg["xy"], mg["tris"], patches, c2p, nprocs)
    write_processor_fields(case, nprocs)
    par = subprocess.run([str(ROOT / "tools" / "hoperun"), "-np", str(nprocs), app, "-parallel", "-case", str(case)], capture_output=True, text=True,
                         timeout=900)
    assert par.returncode == 0, par.stdout[-3000:] + par.stderr[-3000:]
    worst = 0.0
    n1_master = None
    for r in range(nprocs):
        pdir = case / f"processor{r}"
        t = (pdir / "constant" / "polyMesh" / "cellProcAddressing").read_text()
        addr = np.array(t[t.index("(", t.index("// *")) + 1: t.rindex(")")].split(), dtype=int)
        if r == 0:
            n1_master = addr.size * Np
        xl = node_coords(pdir / "constant" / "polyMesh")
        d = np.linalg.norm(xl[:, :, None, :] - xg[addr][:, None, :, :], axis=-1)        # (Kr, Np local, Np global)
        perm = d.argmin(-1)
        assert d.min(-1).max() < 1e-12
        for f, c in fields:
            loc = read_field(pdir / tname / f, c).reshape((addr.size, Np) + ((c,) if c > 1 else ()))
            ref = np.take_along_axis(glob[f][addr], perm[..., None] if c > 1 else perm, axis=1)
            worst = max(worst, float(np.abs(loc - ref).max() / np.abs(ref).max()))
    es, ep = errors(ser.stdout, err_names), errors(par.stdout, err_names)
    for a, b in zip(es, ep):
        assert abs(a * K * Np - b * n1_master) <= 1e-10 * a * K * Np, (es, ep, n1_master)
    assert worst <= 1e-13, worst
    print(f"{Path(app).name}: {nprocs} ranks, fields equal the serial run to {worst:.2e}; master prints {ep} (serial {es})")


def check_with_tools(nprocs, tmp):
    """The whole workflow with the repo's own tools: hopeDgDecomposePar (system/decomposeParDict, method simple) -> hoperun -parallel ->
    hopeDgReconstructPar, against the serial run."""
    bindir = ROOT / "hopefoam_b200" / "apps" / "bin"
    mg = meshgen.jittered_square(12)
    case = write_euler_case(tmp / "case_tools", mg, N, DT, DT * STEPS, write_interval=STEPS)
    K = mg["tris"].shape[0]
    tname = f"{DT * STEPS:.6g}"
    ser = subprocess.run([str(bindir / "hopeEulerFoam"), "-case", str(case)], capture_output=True, text=True, timeout=600)
    assert ser.returncode == 0, ser.stdout[-2000:] + ser.stderr[-2000:]
    fields = (("rho", 1), ("rhoU", 3), ("Ener", 1))
    serial = {f: read_field(case / tname / f, c) for f, c in fields}
    for f, _ in fields:
        (case / tname / f).unlink()
    (case / "system" / "decomposeParDict").write_text(HDR.format(cls="dictionary", obj="decomposeParDict") +
                                                      f"\nnumberOfSubdomains {nprocs};\nmethod simple;\nsimpleCoeffs\n{{\n    n ({nprocs} 1 1);\n    delta 0.001;\n}}\n")
    for cmd in ([str(bindir / "hopeDgDecomposePar"), "-case", str(case)],
                [str(ROOT / "tools" / "hoperun"), "-np", str(nprocs), str(bindir / "hopeEulerFoam"), "-parallel", "-case", str(case)],
                [str(bindir / "hopeDgReconstructPar"), "-case", str(case), "-time", tname, "rho", "rhoU", "Ener"]):
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, " ".join(cmd) + "\n" + out.stdout[-3000:] + out.stderr[-3000:]
    worst = max(float(np.abs(read_field(case / tname / f, c) - serial[f]).max() / np.abs(serial[f]).max()) for f, c in fields)
    assert worst <= 1e-13, worst
    print(f"hopeDgDecomposePar -> hoperun -np {nprocs} hopeEulerFoam -parallel -> hopeDgReconstructPar: fields equal the serial run to {worst:.2e}")


def main():
    nprocs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    subprocess.run(["make", "-C", str(ROOT / "hopefoam_b200" / "csrc"), "apps"], check=True, capture_output=True)
    with tempfile.TemporaryDirectory() as tmp:
        check(str(ROOT / "hopefoam_b200" / "apps" / "bin" / "hopeEulerFoam"), nprocs, Path(tmp))
        check(str(ROOT / "hopefoam_b200" / "apps" / "bin" / "hopeScalarTransportFoam"), nprocs, Path(tmp), fields=(("T", 1),), err_names=("TError",),
              box=dict(x0=-1, x1=1, y0=-1, y1=1))
        check_with_tools(nprocs, Path(tmp))
        ref = ROOT / "oracle" / "_ref" / "dgEulerFoam"
        if ref.exists():
            check(str(ref), nprocs, Path(tmp))
        else:
            print("oracle/_ref/dgEulerFoam not built: reference-solver leg skipped")
    print("MGPU_FACADE_PARALLEL PASS")


if __name__ == "__main__":
    main()
