"""The C++ facade (dgCFD.H: dgScalarField, dgm::ddt, dgc::div, dgc::grad, dg::godunovScheme, dg::solveEquation) driven by the
hopeEulerFoam solver on a HopeFOAM-format case directory, against the oracle's restatement of dgEulerFoam's main loop."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from hopefoam_b200 import meshgen
from oracle import dg_oracle as o
from tests import helpers as H
from tests.case_writer import read_field, write_euler_case

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
APP = ROOT / "hopefoam_b200" / "apps" / "bin" / "hopeEulerFoam"


def _build():
    if not APP.exists():
        subprocess.run(["make", "-C", str(ROOT / "hopefoam_b200" / "csrc"), "apps"], check=True)


@pytest.mark.parametrize("N", [2, 4])
def test_solver_on_case_directory(tmp_path, built_library, N):
    _build()
    mg = meshgen.jittered_square(6)
    dt, steps = 2e-3, 12
    case = write_euler_case(tmp_path / "case", mg, N, dt, dt * steps)
    out = subprocess.run([str(APP), "-case", str(case)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    er = float(re.search(r"rhoError:\s*([0-9.eE+-]+)", out.stdout).group(1))
    eu = float(re.search(r"rhoUError:\s*([0-9.eE+-]+)", out.stdout).group(1))
    # oracle on the SAME polyMesh (cell vertex order as stored in the files)
    om = o.mesh_from_polymesh(case / "constant" / "polyMesh")
    om.patches = [p for p in om.patches if p["type"] != "empty"]
    run = o.VortexRun(o.Case(om, N), dt)
    for _ in range(steps):
        run.step()
    r_er, r_eu = run.errors()
    assert abs(er - r_er) <= 1e-9 * r_er and abs(eu - r_eu) <= 1e-9 * r_eu, (er, r_er, eu, r_eu)
    tdir = case / f"{dt * steps:.6g}"
    rho = read_field(tdir / "rho", 1).reshape(run.rho.shape)
    rhoU = read_field(tdir / "rhoU", 3).reshape(run.rho.shape + (3,))
    E = read_field(tdir / "Ener", 1).reshape(run.rho.shape)
    assert H.rel_l2(rho, run.rho) <= 1e-12 and H.rel_l2(rhoU[..., :2], run.rhoU) <= 1e-12 and H.rel_l2(E, run.E) <= 1e-12
    assert np.abs(rhoU[..., 2]).max() == 0.0


def test_lazy_evaluation_is_bit_identical_to_statement_by_statement(tmp_path, built_library):
    """The facade records `rho1 = rho`, the three solves of a stage and `rho = 0.5*rho + 0.5*rho1` and issues ONE launch per stage (the copy
    is never made, the combination leaves the stage kernel as a second result).  HOPEDG_LAZY=0 executes every statement where it stands
    (3 copies + 2 launches + 3 axpby per step): the written fields must be the same bits, for the reference's UNMODIFIED tutorial solver
    (oracle/_ref/dgEulerFoam) and for the doubleMach variant with its limiter calls between the stages."""
    import os
    _build()
    ref = ROOT / "oracle" / "_ref"
    mg = meshgen.jittered_square(6)
    dt, steps = 2e-3, 8
    for binary in (ref / "dgEulerFoam", APP):
        if not binary.exists():
            continue
        outs = []
        for lazy in ("1", "0"):
            case = write_euler_case(tmp_path / f"case_{binary.name}_{lazy}", mg, 3, dt, dt * steps, write_interval=steps)
            out = subprocess.run([str(binary), "-case", str(case)], capture_output=True, text=True, timeout=300, env=dict(os.environ, HOPEDG_LAZY=lazy))
            assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
            tdir = case / f"{dt * steps:.6g}"
            outs.append([read_field(tdir / f, n) for f, n in (("rho", 1), ("rhoU", 3), ("Ener", 1))])
        for a, b in zip(*outs):
            assert np.array_equal(a, b)


def test_solver_with_slip_wall_patch_types(tmp_path, built_library):
    """Boundary-condition plug-ins selected by the `type` word of each field file, as in the reference: a slip wall
    (rho, Ener: zeroGradient; rhoU: reflective) next to exact-solution fixedValue patches."""
    _build()
    N, dt, steps = 4, 2e-3, 10
    mg = meshgen.jittered_square(6)
    e = mg["patch_edges"][0]
    patches = [("wall", "wall", e[:6]), ("farField", "patch", e[6:])]
    slip = {"wall": {"rho": "zeroGradient", "rhoU": "reflective", "Ener": "zeroGradient", "T": "zeroGradient", "U": "reflective"}}
    case = write_euler_case(tmp_path / "case", mg, N, dt, dt * steps, patches=patches, bc_types=slip)
    out = subprocess.run([str(APP), "-case", str(case)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    om = o.mesh_from_polymesh(case / "constant" / "polyMesh")
    om.patches = [p for p in om.patches if p["type"] != "empty"]
    run = o.VortexRun(o.Case(om, N, bc_kinds=[o.BC_REFLECTIVE, o.BC_FIXED]), dt)
    for _ in range(steps):
        run.step()
    tdir = case / f"{dt * steps:.6g}"
    rho = read_field(tdir / "rho", 1).reshape(run.rho.shape)
    rhoU = read_field(tdir / "rhoU", 3).reshape(run.rho.shape + (3,))
    E = read_field(tdir / "Ener", 1).reshape(run.rho.shape)
    assert H.rel_l2(rho, run.rho) <= 1e-12 and H.rel_l2(rhoU[..., :2], run.rhoU) <= 1e-12 and H.rel_l2(E, run.E) <= 1e-12
    assert "type            reflective;" in (tdir / "rhoU").read_text()


def test_scalar_transport_solver_on_case_directory(tmp_path, built_library):
    """BASELINE configs[0] through the facade: dg::solveEquation(dgm::ddt(T) + dgc::div(U,T)) with `div(U,T) default LF;`."""
    _build()
    app = APP.parent / "hopeScalarTransportFoam"
    N, dt, steps = 4, 1e-3, 20
    mg = meshgen.jittered_square(7, x0=-1, x1=1, y0=-1, y1=1)
    case = write_euler_case(tmp_path / "case", mg, N, dt, dt * steps)
    out = subprocess.run([str(app), "-case", str(case)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    om = o.mesh_from_polymesh(case / "constant" / "polyMesh")
    om.patches = [p for p in om.patches if p["type"] != "empty"]
    oc = o.Case(om, N)
    x, y = oc.geo.x[..., 0], oc.geo.x[..., 1]
    exact = lambda xx, yy, t: np.exp(-((xx + 0.3 - 1.0 * t) ** 2 + (yy + 0.3 - 0.5 * t) ** 2) / (2 * 0.1 ** 2))
    T, Ux, Uy = exact(x, y, 0.0), np.full_like(x, 1.0), np.full_like(x, 0.5)
    pxy = oc.patch_internal(oc.geo.x, 0)
    bUx, bUy = [np.full(pxy.shape[0], 1.0)], [np.full(pxy.shape[0], 0.5)]
    t = 0.0
    for _ in range(steps):
        bT = [exact(pxy[:, 0], pxy[:, 1], t)]
        T1 = o.advect_stage(oc, T, Ux, Uy, bT, bUx, bUy, dt)
        T2 = o.advect_stage(oc, T1, Ux, Uy, bT, bUx, bUy, dt)
        T = 0.5 * T + 0.5 * T2
        t += dt
    got = read_field(case / f"{dt * steps:.6g}" / "T", 1).reshape(T.shape)
    assert H.rel_l2(got, T) <= 1e-12
    err = float(re.search(r"TError:\s*([0-9.eE+-]+)", out.stdout).group(1))
    assert abs(err - np.abs(T - exact(x, y, t)).sum() / T.size) <= 1e-9 * max(err, 1e-30) + 1e-16


def test_solver_error_behaviour(tmp_path, built_library):
    """FatalError conventions: unknown flux scheme / missing dictionary entry abort with the reference-style message."""
    _build()
    mg = meshgen.jittered_square(3)
    case = write_euler_case(tmp_path / "case", mg, 2, 1e-3, 2e-3)
    p = case / "system" / "dgSchemes"
    p.write_text(p.read_text().replace("fluxScheme      Roe", "fluxScheme      HLLC"))
    out = subprocess.run([str(APP), "-case", str(case)], capture_output=True, text=True, timeout=120)
    assert out.returncode != 0
    assert "FOAM FATAL ERROR" in out.stderr and "Unknown fluxSchemes type HLLC" in out.stderr and "Valid fluxSchemes types are" in out.stderr
    (case / "system" / "dgSolution").write_text("FoamFile{version 2.0; format ascii; class dictionary; object dgSolution;}\nDG { baseOrder 2; }\n")
    out = subprocess.run([str(APP), "-case", str(case)], capture_output=True, text=True, timeout=120)
    assert out.returncode != 0 and "DG parameters meshDimension have not been set" in out.stderr
