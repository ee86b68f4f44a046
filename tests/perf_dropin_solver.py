#!/usr/bin/env python3
"""Throughput of the SOLVERS on the facade (not of the bare stage): the reference's unmodified tutorial solver source
(oracle/_ref/dgEulerFoam) and hopeEulerFoam on a generated HopeFOAM case directory, one B200.

    /usr/local/graft/bin/gpurun -- python tests/perf_dropin_solver.py [n]      (n x n x 2 triangles, default 354 -> 250 632)

Per step the solver does what the reference does around the three dg::solveEquation calls: exact boundary values evaluated on the
host and sent to the patches, field copies, the SSP-RK2 combination, runTime.write().  The time per step is taken from the solver's own
`ClockTime` print-out (differences between late steps, so that mesh reading and the first-launch costs are excluded)."""
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hopefoam_b200 import meshgen  # noqa: E402
from tests.case_writer import write_euler_case  # noqa: E402

N, STEPS = 4, 30


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 354
    mg = meshgen.jittered_square(n)
    K = mg["tris"].shape[0]
    dt = 0.09 / n
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.time()
        case = write_euler_case(Path(tmp) / "case", mg, N, dt, dt * STEPS)
        print(f"case written in {time.time() - t0:.1f} s: {K} triangles, N={N}", flush=True)
        for app in (ROOT / "oracle" / "_ref" / "dgEulerFoam", ROOT / "hopefoam_b200" / "apps" / "bin" / "hopeEulerFoam"):
            if not app.exists():
                print(f"{app.name}: not built, skipped")
                continue
            t0 = time.time()
            out = subprocess.run([str(app), "-case", str(case)], capture_output=True, text=True, timeout=45)
            wall = time.time() - t0
            assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
            clock = [float(x) for x in re.findall(r"ClockTime = ([0-9.eE+-]+) s", out.stdout)]
            assert len(clock) == STEPS, len(clock)
            per_step = (clock[-1] - clock[9]) / (STEPS - 10)
            gdof = 2 * 4 * 15 * K / per_step / 1e9
            err = re.search(r"rhoError:\s*([0-9.eE+-]+)", out.stdout).group(1)
            print(f"{app.name}: {per_step * 1e3:.3f} ms per SSP-RK2 step ({gdof:.1f} GDOF-updates/s per stage), whole run incl. mesh set-up {wall:.1f} s, "
                  f"rhoError {err}", flush=True)


if __name__ == "__main__":
    main()
