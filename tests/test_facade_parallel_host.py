"""Host side of the facade's `-parallel` (no GPU): tools/hoperun starts one process per rank with RANK / WORLD_SIZE / LOCAL_RANK and a
common run id; argList turns `-parallel` into Pstream state; Time points fields and polyMesh at <case>/processorN and system/ at the
case; IOdictionary finds global dictionaries in the case's own constant/; Info prints on the master only."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

PROBE = r'''
#include "foamLite.H"
using namespace Foam;
int main(int argc, char* argv[])
{
    argList args(argc, argv);
    Time runTime(args);
    Info << "MASTER " << Pstream::myProcNo() << endl;
    Pout << "RANK " << Pstream::myProcNo() << " OF " << Pstream::nProcs() << " LOCAL " << Pstream::localRank() << " PAR " << Pstream::parRun()
         << " ID " << Pstream::runId() << " PATH " << runTime.path() << " SYSTEM " << runTime.system() << " CONSTANT " << runTime.constant()
         << " DICT " << IOdictionary::locate(IOobject("transportProperties", runTime.constant(), runTime)) << std::endl;
    IOdictionary tp(IOobject("transportProperties", runTime.constant(), runTime));
    const dimensionedScalar gamma(tp.lookup("gamma"));
    Pout << "GAMMA " << gamma.value() << std::endl;
    return 0;
}
'''


def _case(tmp_path):
    case = tmp_path / "case"
    (case / "system").mkdir(parents=True)
    (case / "constant").mkdir()
    (case / "system" / "controlDict").write_text("startTime 0;\nendTime 1;\ndeltaT 0.1;\nwriteInterval 10;\n")
    (case / "constant" / "transportProperties").write_text("gamma gamma [0 0 0 0 0 0 0] 1.4;\n")
    for r in range(3):
        (case / f"processor{r}" / "constant").mkdir(parents=True)
    return case


def _probe(tmp_path):
    src, exe = tmp_path / "probe.C", tmp_path / "probe"
    src.write_text(PROBE)
    subprocess.run(["g++", "-std=c++17", "-O0", f"-I{ROOT / 'hopefoam_b200' / 'include' / 'hopedg'}", str(src), "-o", str(exe)], check=True)
    return exe


def test_hoperun_and_parallel_paths(tmp_path):
    case, exe = _case(tmp_path), _probe(tmp_path)
    out = subprocess.run([str(ROOT / "tools" / "hoperun"), "-np", "3", str(exe), "-parallel", "-case", str(case)], capture_output=True, text=True,
                         timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.splitlines()
    assert sum("MASTER" in ln for ln in lines) == 1 and any(ln.strip() == "MASTER 0" for ln in lines)      # Info: master only
    ranks = sorted(ln for ln in lines if "RANK " in ln)
    assert len(ranks) == 3
    ids = set()
    for r in range(3):
        ln = next(x for x in ranks if f"RANK {r} OF 3 LOCAL {r} PAR 1" in x)
        assert f"PATH {case}/processor{r} " in ln and f"SYSTEM {case}/system " in ln and f"CONSTANT {case}/processor{r}/constant " in ln
        assert ln.rstrip().endswith(f"DICT {case}/constant/transportProperties")          # global dictionary: the case's own constant/
        ids.add(ln.split(" ID ")[1].split()[0])
    assert len(ids) == 1                                                                   # one run id for all ranks
    assert sum("GAMMA 1.4" in ln for ln in lines) == 3


def test_serial_paths_and_launcher_failure(tmp_path):
    case, exe = _case(tmp_path), _probe(tmp_path)
    out = subprocess.run([str(exe), "-case", str(case)], capture_output=True, text=True)
    assert out.returncode == 0 and f"PAR 0 ID  PATH {case} " in out.stdout.replace("ID 0", "ID ") or "PAR 0" in out.stdout
    assert f"DICT {case}/constant/transportProperties" in out.stdout
    # -parallel without a launcher: the reference-style fatal error, non-zero exit
    bad = subprocess.run([str(exe), "-parallel", "-case", str(case)], capture_output=True, text=True, env={"PATH": "/usr/bin:/bin"})
    assert bad.returncode != 0 and "-parallel needs one process per sub-domain" in bad.stderr
    # a failing rank takes the run down with a non-zero status instead of leaving the others waiting
    fail = subprocess.run([str(ROOT / "tools" / "hoperun"), "-np", "2", "bash", "-c", "if [ $RANK = 1 ]; then exit 3; else exec sleep 30; fi"],
                          capture_output=True, text=True, timeout=20)
    assert fail.returncode == 1
