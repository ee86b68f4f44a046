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


VIEWS = r'''
#include "dgCFD.H"
#include <chrono>
int main()
{
    const label K = 200000, Np = 15;
    Field<scalar> f(K * Np, 0.0);
    dgPatchField<scalar> pf;
    pf.setSize(100);
    pf.markClean();
    const auto t0 = std::chrono::steady_clock::now();
    double s = 0;
    for (label k = 0; k < K; ++k) {
        SubField<scalar> sub(f, Np, k * Np);      // setNonUniformInlet.H:13-15 builds three of these per cell
        sub[0] = k;
        s += sub[0];
    }
    SubList<scalar> sp(pf, 5, 10);                // setBoundaryValues.H:39-41: writing through a SubList marks the patch for upload
    sp[0] = 1;
    std::cout << "SECONDS " << std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() << " DIRTY " << pf.dirty()
              << " SUM " << s << " VALUE " << f[(K - 1) * Np] << std::endl;
    return 0;
}
'''


def test_subfield_windows_do_not_copy(tmp_path):
    """A SubField over a K*Np field must be a view: the unmodified tutorial solver builds 3 K of them at start-up (an accidental by-value
    pass made that quadratic: 25 s at 29 k cells)."""
    src, exe = tmp_path / "views.C", tmp_path / "views"
    src.write_text(VIEWS)
    subprocess.run(["g++", "-std=c++17", "-O1", f"-I{ROOT / 'hopefoam_b200' / 'include' / 'hopedg'}", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0
    tok = out.stdout.split()
    assert float(tok[1]) < 1.0 and tok[3] == "1" and float(tok[7]) == 199999.0, out.stdout
