"""The C-ABI library loads on a CPU-only box and exports every symbol include/hopedg.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

from hopefoam_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / "include" / "hopedg.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hdg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_library):
    names = declared_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(built_library, n), f"{n} declared in hopedg.h but not exported by libhopedg.so"


def test_python_binding_covers_header():
    assert sorted(capi.SIGNATURES) == declared_functions()


def test_no_cpu_fallback(built_library):
    """Without a CUDA device hdg_create fails loudly; a host-only context refuses every compute call."""
    import torch
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    assert built_library.hdg_create(0, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in built_library.hdg_last_error(None)
    from tests.helpers import HostContext
    c = HostContext()
    c.set_order(2)
    try:
        c.state_create(4)
        raise AssertionError("compute call on a host-only context must fail")
    except capi.HdgError as e:
        assert "no CPU fallback" in str(e)


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's baseline legs may import / link / execute anything under oracle/."""
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#\s*include\s*[\"<][^\">]*oracle)|(oracle/)|(libref_cpu)", re.M)
    for p in list((ROOT / "hopefoam_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".H", ".sh") or p.name == "Makefile":
            if p.is_file():
                assert not pat.search(p.read_text()), f"{p} reaches into oracle/"


def test_header_is_valid_c99_and_links(tmp_path, built_library):
    """include/hopedg.h is a plain C header (the boundary a cgo/JNI/ctypes/Fortran binding would consume): compile a C99
    translation unit against it, link with libhopedg.so and run a host-only session (operators + connectivity, no GPU)."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include "hopedg.h"
int main(void) {
    hdg_context* c = 0;
    if (hdg_create(-1, &c)) { printf("create failed: %s\n", hdg_last_error(0)); return 1; }
    if (hdg_set_order(c, 4)) return 2;
    int32_t Np, Nfp, Ng, Nfg;
    hdg_get_sizes(c, &Np, &Nfp, &Ng, &Nfg);
    double xy[8] = {0,0, 1,0, 1,1, 0,1};
    int32_t tris[6] = {0,1,2, 0,2,3}, start[2] = {0,4}, cell[4] = {0,0,1,1}, pts[8] = {0,1, 1,2, 2,3, 3,0};
    if (hdg_set_mesh_triangles(c, 4, xy, 2, tris, 0, 1, start, cell, pts)) { printf("%s\n", hdg_last_error(c)); return 3; }
    int64_t K, F, ng; int32_t np_;
    hdg_mesh_counts(c, &K, &F, &np_, &ng);
    int32_t sid;
    int rc = hdg_state_create(c, 4, &sid);          /* must fail loudly: host-only context */
    printf("%d %d %d %d %lld %lld %d %s\n", Np, Nfp, Ng, Nfg, (long long)K, (long long)F, rc, hdg_version());
    if (rc) printf("ERR %s\n", hdg_last_error(c));
    hdg_destroy(c);
    return 0;
}
''')
    exe = tmp_path / "t"
    lib = ROOT / "hopefoam_b200"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(src), "-o", str(exe), f"-L{lib}", "-lhopedg",
                    f"-Wl,-rpath,{lib}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    assert out.startswith("15 5 54 6 2 5 1 hopedg-b200")
    assert "no CPU fallback" in out
