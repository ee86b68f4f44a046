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
