"""Shared-memory tile maps of the TMA advection kernels (hopefoam_b200/csrc/dg_advect_tiles.hpp, the functions the kernels of
dg_advect_tma.cu index their tiles with) checked on the host through tests/native/tile_map_check.cpp:

* each map is a bijection onto its tile and agrees with what the TMA tensor copy writes - a row-major tile for the unswizzled maps,
  16-B chunk index XOR (128-B row index mod 8) for CU_TENSOR_MAP_SWIZZLE_128B (box rows of 128 B, tile aligned to 1 KB);
* the DMMA-row -> element map is a permutation of the octet;
* the fragment reads of a quarter warp (two DMMA rows x four lanes, served together for 16-B accesses) never meet in a bank group:
  T node pairs (chunk 4 nt + j), velocity pairs of the volume nodes (8 nt + 2 j + h, with the h swap of the odd row where the kernel
  applies it), and - for the rows that are not one 128-B line - the velocity pair of one and the same node read by both rows
  (the own-trace reads).

A wrong offset shows up in the GPU parity tests; a bank conflict does not (it only costs time), so it is pinned here."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def maps(tmp_path_factory):
    so = tmp_path_factory.mktemp("tiles") / "libtile_map_check.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Werror", str(ROOT / "tests/native/tile_map_check.cpp"), "-o", str(so)],
                   check=True)
    lib = C.CDLL(str(so))
    ip = C.POINTER(C.c_int32)
    lib.tile_maps.restype = C.c_int
    lib.tile_maps.argtypes = [C.c_int, ip, ip, ip, ip, ip]

    def get(NT):
        offT = np.zeros((8, 8 * NT), dtype=np.int32)
        offU = np.zeros((8, 8 * NT), dtype=np.int32)
        elem = np.zeros(8, dtype=np.int32)
        sw, tb = C.c_int32(0), C.c_int32(0)
        rc = lib.tile_maps(NT, offT.ctypes.data_as(ip), offU.ctypes.data_as(ip), elem.ctypes.data_as(ip), C.byref(sw), C.byref(tb))
        assert rc == 0
        return offT, offU, elem, bool(sw.value), tb.value

    return get


def _swizzle128(linear_bytes):
    """Where CU_TENSOR_MAP_SWIZZLE_128B puts byte `linear_bytes` of a row-major box of 128-B rows (tile aligned to 1 KB)."""
    row, chunk, rest = linear_bytes // 128, (linear_bytes % 128) // 16, linear_bytes % 16
    return row * 128 + ((chunk ^ (row & 7)) << 4) + rest


@pytest.mark.parametrize("NT", [1, 2, 3, 4, 5])
def test_tile_maps_are_the_tensor_copy_layout(maps, NT):
    offT, offU, elem, swizzled, tbytes = maps(NT)
    assert tbytes == 512 * NT
    assert sorted(elem.tolist()) == list(range(8))
    # T tile: element e, double d is byte 64 NT e + 8 d of the box
    lin = (np.arange(8)[:, None] * 64 * NT + np.arange(8 * NT)[None, :] * 8)
    want = np.vectorize(_swizzle128)(lin) if swizzled else lin
    assert (offT == want).all()
    assert sorted(offT.reshape(-1).tolist()) == list(range(0, 512 * NT, 8))
    # velocity tile: element e, node n is the 16-B pair at byte 128 NT e + 16 n of a box of 128-B rows, always swizzled
    linU = (np.arange(8)[:, None] * 128 * NT + np.arange(8 * NT)[None, :] * 16)
    assert (offU == np.vectorize(_swizzle128)(linU)).all()
    assert sorted(offU.reshape(-1).tolist()) == list(range(0, 1024 * NT, 16))


@pytest.mark.parametrize("NT", [1, 2, 3, 4, 5])
def test_quarter_warp_reads_are_conflict_free(maps, NT):
    offT, offU, elem, swizzled, _ = maps(NT)
    group = lambda byte: (byte // 16) % 8      # the eight 16-B bank groups of the 128-B shared-memory data path
    for q in range(4):
        rows = (2 * q, 2 * q + 1)
        for nt in range(NT):
            # T node pairs: lane (g, j) reads doubles 8 nt + 2 j, + 1 (one 16-B chunk)
            g_ = [group(offT[elem[g], 8 * nt + 2 * j]) for g in rows for j in range(4)]
            assert len(set(g_)) == 8, (NT, q, nt, g_)
            # velocity pairs of the volume nodes 8 nt + 2 j + h: both rows read the same h, except under the swizzled T layout, where
            # the odd row takes h ^ 1 first (oddU in the kernels)
            for h in range(2):
                g_ = [group(offU[elem[g], 8 * nt + 2 * j + (h ^ ((g & 1) if swizzled else 0))]) for g in rows for j in range(4)]
                assert len(set(g_)) == 8, (NT, q, nt, h, g_)
        if NT != 2:
            # own-trace reads: both rows read the pair of the SAME node; the wide-row maps keep the two elements apart
            for n in range(8 * NT):
                assert group(offU[elem[rows[0]], n]) != group(offU[elem[rows[1]], n]), (NT, q, n)
