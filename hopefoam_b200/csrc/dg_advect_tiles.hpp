// Shared-memory tile geometry of the TMA advection kernels (dg_advect_tma.cu): where double d / node pair n of element e of an octet
// lands inside a tile that a TMA tensor copy has written, and which element a DMMA row carries.  Plain integer functions, host and
// device: tests/native/tile_map_check.cpp walks them on the CPU (tests/test_tile_maps_host.py: the maps are bijections onto the tile
// that agree with what the tensor copy writes, and the lanes of a quarter warp never meet in a bank group on the fragment reads).
#pragma once

#if defined(__CUDACC__)
#define HDG_TILE_HD __host__ __device__ __forceinline__
#else
#define HDG_TILE_HD inline
#endif

namespace hdg {

constexpr int kTile = 1024;        // one octet of one 128-B-row plane: 8 element rows of 128 B = one 128B-swizzle atom

// ---- 128-B rows (NpPad = 16: N = 3, 4) ---------------------------------------------------------------------------------------------
// byte offset of double `d` (0..15) of element row `e` (0..7) inside a swizzled tile: 16-B chunk index XOR row
HDG_TILE_HD int swz(int e, int d) { return e * 128 + ((((d >> 1) ^ e) & 7) << 4) + (d & 1) * 8; }
// velocity tile of an octet: 16 rows of 128 B, element e = rows 2e, 2e+1, node n = (x,y) pair in row 2e + (n >> 3), chunk n & 7
HDG_TILE_HD int swzU(int e, int n)
{
    const int row = 2 * e + (n >> 3);
    return row * 128 + ((((n & 7) ^ row) & 7) << 4);
}
// element carried by DMMA row g: rows 2q and 2q+1 of a quarter warp hold elements q and q + 4 (bit 2 of the swizzle XOR differs)
HDG_TILE_HD int elemOfRow128(int g) { return (g & 1) * 4 + (g >> 1); }

// ---- rows of NT x 64 B (NT = NpPad / 8 != 2) -----------------------------------------------------------------------------------------
template <int NT>
struct WideTile {
    static constexpr bool swizzled = (NT % 2) == 0;
    static constexpr int tBytes = 512 * NT;              // one octet of one plane
    static constexpr int oU = 0;                         // velocity pairs (2 tBytes)
    static constexpr int oTin = 2 * tBytes;
    static constexpr int oAux = 3 * tBytes;
    static constexpr int oGeo = 4 * tBytes;              // 1 KB, 128B swizzle (a multiple of 1 KB for every NT)
    static constexpr int stageBytes = 4 * tBytes + kTile;
    // element carried by DMMA row g.  NT even: the two rows of a quarter warp carry elements e, e ^ 3 - bit 1 separates them in the T
    // tile (2e enters the swizzle XOR), bit 0 in the velocity tile (4e), where both rows read the SAME own-trace nodes
    static HDG_TILE_HD int elemOfRow(int g)
    {
        if (!swizzled) return g;
        const int q = g >> 1;
        return ((q & 1) | ((q & 2) << 1)) ^ ((g & 1) ? 3 : 0);
    }
    // byte offset of double d of element e inside a T tile
    static HDG_TILE_HD int offT(int e, int d)
    {
        if (!swizzled) return e * (NT * 64) + d * 8;
        const int c = d >> 1, row = e * (NT / 2) + (c >> 3);
        return row * 128 + ((((c & 7) ^ row) & 7) << 4) + (d & 1) * 8;
    }
    // byte offset of the (x,y) pair of node n of element e inside the velocity tile: always NT swizzled 128-B rows per element (an
    // unswizzled tile would put the same node of every element on the same banks: 2-way conflicts on every own-trace read)
    static HDG_TILE_HD int offU(int e, int n)
    {
        const int row = e * NT + (n >> 3);
        return row * 128 + ((((n & 7) ^ row) & 7) << 4);
    }
};

}  // namespace hdg
