// Test harness (not shipped): the shared-memory tile maps of the TMA advection kernels (hopefoam_b200/csrc/dg_advect_tiles.hpp) walked
// on the host.  Built and driven by tests/test_tile_maps_host.py.
#include <cstdint>

#include "../../hopefoam_b200/csrc/dg_advect_tiles.hpp"

using namespace hdg;

namespace {
template <int NT>
void fill(int32_t* offT, int32_t* offU, int32_t* elem)
{
    using G = WideTile<NT>;
    for (int e = 0; e < 8; ++e) {
        for (int d = 0; d < 8 * NT; ++d) offT[e * 8 * NT + d] = G::offT(e, d);
        for (int n = 0; n < 8 * NT; ++n) offU[e * 8 * NT + n] = G::offU(e, n);
        elem[e] = G::elemOfRow(e);
    }
}
}  // namespace

extern "C" {
// offT[8][8 NT], offU[8][8 NT] (byte offsets inside the T / velocity tile), elem[8] (element of DMMA row g), swizzled flag, tile bytes.
// NT = 2 reports the 128-B-row kernel's maps (swz, swzU, elemOfRow128).
int tile_maps(int NT, int32_t* offT, int32_t* offU, int32_t* elem, int32_t* swizzledT, int32_t* tBytes)
{
    switch (NT) {
        case 1: fill<1>(offT, offU, elem); *swizzledT = WideTile<1>::swizzled; *tBytes = WideTile<1>::tBytes; return 0;
        case 3: fill<3>(offT, offU, elem); *swizzledT = WideTile<3>::swizzled; *tBytes = WideTile<3>::tBytes; return 0;
        case 4: fill<4>(offT, offU, elem); *swizzledT = WideTile<4>::swizzled; *tBytes = WideTile<4>::tBytes; return 0;
        case 5: fill<5>(offT, offU, elem); *swizzledT = WideTile<5>::swizzled; *tBytes = WideTile<5>::tBytes; return 0;
        case 2:
            for (int e = 0; e < 8; ++e) {
                for (int d = 0; d < 16; ++d) offT[e * 16 + d] = swz(e, d);
                for (int n = 0; n < 16; ++n) offU[e * 16 + n] = swzU(e, n);
                elem[e] = elemOfRow128(e);
            }
            *swizzledT = 1;
            *tBytes = kTile;
            return 0;
        default: return 1;
    }
}
}
