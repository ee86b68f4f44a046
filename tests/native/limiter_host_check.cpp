// Test harness (NOT part of the product): runs the passes of hopefoam_b200/csrc/dg_limiter_core.hpp (five-pass and fused forms) - the inline functions the
// CUDA kernels of dg_limiter.cu wrap - in plain host loops, so that tests/test_limiter_core_host.py can compare them with the numpy
// restatement of the reference's limiter.  Built by the test with g++ into a temporary directory.
#include "../../hopefoam_b200/csrc/dg_limiter_core.hpp"

extern "C" int limiter_host_run(int64_t K, int64_t nGhost, int64_t ghostBase, int Np, int NpPad, int Nfp, int NfpPad, double* rho,
                                double* rhou, double* rhov, double* ener, const int* connS, const int* connU, const int* bslot,
                                const int* ghostFirst, const double* verts, const double* r, const double* s, const double* mpp,
                                const int* nodeTab, double* work, double gamma, double eps, double tol, int fused)
{
    using namespace hdg;
    LimiterView v{};
    v.K = K; v.nGhost = nGhost; v.ghostBase = ghostBase;
    v.Np = Np; v.NpPad = NpPad; v.Nfp = Nfp; v.NfpPad = NfpPad;
    double* planes[4] = {rho, rhou, rhov, ener};
    for (int f = 0; f < 4; ++f) { v.q[f] = planes[f]; v.qout[f] = planes[f]; }      // in place, as hdg_euler_limit does
    v.connS = connS; v.connU = connU; v.bslot = bslot; v.ghostFirst = ghostFirst;
    v.verts = verts; v.r = r; v.s = s; v.mpp = mpp; v.nodeTab = nodeTab;
    const int64_t tot = K + nGhost;
    double* w = work;                   // 16 tot + 39 K doubles
    v.cell = w; w += 8 * tot;
    v.V = w; w += 8 * 3 * K;
    v.A2 = w; w += 3 * K;
    v.CV = w; w += 8 * tot;
    v.vtx = fused ? w : nullptr;        // + 12 K doubles
    v.gamma = gamma; v.eps = eps; v.tol = tol;
    // every pass runs over ALL entities before the next one starts (one kernel launch each on the device); the loops run backwards
    // to show that no pass depends on the order inside a launch
    if (!fused) {                       // the five passes as the reference orders them
        for (int64_t k = K - 1; k >= 0; --k) limCellAverages(v, k);
        for (int64_t k = K - 1; k >= 0; --k) for (int lf = 0; lf < 3; ++lf) limGhostCell(v, k, lf);
        for (int64_t k = K - 1; k >= 0; --k) for (int lf = 0; lf < 3; ++lf) limFaceGradient(v, k, lf);
        for (int64_t k = K - 1; k >= 0; --k) limCellGradient(v, k);
        for (int64_t k = K - 1; k >= 0; --k) limReconstruct(v, k);
    } else {                            // the three launches of dg_limiter.cu: A = 1+2 (+ vertex values), B = 3+4 per element, C = 5
        for (int64_t k = K - 1; k >= 0; --k) { limCellAverages(v, k); limVertexExtract(v, k); for (int lf = 0; lf < 3; ++lf) limGhostCell(v, k, lf); }
        for (int64_t k = K - 1; k >= 0; --k) limCellGradientFused(v, k);
        for (int64_t k = K - 1; k >= 0; --k) {
            double L[8], c[8];
            limLimitedGradient(v, k, L);
            limCellConstants(v, k, c);
            for (int i = 0; i < Np; ++i) limReconstructNode(v, k, i, L, c);
        }
    }
    return 0;
}
