// Test harness (NOT part of the product): runs the five passes of hopefoam_b200/csrc/dg_limiter_core.hpp - the inline functions the
// CUDA kernels of dg_limiter.cu wrap - in plain host loops, so that tests/test_limiter_core_host.py can compare them with the numpy
// restatement of the reference's limiter.  Built by the test with g++ into a temporary directory.
#include "../../hopefoam_b200/csrc/dg_limiter_core.hpp"

extern "C" int limiter_host_run(int64_t K, int64_t nGhost, int64_t ghostBase, int Np, int NpPad, int Nfp, int NfpPad, double* rho,
                                double* rhou, double* rhov, double* ener, const int* connS, const int* connU, const int* bslot,
                                const int* ghostFirst, const double* verts, const double* r, const double* s, const double* mpp,
                                const int* nodeTab, double* work, double gamma, double eps, double tol, int split)
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
    double* w = work;                   // same carving as hdg_euler_limit: 50 K + 14 nGhost doubles
    v.ave = w; w += 4 * tot;
    v.cx = w; w += tot;
    v.cy = w; w += tot;
    v.A0 = w; w += K;
    v.V = w; w += 8 * 3 * K;
    v.A2 = w; w += 3 * K;
    v.CV = w; w += 8 * tot;
    v.L = split ? w : nullptr;          // + 8 K doubles
    v.gamma = gamma; v.eps = eps; v.tol = tol;
    // every pass runs over ALL entities before the next one starts (one kernel launch each on the device); the loops run backwards
    // to show that no pass depends on the order inside a launch
    for (int64_t k = K - 1; k >= 0; --k) limCellAverages(v, k);
    for (int64_t k = K - 1; k >= 0; --k) for (int lf = 0; lf < 3; ++lf) limGhostCell(v, k, lf);
    for (int64_t k = K - 1; k >= 0; --k) for (int lf = 0; lf < 3; ++lf) limFaceGradient(v, k, lf);
    for (int64_t k = K - 1; k >= 0; --k) limCellGradient(v, k);
    if (!split)
        for (int64_t k = K - 1; k >= 0; --k) limReconstruct(v, k);
    else {                              // HDG_LIMITER_CFG=1: limited gradients per cell, then one thread per node slot
        for (int64_t k = K - 1; k >= 0; --k) limStoreGradient(v, k);
        for (int64_t sl = K * NpPad - 1; sl >= 0; --sl) limReconstructSlot(v, sl);
    }
    return 0;
}
