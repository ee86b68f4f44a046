// `Triangle` slope limiter (HopeFOAM-0.1/src/DG/DG/godunovFlux/limiteSchemes/scheme/Trianglelimite/Trianglelimite.C:61-864) as five
// per-entity passes over plain arrays.  Every pass is a host+device inline function: dg_limiter.cu wraps them in kernels (one thread per
// element), tests/limiter_host_check.cpp runs the SAME functions in host loops against the numpy restatement (oracle.triangle_limit), so
// the arithmetic and the indexing are verified on the CPU.
//
// Verified against the oracle both ways: in host loops (tests/test_limiter_core_host.py) and on the device (tests/test_gpu_limiter.py).
//
// Cells 0..K-1 are the elements, cells K..K+nGhost-1 the virtual cells behind the boundary faces (ghost slot order = patch order, faces
// in dgFaceIndex order, :153-258).  Passes (each needs the previous one complete for ALL entities):
//   1 cellAverages   per element        averages of (rho, rho u, rho v, E), centroid, A0                                   :100-140
//   2 ghostCell      per boundary face  mirrored centroid; averages by patch kind                                          :153-258
//   3 faceGradient   per owner face     end-point states -> primitives, diamond area, gradients of (rho, u, v, p), A_2     :341-560
//   4 cellGradient   per element        A_2-weighted mean of its three face gradients; ghost cell := its face's gradient    :570-640
//   5 reconstruct    per element        neighbour-gradient weights, P1 field about the averages, back to conserved         :727-850
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define HDG_HD __host__ __device__ __forceinline__
#else
#define HDG_HD inline
#endif

namespace hdg {

// per-face connectivity byte, as in dg_kernels.cuh (repeated here so that the header stands alone for the host check)
enum : unsigned { kLimFaceMask = 0x3, kLimRev = 0x4, kLimGhost = 0x8, kLimReflect = 0x10, kLimOwner = 0x20 };

struct LimiterView {
    int64_t K, nGhost, ghostBase;      // ghostBase: offset of the ghost region inside a plane
    int Np, NpPad, Nfp, NfpPad;
    const double* q[4];                // rho, rhoU.x, rhoU.y, Ener planes ([Kpad][NpPad] | ghost traces [nGhost][NfpPad])
    double* qout[4];                   // may alias q (pass 5 reads only the work arrays)
    const int* connS;                  // [K][4]: connectivity of the density state (patch kinds of rho decide the ghost cells, :166-176)
    const int* connU;                  // [K][4]: connectivity of the momentum state (its boundary values at the face end points)
    const int* bslot;                  // [K][3]: ghost slot of a boundary face, -1 for interior faces
    const int* ghostFirst;             // [nGhost]: first ghost slot of the patch the slot belongs to (tPatchf[0], :222-225)
    const double* verts;               // [K][6]: x0 y0 x1 y1 x2 y2
    const double* r;                   // [Np] reference nodes
    const double* s;
    const double* mpp;                 // [Np] column sums of the reference mass matrix / 2 (:109-116)
    const int* nodeTab;                // [3][2][NfpPad] faceToCellIndex
    // work arrays, tot = K + nGhost
    double* ave;                       // [4][tot]
    double* cx;                        // [tot]
    double* cy;
    double* A0;                        // [K]
    double* V;                         // [8][3K]: (variable, direction) x owner element-face
    double* A2;                        // [3K]
    double* CV;                        // [8][tot]
    double* L;                         // [8][K] limited gradients (split reconstruction only, else nullptr)
    double gamma, eps, tol;
};

HDG_HD int64_t limTot(const LimiterView& v) { return v.K + v.nGhost; }

HDG_HD void limNode(const LimiterView& v, int64_t k, int i, double& x, double& y)
{
    const double* p = v.verts + 6 * k;
    const double a = -(v.r[i] + v.s[i]) * 0.5, b = (v.r[i] + 1.0) * 0.5, c = (v.s[i] + 1.0) * 0.5;      // triangleBaseFunction.C:303-313
    x = a * p[0] + b * p[2] + c * p[4];
    y = a * p[1] + b * p[3] + c * p[5];
}

// outward unit normal of local face lf (triangleBaseFunction.C:354-391)
HDG_HD void limNormal(const LimiterView& v, int64_t k, int lf, double& nx, double& ny)
{
    const double* p = v.verts + 6 * k;
    const double xr = 0.5 * (p[2] - p[0]), yr = 0.5 * (p[3] - p[1]), xs = 0.5 * (p[4] - p[0]), ys = 0.5 * (p[5] - p[1]);
    double ax = lf == 0 ? yr : (lf == 1 ? ys - yr : -ys), ay = lf == 0 ? -xr : (lf == 1 ? xr - xs : xs);
    double len = ax * ax + ay * ay;
#if defined(__CUDA_ARCH__)
    len = sqrt(len);
#else
    len = __builtin_sqrt(len);
#endif
    nx = ax / len;
    ny = ay / len;
}

HDG_HD void limCellAverages(const LimiterView& v, int64_t k)
{
    const int64_t tot = limTot(v);
    const double* p = v.verts + 6 * k;
    const double J = 0.25 * ((p[2] - p[0]) * (p[5] - p[1]) - (p[3] - p[1]) * (p[4] - p[0]));
    double a[4] = {0, 0, 0, 0}, sx = 0, sy = 0, sa = 0;
    for (int j = 0; j < v.Np; ++j) {
        const double w = v.mpp[j];
        for (int f = 0; f < 4; ++f) a[f] += v.q[f][k * v.NpPad + j] * w;
        double x, y;
        limNode(v, k, j, x, y);
        sx += x * w;
        sy += y * w;
        sa += w * J * 2.0 / 3.0;
    }
    for (int f = 0; f < 4; ++f) v.ave[f * tot + k] = a[f];
    v.cx[k] = sx;
    v.cy[k] = sy;
    v.A0[k] = sa;
}

// boundary value of field f at trace node i of boundary face (k, lf), by the kind of that field's patch (fixedValue / processor: ghost
// slot; reflective: mirrored interior trace for the momentum; otherwise the interior trace)
HDG_HD double limBoundaryValue(const LimiterView& v, int64_t k, int lf, int f, int i)
{
    const int* cn = (f == 1 || f == 2) ? v.connU + 4 * k : v.connS + 4 * k;
    const unsigned code = ((unsigned)cn[3] >> (8 * lf)) & 0xffu;
    if (code & kLimGhost) return v.q[f][v.ghostBase + (int64_t)cn[lf] * v.NfpPad + i];
    const int node = v.nodeTab[(lf * 2) * v.NfpPad + i];
    const double own = v.q[f][k * v.NpPad + node];
    if ((code & kLimReflect) && (f == 1 || f == 2)) {
        double nx, ny;
        limNormal(v, k, lf, nx, ny);
        const double d2 = 2.0 * (v.q[1][k * v.NpPad + node] * nx + v.q[2][k * v.NpPad + node] * ny);
        return own - d2 * (f == 1 ? nx : ny);
    }
    return own;
}

HDG_HD void limGhostCell(const LimiterView& v, int64_t k, int lf)
{
    const int slot = v.bslot[3 * k + lf];
    if (slot < 0) return;
    const int64_t tot = limTot(v), g = v.K + slot;
    double A, B;
    limNormal(v, k, lf, A, B);
    double px, py;
    limNode(v, k, v.nodeTab[(lf * 2) * v.NfpPad], px, py);
    const double C = -px * A - py * B;
    v.cx[g] = (B * B - A * A) * v.cx[k] - 2 * A * B * v.cy[k] - 2 * A * C;
    v.cy[g] = (-B * B + A * A) * v.cy[k] - 2 * A * B * v.cx[k] - 2 * B * C;
    const unsigned code = ((unsigned)v.connS[4 * k + 3] >> (8 * lf)) & 0xffu;
    const double a0 = v.ave[k], a1 = v.ave[tot + k], a2 = v.ave[2 * tot + k], a3 = v.ave[3 * tot + k];
    if (code & kLimReflect) {                        // the reference tests reflective() first (:176)
        const double un = A * a1 + B * a2;
        v.ave[g] = a0;
        v.ave[tot + g] = a1 - A * un;
        v.ave[2 * tot + g] = a2 - B * un;
        v.ave[3 * tot + g] = a3;
    } else if (code & kLimGhost) {                   // fixesValue(): the FIRST value of the patch field for every face of the patch
        const int64_t first = v.ghostBase + (int64_t)v.ghostFirst[slot] * v.NfpPad;
        for (int f = 0; f < 4; ++f) v.ave[f * tot + g] = v.q[f][first];
    } else {
        v.ave[g] = a0; v.ave[tot + g] = a1; v.ave[2 * tot + g] = a2; v.ave[3 * tot + g] = a3;
    }
}

HDG_HD void limPrimitive(const LimiterView& v, double q[4])
{
    q[1] /= q[0];
    q[2] /= q[0];
    q[3] = (v.gamma - 1.0) * (q[3] - 0.5 * q[0] * (q[1] * q[1] + q[2] * q[2]));
}

// cell averages in primitive form (:294-304)
HDG_HD void limAvePrim(const LimiterView& v, int64_t c, double p[4])
{
    const int64_t tot = limTot(v);
    const double a0 = v.ave[c], a1 = v.ave[tot + c], a2 = v.ave[2 * tot + c], a3 = v.ave[3 * tot + c];
    p[0] = a0;
    p[1] = a1 / a0;
    p[2] = a2 / a0;
    p[3] = (v.gamma - 1.0) * (a3 - 0.5 * (a1 * a1 + a2 * a2) / a0);
}

HDG_HD void limFaceGradient(const LimiterView& v, int64_t k, int lf)
{
    const int* cn = v.connS + 4 * k;
    const unsigned code = ((unsigned)cn[3] >> (8 * lf)) & 0xffu;
    const int slot = v.bslot[3 * k + lf];
    if (!(code & kLimOwner) && slot < 0) return;      // an interior face is handled by its dgFace owner; a boundary face by its cell
    const int iS = v.nodeTab[(lf * 2) * v.NfpPad], iE = v.nodeTab[(lf * 2) * v.NfpPad + v.Nfp - 1];
    double S[4], E[4];
    int64_t n;
    double a2;
    if (slot < 0) {
        const int64_t nb = cn[lf];
        const int nf = code & kLimFaceMask, rev = (code & kLimRev) ? 1 : 0;
        const int jS = v.nodeTab[(nf * 2 + rev) * v.NfpPad], jE = v.nodeTab[(nf * 2 + rev) * v.NfpPad + v.Nfp - 1];
        for (int f = 0; f < 4; ++f) {
            S[f] = 0.5 * v.q[f][k * v.NpPad + iS] + 0.5 * v.q[f][nb * v.NpPad + jS];
            E[f] = 0.5 * v.q[f][k * v.NpPad + iE] + 0.5 * v.q[f][nb * v.NpPad + jE];
        }
        n = nb;
        a2 = v.A0[k] + v.A0[nb];
    } else {
        for (int f = 0; f < 4; ++f) {
            S[f] = 0.5 * v.q[f][k * v.NpPad + iS] + 0.5 * limBoundaryValue(v, k, lf, f, 0);
            E[f] = 0.5 * v.q[f][k * v.NpPad + iE] + 0.5 * limBoundaryValue(v, k, lf, f, v.Nfp - 1);
        }
        n = v.K + slot;
        a2 = v.A0[k] + v.A0[k];
    }
    limPrimitive(v, S);
    limPrimitive(v, E);
    double x0, y0, x1, y1;
    limNode(v, k, iS, x0, y0);
    limNode(v, k, iE, x1, y1);
    const double dcx = v.cx[n] - v.cx[k], dcy = v.cy[n] - v.cy[k];
    const double Ad = (dcx * (y1 - y0) - (x1 - x0) * dcy) * 0.5;                                  // :428
    double po[4], pn[4];
    limAvePrim(v, k, po);
    limAvePrim(v, n, pn);
    const int64_t e = 3 * k + lf, nE3 = 3 * v.K;
    for (int f = 0; f < 4; ++f) {
        const double dc = pn[f] - po[f], df = S[f] - E[f];
        v.V[(2 * f) * nE3 + e] = 0.5 * (dc * (y1 - y0) + df * dcy) / Ad;
        v.V[(2 * f + 1) * nE3 + e] = -0.5 * (dc * (x1 - x0) + df * dcx) / Ad;
    }
    v.A2[e] = a2;
}

// index into V / A2 of the gradient of face lf of element k (stored with the dgFace owner's element-face)
HDG_HD int64_t limFaceEntry(const LimiterView& v, int64_t k, int lf)
{
    const int* cn = v.connS + 4 * k;
    const unsigned code = ((unsigned)cn[3] >> (8 * lf)) & 0xffu;
    if ((code & kLimOwner) || v.bslot[3 * k + lf] >= 0) return 3 * k + lf;
    return 3 * (int64_t)cn[lf] + (code & kLimFaceMask);
}

HDG_HD void limCellGradient(const LimiterView& v, int64_t k)
{
    const int64_t tot = limTot(v), nE3 = 3 * v.K;
    double cellA2 = 0;
    for (int lf = 0; lf < 3; ++lf) cellA2 += v.A2[limFaceEntry(v, k, lf)];
    for (int c = 0; c < 8; ++c) {
        double s = 0;
        for (int lf = 0; lf < 3; ++lf) {
            const int64_t e = limFaceEntry(v, k, lf);
            s += v.A2[e] * v.V[c * nE3 + e] / cellA2;
        }
        v.CV[c * tot + k] = s;
    }
    for (int lf = 0; lf < 3; ++lf) {                   // a ghost cell takes the gradient of its face (:606-637)
        const int slot = v.bslot[3 * k + lf];
        if (slot < 0) continue;
        for (int c = 0; c < 8; ++c) v.CV[c * tot + v.K + slot] = v.V[c * nE3 + 3 * k + lf];
    }
}

// limited gradient of the four primitives in cell k (:727-798): each neighbour gradient weighted by the squared magnitudes of the other two
HDG_HD void limLimitedGradient(const LimiterView& v, int64_t k, double L[8])
{
    const int64_t tot = limTot(v);
    int64_t c[3];
    for (int lf = 0; lf < 3; ++lf) {
        const int slot = v.bslot[3 * k + lf];
        c[lf] = slot >= 0 ? v.K + slot : (int64_t)v.connS[4 * k + lf];
    }
    for (int f = 0; f < 4; ++f) {
        double g[3];
        for (int i = 0; i < 3; ++i) {
            const double gx = v.CV[(2 * f) * tot + c[i]], gy = v.CV[(2 * f + 1) * tot + c[i]];
            g[i] = gx * gx + gy * gy;
        }
        const double fac = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
        const double w[3] = {(g[1] * g[2] + v.eps) / (fac + 3 * v.eps), (g[0] * g[2] + v.eps) / (fac + 3 * v.eps), (g[1] * g[0] + v.eps) / (fac + 3 * v.eps)};
        for (int d = 0; d < 2; ++d)
            L[2 * f + d] = w[0] * v.CV[(2 * f + d) * tot + c[0]] + w[1] * v.CV[(2 * f + d) * tot + c[1]] + w[2] * v.CV[(2 * f + d) * tot + c[2]];
    }
}

// per-cell constants of the reconstruction: averages, mean velocity, centroid
HDG_HD void limCellConstants(const LimiterView& v, int64_t k, double c[8])
{
    const int64_t tot = limTot(v);
    c[0] = v.ave[k]; c[1] = v.ave[tot + k]; c[2] = v.ave[2 * tot + k]; c[3] = v.ave[3 * tot + k];
    c[4] = c[1] / c[0];
    c[5] = c[2] / c[0];
    c[6] = v.cx[k];
    c[7] = v.cy[k];
}

// P1 field about the cell averages at node i of cell k, back to conserved variables (:803-850)
HDG_HD void limReconstructNode(const LimiterView& v, int64_t k, int i, const double L[8], const double c[8])
{
    const double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3], ub = c[4], vb = c[5];
    double x, y;
    limNode(v, k, i, x, y);
    const double dx = x - c[6], dy = y - c[7];
    double du = dx * L[0] + dy * L[1];
    const double du1 = dx * L[2] + dy * L[3], du2 = dx * L[4] + dy * L[5], du3 = dx * L[6] + dy * L[7];
    // "crroect negative density" (:823-827).  The reference loops forever when the cell MEAN is below tol; bounded here: after
    // ~1075 halvings du is exactly 0 and the node takes the mean
    for (int it = 0; a0 + du < v.tol && it < 1200; ++it) du *= 0.5;
    v.qout[0][k * v.NpPad + i] = a0 + du;
    v.qout[1][k * v.NpPad + i] = a1 + a0 * du1 + du * ub;
    v.qout[2][k * v.NpPad + i] = a2 + a0 * du2 + du * vb;
    v.qout[3][k * v.NpPad + i] = a3 + du3 / (v.gamma - 1.0) + 0.5 * du * (ub * ub + vb * vb) + a0 * (ub * du1 + vb * du2);
}

HDG_HD void limReconstruct(const LimiterView& v, int64_t k)
{
    double L[8], c[8];
    limLimitedGradient(v, k, L);
    limCellConstants(v, k, c);
    for (int i = 0; i < v.Np; ++i) limReconstructNode(v, k, i, L, c);
}

// split form of pass 5 (one thread per cell, then one thread per node slot with coalesced stores): 5a stores the limited gradients
HDG_HD void limStoreGradient(const LimiterView& v, int64_t k)
{
    double L[8];
    limLimitedGradient(v, k, L);
    for (int c = 0; c < 8; ++c) v.L[c * v.K + k] = L[c];
}
HDG_HD void limReconstructSlot(const LimiterView& v, int64_t slot)      // slot = k * NpPad + i
{
    const int64_t k = slot / v.NpPad;
    const int i = (int)(slot - k * v.NpPad);
    if (k >= v.K || i >= v.Np) return;
    double L[8], c[8];
    for (int j = 0; j < 8; ++j) L[j] = v.L[j * v.K + k];
    limCellConstants(v, k, c);
    limReconstructNode(v, k, i, L, c);
}

}  // namespace hdg
