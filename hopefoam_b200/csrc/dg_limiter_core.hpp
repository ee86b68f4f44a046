// `Triangle` slope limiter (HopeFOAM-0.1/src/DG/DG/godunovFlux/limiteSchemes/scheme/Trianglelimite/Trianglelimite.C:61-864) as five
// per-entity passes over plain arrays.  Every pass is a host+device inline function: dg_limiter.cu wraps them in kernels (one thread per
// element), tests/limiter_host_check.cpp runs the SAME functions in host loops against the numpy restatement (oracle.triangle_limit), so
// the arithmetic and the indexing are verified on the CPU.
//
// Verified against the oracle both ways: in host loops (tests/test_limiter_core_host.py) and on the device (tests/test_gpu_limiter.py).
//
// The device runs the FUSED form (dg_limiter.cu, three launches): A = passes 1+2 plus the extraction of the three vertex values of every
// field into a compact array (`vtx`), B = passes 3+4 with every element evaluating the gradients of its three faces itself in the dgFace
// owner's role (bit-identical on both sides of a face, nothing stored per face), C = pass 5.  The same functions run in host loops in
// the harness (mode 1).
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
    // work arrays, tot = K + nGhost.  Records (AoS): whatever one cell contributes to a neighbour's computation is ONE 64-B (or two
    // 32-B) gather instead of one sector per scalar
    double* cell;                      // [tot][8]: ave rho, rho u, rho v, E | centroid x, y | A0 | -
    double* V;                         // [3K][8]: (variable, direction) per owner element-face (five-pass form only)
    double* A2;                        // [3K]                                                  (five-pass form only)
    double* CV;                        // [tot][8]: cell gradients (variable, direction)
    double* vtx;                       // [K][3][4]: the four fields at the vertices v0, v1, v2 of every element (fused form), else nullptr:
                                       // the face end points are the only nodal values passes 3-4 read
    double gamma, eps, tol;
    int streamPlanes;                  // device: plane accesses carry an L2 evict-first policy (dg_limiter.cu)
    double cabc[4];                    // sum_i mpp_i (a_i, b_i, c_i) of the affine node map and sum_i mpp_i: the centroid is affine in the vertices
};

HDG_HD int64_t limTot(const LimiterView& v) { return v.K + v.nGhost; }

// n doubles (n even) from a 16-B aligned record: 16-B vector loads on the device
HDG_HD void limLoad(const double* p, int n, double* out)
{
#if defined(__CUDA_ARCH__)
    for (int i = 0; i < n; i += 2) {
        const double2 t = *reinterpret_cast<const double2*>(p + i);
        out[i] = t.x;
        out[i + 1] = t.y;
    }
#else
    for (int i = 0; i < n; ++i) out[i] = p[i];
#endif
}
HDG_HD void limStore(double* p, int n, const double* in)
{
#if defined(__CUDA_ARCH__)
    for (int i = 0; i < n; i += 2) *reinterpret_cast<double2*>(p + i) = make_double2(in[i], in[i + 1]);
#else
    for (int i = 0; i < n; ++i) p[i] = in[i];
#endif
}

// 1/x: MUFU seed + two Newton steps on the device (about 1 ulp; the IEEE division subroutine costs ~4x the FP64 pipe slots and the
// limiter divides ~40 times per element), plain division on the host
HDG_HD double limRcp(double x)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
#else
    return 1.0 / x;
#endif
}

// node of vertex vert (0,1,2): v0 = node 0, v1 = node N, v2 = node Np-1 = first nodes of faces 0, 1, 2 (triangleBaseFunction.C:79-88)
HDG_HD int limVertexNode(const LimiterView& v, int vert) { return v.nodeTab[(vert * 2) * v.NfpPad]; }
// the four fields at vertex vert of element k
HDG_HD void limVertexState(const LimiterView& v, int64_t k, int vert, double q[4])
{
    if (v.vtx) { limLoad(v.vtx + k * 12 + vert * 4, 4, q); return; }
    const int node = limVertexNode(v, vert);
    for (int f = 0; f < 4; ++f) q[f] = v.q[f][k * v.NpPad + node];
}
HDG_HD void limVertexExtract(const LimiterView& v, int64_t k)
{
    for (int vert = 0; vert < 3; ++vert)
        for (int f = 0; f < 4; ++f) v.vtx[k * 12 + vert * 4 + f] = v.q[f][k * v.NpPad + limVertexNode(v, vert)];
}

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
    double* c = v.cell + 8 * k;
    for (int f = 0; f < 4; ++f) c[f] = a[f];
    c[4] = sx;
    c[5] = sy;
    c[6] = sa;
    c[7] = 0.0;
}

// boundary state at END POINT `end` (0 = first, 1 = last trace node) of boundary face (k, lf): per field by the kind of that field's
// patch (fixedValue / processor: ghost slot; reflective: mirrored interior trace for the momentum; otherwise the interior trace)
HDG_HD void limBoundaryState(const LimiterView& v, int64_t k, int lf, int end, const double own[4], double out[4])
{
    for (int f = 0; f < 4; ++f) {
        const int* cn = (f == 1 || f == 2) ? v.connU + 4 * k : v.connS + 4 * k;
        const unsigned code = ((unsigned)cn[3] >> (8 * lf)) & 0xffu;
        if (code & kLimGhost) out[f] = v.q[f][v.ghostBase + (int64_t)cn[lf] * v.NfpPad + (end ? v.Nfp - 1 : 0)];
        else if ((code & kLimReflect) && (f == 1 || f == 2)) {
            double nx, ny;
            limNormal(v, k, lf, nx, ny);
            const double d2 = 2.0 * (own[1] * nx + own[2] * ny);
            out[f] = own[f] - d2 * (f == 1 ? nx : ny);
        } else
            out[f] = own[f];
    }
}

HDG_HD void limGhostCell(const LimiterView& v, int64_t k, int lf)
{
    const int slot = v.bslot[3 * k + lf];
    if (slot < 0) return;
    const int64_t g = v.K + slot;
    double A, B;
    limNormal(v, k, lf, A, B);
    double px, py;
    limNode(v, k, v.nodeTab[(lf * 2) * v.NfpPad], px, py);
    const double C = -px * A - py * B;
    double c[8], o[8];
    limLoad(v.cell + 8 * k, 8, c);
    o[4] = (B * B - A * A) * c[4] - 2 * A * B * c[5] - 2 * A * C;
    o[5] = (-B * B + A * A) * c[5] - 2 * A * B * c[4] - 2 * B * C;
    o[6] = o[7] = 0.0;
    const unsigned code = ((unsigned)v.connS[4 * k + 3] >> (8 * lf)) & 0xffu;
    if (code & kLimReflect) {                        // the reference tests reflective() first (:176)
        const double un = A * c[1] + B * c[2];
        o[0] = c[0];
        o[1] = c[1] - A * un;
        o[2] = c[2] - B * un;
        o[3] = c[3];
    } else if (code & kLimGhost) {                   // fixesValue(): the FIRST value of the patch field for every face of the patch
        const int64_t first = v.ghostBase + (int64_t)v.ghostFirst[slot] * v.NfpPad;
        for (int f = 0; f < 4; ++f) o[f] = v.q[f][first];
    } else {
        for (int f = 0; f < 4; ++f) o[f] = c[f];
    }
    limStore(v.cell + 8 * g, 8, o);
}

HDG_HD void limPrimitive(const LimiterView& v, double q[4])
{
    const double ir = limRcp(q[0]);
    q[1] *= ir;
    q[2] *= ir;
    q[3] = (v.gamma - 1.0) * (q[3] - 0.5 * q[0] * (q[1] * q[1] + q[2] * q[2]));
}

// cell averages (first four entries of a cell record) in primitive form (:294-304)
HDG_HD void limAvePrim(const LimiterView& v, const double a[4], double p[4])
{
    const double ir = limRcp(a[0]);
    p[0] = a[0];
    p[1] = a[1] * ir;
    p[2] = a[2] * ir;
    p[3] = (v.gamma - 1.0) * (a[3] - 0.5 * (a[1] * a[1] + a[2] * a[2]) * ir);
}

// topology of face lf of element k from ONE connectivity record: interior -> neighbour element nb, its local face nf, reversed flag,
// slot = -1; boundary -> slot = ghost slot (the only case that reads bslot).  owner: k is the dgFace owner (or the face is a boundary face)
struct LimFace { int64_t nb; int nf, rev, slot; bool owner; };
HDG_HD LimFace limFaceTopo(const LimiterView& v, int64_t k, int lf)
{
    const int* cn = v.connS + 4 * k;
    const unsigned code = ((unsigned)cn[3] >> (8 * lf)) & 0xffu;
    LimFace t;
    t.nb = cn[lf];
    t.nf = (int)(code & kLimFaceMask);
    t.rev = (code & kLimRev) ? 1 : 0;
    const bool boundary = (code & kLimGhost) || t.nb == k;      // ghost-type patch, or a patch evaluated from the interior (nb = k itself)
    t.slot = boundary ? v.bslot[3 * k + lf] : -1;
    t.owner = boundary || (code & kLimOwner);
    return t;
}

// the arithmetic of limFaceGradientValue on data already in registers: S, E = this element's end-point states (overwritten by the
// face averages in primitive form), oS, oE = the other side's, ck / cn8 = the two cell records, (x0,y0)-(x1,y1) = the face
HDG_HD void limFaceGradientCompute(const LimiterView& v, double S[4], double E[4], const double oS[4], const double oE[4], const double ck[8],
                                   const double cn8[8], double x0, double y0, double x1, double y1, double V[8])
{
    for (int f = 0; f < 4; ++f) {
        S[f] = 0.5 * S[f] + 0.5 * oS[f];
        E[f] = 0.5 * E[f] + 0.5 * oE[f];
    }
    limPrimitive(v, S);
    limPrimitive(v, E);
    const double dcx = cn8[4] - ck[4], dcy = cn8[5] - ck[5];
    const double Ad = (dcx * (y1 - y0) - (x1 - x0) * dcy) * 0.5;                                  // :428
    const double iAd = limRcp(Ad);
    double po[4], pn[4];
    limAvePrim(v, ck, po);
    limAvePrim(v, cn8, pn);
    for (int f = 0; f < 4; ++f) {
        const double dc = pn[f] - po[f], df = S[f] - E[f];
        V[2 * f] = 0.5 * (dc * (y1 - y0) + df * dcy) * iAd;
        V[2 * f + 1] = -0.5 * (dc * (x1 - x0) + df * dcx) * iAd;
    }
}

// gradients of (rho, u, v, p) on the diamond of face lf of element k and the weight A_2, evaluated in the role of element k (:341-560).
// Callers pass the dgFace owner's (k, lf) for an interior face, so that both sides of a face see bit-identical numbers; t = the face's
// topology seen from k.
HDG_HD void limFaceGradientValue(const LimiterView& v, int64_t k, int lf, const LimFace& t, double V[8], double& a2)
{
    const int slot = t.slot;
    const int vS = lf, vE = (lf + 1) % 3;                                  // the face runs from vertex lf to vertex lf+1
    double S[4], E[4], oS[4], oE[4], ck[8], cn8[8];
    limVertexState(v, k, vS, S);
    limVertexState(v, k, vE, E);
    limLoad(v.cell + 8 * k, 8, ck);
    if (slot < 0) {
        const int64_t nb = t.nb;
        const int nf = t.nf, rev = t.rev;
        limVertexState(v, nb, rev ? (nf + 1) % 3 : nf, oS);               // the neighbour's trace in this element's direction
        limVertexState(v, nb, rev ? nf : (nf + 1) % 3, oE);
        limLoad(v.cell + 8 * nb, 8, cn8);
        a2 = ck[6] + cn8[6];
    } else {
        limBoundaryState(v, k, lf, 0, S, oS);
        limBoundaryState(v, k, lf, 1, E, oE);
        limLoad(v.cell + 8 * (v.K + slot), 8, cn8);
        a2 = ck[6] + ck[6];
    }
    const double* p = v.verts + 6 * k;                                     // vertex coordinates ARE the end-point node coordinates
    limFaceGradientCompute(v, S, E, oS, oE, ck, cn8, p[2 * vS], p[2 * vS + 1], p[2 * vE], p[2 * vE + 1], V);
}

HDG_HD void limFaceGradient(const LimiterView& v, int64_t k, int lf)
{
    const LimFace t = limFaceTopo(v, k, lf);
    if (!t.owner) return;      // an interior face is handled by its dgFace owner; a boundary face by its cell
    double V[8], a2;
    limFaceGradientValue(v, k, lf, t, V, a2);
    const int64_t e = 3 * k + lf;
    for (int c = 0; c < 8; ++c) v.V[8 * e + c] = V[c];
    v.A2[e] = a2;
}

// gradient of face lf of element k evaluated from k's own side (fused form).  The formula is symmetric under exchanging the two
// elements: every difference changes sign together with the traversal direction of the face and the averages commute, so both sides
// compute bit-identical numbers (up to the vertex coordinates across a periodic wrap) - nothing needs to be stored per face and no
// data of the neighbour beyond its cell record and its two end-point states is read.  Returns the ghost slot (-1: interior face).
HDG_HD int limFaceGradientOwnSide(const LimiterView& v, int64_t k, int lf, double V[8], double& a2)
{
    const LimFace t = limFaceTopo(v, k, lf);
    limFaceGradientValue(v, k, lf, t, V, a2);
    return t.slot;
}

// fused passes 3+4 for one element: its three face gradients evaluated here, A_2-weighted mean -> CV; ghost cells as in pass 4
HDG_HD void limCellGradientFused(const LimiterView& v, int64_t k)
{
    double V[3][8], a2[3], cellA2 = 0;
    int slot[3];
    for (int lf = 0; lf < 3; ++lf) {
        slot[lf] = limFaceGradientOwnSide(v, k, lf, V[lf], a2[lf]);
        cellA2 += a2[lf];
    }
    const double iA = limRcp(cellA2);
    double s[8];
    for (int c = 0; c < 8; ++c) {
        s[c] = 0;
        for (int lf = 0; lf < 3; ++lf) s[c] += a2[lf] * V[lf][c] * iA;
    }
    limStore(v.CV + 8 * k, 8, s);
    for (int lf = 0; lf < 3; ++lf)
        if (slot[lf] >= 0) limStore(v.CV + 8 * (v.K + slot[lf]), 8, V[lf]);
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
    double cellA2 = 0;
    for (int lf = 0; lf < 3; ++lf) cellA2 += v.A2[limFaceEntry(v, k, lf)];
    const double iA = limRcp(cellA2);
    for (int c = 0; c < 8; ++c) {
        double s = 0;
        for (int lf = 0; lf < 3; ++lf) {
            const int64_t e = limFaceEntry(v, k, lf);
            s += v.A2[e] * v.V[8 * e + c] * iA;
        }
        v.CV[8 * k + c] = s;
    }
    for (int lf = 0; lf < 3; ++lf) {                   // a ghost cell takes the gradient of its face (:606-637)
        const int slot = v.bslot[3 * k + lf];
        if (slot < 0) continue;
        for (int c = 0; c < 8; ++c) v.CV[8 * (v.K + slot) + c] = v.V[8 * (3 * k + lf) + c];
    }
}

// neighbour cells of element k for the limited gradient (ghost cell behind a boundary face)
HDG_HD void limNeighbourCells(const LimiterView& v, int64_t k, int64_t c[3])
{
    for (int lf = 0; lf < 3; ++lf) {
        const LimFace t = limFaceTopo(v, k, lf);
        c[lf] = t.slot >= 0 ? v.K + t.slot : t.nb;
    }
}
// limited gradient of ONE primitive f from the gradients (gx_i, gy_i) of the three neighbour cells (:727-798): each neighbour gradient
// weighted by the squared magnitudes of the other two
HDG_HD void limLimitedGradientField(const LimiterView& v, const double gx[3], const double gy[3], double& Lx, double& Ly)
{
    double g[3];
    for (int i = 0; i < 3; ++i) g[i] = gx[i] * gx[i] + gy[i] * gy[i];
    const double fac = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
    const double iw = limRcp(fac + 3 * v.eps);
    const double w[3] = {(g[1] * g[2] + v.eps) * iw, (g[0] * g[2] + v.eps) * iw, (g[1] * g[0] + v.eps) * iw};
    Lx = w[0] * gx[0] + w[1] * gx[1] + w[2] * gx[2];
    Ly = w[0] * gy[0] + w[1] * gy[1] + w[2] * gy[2];
}

// limited gradient of the four primitives in cell k
HDG_HD void limLimitedGradient(const LimiterView& v, int64_t k, double L[8])
{
    int64_t c[3];
    limNeighbourCells(v, k, c);
    for (int f = 0; f < 4; ++f) {
        double gx[3], gy[3];
        for (int i = 0; i < 3; ++i) { gx[i] = v.CV[8 * c[i] + 2 * f]; gy[i] = v.CV[8 * c[i] + 2 * f + 1]; }
        limLimitedGradientField(v, gx, gy, L[2 * f], L[2 * f + 1]);
    }
}

// per-cell constants of the reconstruction: averages, mean velocity, centroid
HDG_HD void limCellConstants(const LimiterView& v, int64_t k, double c[8])
{
    double r[8];
    limLoad(v.cell + 8 * k, 8, r);
    c[0] = r[0]; c[1] = r[1]; c[2] = r[2]; c[3] = r[3];
    const double ir = limRcp(c[0]);
    c[4] = c[1] * ir;
    c[5] = c[2] * ir;
    c[6] = r[4];
    c[7] = r[5];
}

// P1 field about the cell averages at the point (x, y) of cell k, back to conserved variables (:803-850)
// igm1 = 1 / (gamma - 1), formed once per cell (the reference divides at every node, :846; one rounding apart)
HDG_HD void limReconstructAt(const LimiterView& v, double x, double y, const double L[8], const double c[8], double igm1, double out[4])
{
    const double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3], ub = c[4], vb = c[5];
    const double dx = x - c[6], dy = y - c[7];
    double du = dx * L[0] + dy * L[1];
    const double du1 = dx * L[2] + dy * L[3], du2 = dx * L[4] + dy * L[5], du3 = dx * L[6] + dy * L[7];
    // "crroect negative density" (:823-827).  The reference loops forever when the cell MEAN is below tol; bounded here: after
    // ~1075 halvings du is exactly 0 and the node takes the mean
    for (int it = 0; a0 + du < v.tol && it < 1200; ++it) du *= 0.5;
    out[0] = a0 + du;
    out[1] = a1 + a0 * du1 + du * ub;
    out[2] = a2 + a0 * du2 + du * vb;
    out[3] = a3 + du3 * igm1 + 0.5 * du * (ub * ub + vb * vb) + a0 * (ub * du1 + vb * du2);
}
HDG_HD void limReconstructNode(const LimiterView& v, int64_t k, int i, const double L[8], const double c[8])
{
    double o[4], x, y;
    limNode(v, k, i, x, y);
    limReconstructAt(v, x, y, L, c, 1.0 / (v.gamma - 1.0), o);
    for (int f = 0; f < 4; ++f) v.qout[f][k * v.NpPad + i] = o[f];
}

HDG_HD void limReconstruct(const LimiterView& v, int64_t k)
{
    double L[8], c[8];
    limLimitedGradient(v, k, L);
    limCellConstants(v, k, c);
    for (int i = 0; i < v.Np; ++i) limReconstructNode(v, k, i, L, c);
}

}  // namespace hdg
