// Fused explicit DG stage kernels (sm_100a, FP64, DMMA.8x8x4).  See dg_kernels.cuh for the design notes and
// the reference citations.
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

#include "dg_kernels.cuh"
#include "dg_device.cuh"

namespace hdg {

// ---------------------------------------------------------------------------------------------------------
// Fused Euler stage
// ---------------------------------------------------------------------------------------------------------
// Tuning switches (A/B-measured on B200, 1 M triangles, N=4: 1.316 ms -> 1.282 ms per stage with all three):
//   HDG_VOL_PIPELINED / HDG_FACE_PIPELINED  emit the point-wise fluxes of tile/face n+1 in the same block as the projection / lift
//                                           DMMAs of tile/face n, so that dependent FP64 chains resolve under this warp's own DMMAs
//   HDG_MB4                                 resident blocks per SM at N <= 4 (3 -> 168 registers, no spills with the pipelines)
#ifndef HDG_NO_PIPELINES
#define HDG_VOL_PIPELINED
#define HDG_FACE_PIPELINED
#endif
#ifndef HDG_MB4
#define HDG_MB4 3
#endif
#ifndef HDG_MB_LOW
#define HDG_MB_LOW 3
#endif
#define HDG_EULER_MINBLOCKS(N) ((N) <= 3 ? HDG_MB_LOW : (N) <= 4 ? HDG_MB4 : ((N) <= 6 ? 2 : 1))
// N >= 9 runs the plain loops: the software pipelines double the live point-wise registers, and the accumulators alone are 112 / 144
// threads per block: at N >= 7 the operator tables (128-213 KB) allow one block per SM, so the block is widened to 8 warps
#define HDG_EULER_THREADS(N) ((N) <= 6 ? 128 : 256)      // = Dims<N>::eulerThreads
template <int N>
__global__ void __launch_bounds__(HDG_EULER_THREADS(N), HDG_EULER_MINBLOCKS(N)) eulerStageKernel(const StageParams p)
{
    using D = Dims<N>;
    extern __shared__ __align__(128) double smem[];
    const double* tab = D::big ? p.tables : smem;
    int* nodeTab = reinterpret_cast<int*>(smem + D::eulerSmemDoubles);
    __shared__ unsigned long long tableBar;
    for (int i = threadIdx.x; i < D::nodeTabInts; i += blockDim.x) nodeTab[i] = p.nodeTab[i];
    if constexpr (!D::big) stageTables(smem, p.tables, D::tableDoubles, &tableBar);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    // N >= 9: this warp's A fragments live in shared memory, [f][kt][lane]
    double* aS = smem + (D::big ? (threadIdx.x >> 5) * (4 * D::KT * 32) + lane : 0);
    const int e = lane >> 2, j = lane & 3;
    const int64_t warpsPerGrid = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t warpId = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const double gm1 = p.gamma - 1.0;

    const int64_t n1 = p.octEnd - p.octBegin, nTot = p.octList ? p.nList : n1 + (p.octEnd2 - p.octBegin2);
    auto octOf = [&](int64_t i) -> int64_t { return p.octList ? (int64_t)__ldg(p.octList + i) : (i < n1 ? p.octBegin + i : p.octBegin2 + (i - n1)); };
    for (int64_t it = warpId; it < nTot; it += warpsPerGrid) {
        const int64_t oct = octOf(it);
        const int64_t elem = oct * 8 + e;
        const bool valid = elem < p.K;
        const int64_t el = valid ? elem : p.K - 1;
        const double* geo = p.geo + el * 16;
        const int64_t eoff = el * D::NpPad;

        // A fragments of the element's nodal state: a[f][kt] = q_f[node 4*kt + j]
        double a[D::big ? 1 : 4][D::big ? 1 : D::KT];
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
            for (int kt = 0; kt < D::KT; ++kt) {
                const double v = __ldg(p.qin[f] + eoff + kt * 4 + j);
                if constexpr (D::big) aS[(f * D::KT + kt) * 32] = v;
                else a[f][kt] = v;
            }
        auto aFrag = [&](int f, int kt) -> double {
            if constexpr (D::big) return aS[(f * D::KT + kt) * 32];
            else return a[f][kt];
        };

        double acc[4][D::NT][2];
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
            for (int nt = 0; nt < D::NT; ++nt) acc[f][nt][0] = acc[f][nt][1] = 0.0;

        // connectivity now; prefetch what the surface term and the update will gather (exterior traces, q_aux) so
        // that those loads hit L1 after the volume term instead of stalling the warp on DRAM/L2 latency
        const int4 cn = __ldg(p.conn + el);
        {
#pragma unroll
            for (int face = 0; face < 3; ++face) {
                const int nb = face == 0 ? cn.x : (face == 1 ? cn.y : cn.z);
                const unsigned code = ((unsigned)cn.w >> (8 * face)) & 0xffu;
                const bool ghost = code & kCodeGhost;
                const int64_t nbBase = ghost ? p.ghostBase + (int64_t)nb * D::NfpPad : (int64_t)nb * D::NpPad;
                const int* nt_ = nodeTab + ((code & kCodeFaceMask) * 2 + ((code & kCodeRev) ? 1 : 0)) * D::NfpPad;
                // lane j touches the first / last trace node of field j: covers the (at most two) 128-B lines of a trace
                const int i0 = (j & 1) ? D::Nfp - 1 : 0;
                const int64_t off = nbBase + (ghost ? i0 : nt_[i0]);
                prefetchL1((ghost ? ((j >> 1) ? p.qghost[1] : p.qghost[0]) : ((j >> 1) ? p.qin[1] : p.qin[0])) + off);
                prefetchL1((ghost ? ((j >> 1) ? p.qghost[3] : p.qghost[2]) : ((j >> 1) ? p.qin[3] : p.qin[2])) + off);
            }
            if (p.mode == 0 && p.A != 0.0) prefetchL1((j == 0 ? p.qaux[0] : j == 1 ? p.qaux[1] : j == 2 ? p.qaux[2] : p.qaux[3]) + eoff);
        }

        // ---- volume term -----------------------------------------------------------------------------
        {
            const double2 g01 = __ldg(reinterpret_cast<const double2*>(geo));
            const double2 g23 = __ldg(reinterpret_cast<const double2*>(geo) + 1);
            const double rx = g01.x, ry = g01.y, sx = g23.x, sy = g23.y;
            auto interp = [&](int gt, double (&c)[4][2]) {
#pragma unroll
                for (int f = 0; f < 4; ++f) c[f][0] = c[f][1] = 0.0;
                const double* tv = tab + D::oVg + gt * D::KT * 32 + lane;
#pragma unroll
                for (int kt = 0; kt < D::KT; ++kt) {
                    const double b = tv[kt * 32];
#pragma unroll
                    for (int f = 0; f < 4; ++f) dmma(c[f], aFrag(f, kt), b);
                }
            };
            auto project = [&](int gt, const double (&Gr)[2][4], const double (&Gs)[2][4]) {
                const double* tr = tab + D::oPr + gt * 2 * D::NT * 32 + lane;
                const double* ts = tab + D::oPs + gt * 2 * D::NT * 32 + lane;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int nt = 0; nt < D::NT; ++nt) {
                        const double br = tr[(h * D::NT + nt) * 32];
#pragma unroll
                        for (int f = 0; f < 4; ++f) dmma(acc[f][nt], Gr[h][f], br);
                    }
#pragma unroll
                    for (int nt = 0; nt < D::NT; ++nt) {
                        const double bs = ts[(h * D::NT + nt) * 32];
#pragma unroll
                        for (int f = 0; f < 4; ++f) dmma(acc[f][nt], Gs[h][f], bs);
                    }
                }
            };
#ifdef HDG_VOL_PIPELINED
          if constexpr (!D::big) {
            // software pipeline: the point-wise fluxes of tile gt+1 are independent of the projection DMMAs of tile gt and are
            // emitted in the same block, so their dependent FP64 chains resolve while the DMMAs of this warp occupy the pipe
            double c[4][2], Gr[2][4], Gs[2][4];
            interp(0, c);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double q[4] = {c[0][h], c[1][h], c[2][h], c[3][h]};
                eulerVolumeFlux(q, rx, ry, sx, sy, gm1, Gr[h], Gs[h]);
            }
#pragma unroll 1
            for (int gt = 0; gt < D::GT; ++gt) {
                double Gr2[2][4], Gs2[2][4];
                if (gt + 1 < D::GT) {
                    interp(gt + 1, c);
                    project(gt, Gr, Gs);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const double q[4] = {c[0][h], c[1][h], c[2][h], c[3][h]};
                        eulerVolumeFlux(q, rx, ry, sx, sy, gm1, Gr2[h], Gs2[h]);
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int f = 0; f < 4; ++f) { Gr[h][f] = Gr2[h][f]; Gs[h][f] = Gs2[h][f]; }
                } else
                    project(gt, Gr, Gs);
            }
          } else
#endif
          {
#pragma unroll 1
            for (int gt = 0; gt < D::GT; ++gt) {
                double c[4][2];
                interp(gt, c);
                double Gr[2][4], Gs[2][4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double q[4] = {c[0][h], c[1][h], c[2][h], c[3][h]};
                    eulerVolumeFlux(q, rx, ry, sx, sy, gm1, Gr[h], Gs[h]);
                }
                project(gt, Gr, Gs);
            }
          }
        }

#ifndef HDG_NO_NEXT_PREFETCH
        {   // pull the next octet of this warp (state lines, geometry, connectivity) towards L1 while the surface term runs
            const int64_t itn = it + warpsPerGrid;
            if (itn < nTot) {
                const int64_t octn = octOf(itn);
                const int64_t eln = min(octn * 8 + e, p.K - 1);
                prefetchL1((j == 0 ? p.qin[0] : j == 1 ? p.qin[1] : j == 2 ? p.qin[2] : p.qin[3]) + eln * D::NpPad);
                if (j == 0) prefetchL1(p.geo + eln * 16);
                if (j == 1) prefetchL1(p.conn + eln);
            }
        }
#endif
        // ---- surface term ----------------------------------------------------------------------------
        // interior (own) and exterior traces as A fragments over the face nodes, both in this element's traversal
        // direction (ownerDofMapping / rotated neighborDofMapping, physicalFaceElement.C:80-93); the gathers of face f+1
        // are issued before face f is processed (software pipelining: their latency hides behind the Roe flux of face f)
        auto loadTraces = [&](int face, double (&am_)[4][D::FKT], double (&an_)[4][D::FKT]) {
            const int nb = face == 0 ? cn.x : (face == 1 ? cn.y : cn.z);
            const unsigned code = ((unsigned)cn.w >> (8 * face)) & 0xffu;
            const bool ghost = code & kCodeGhost;
            const int64_t nbBase = ghost ? p.ghostBase + (int64_t)nb * D::NfpPad : (int64_t)nb * D::NpPad;
            const int* nt_ = nodeTab + ((code & kCodeFaceMask) * 2 + ((code & kCodeRev) ? 1 : 0)) * D::NfpPad;
            const int* no_ = nodeTab + (face * 2) * D::NfpPad;
            const double* qn[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) qn[f] = ghost ? p.qghost[f] : p.qin[f];
#pragma unroll
            for (int fkt = 0; fkt < D::FKT; ++fkt) {
                const int i = fkt * 4 + j;
                const bool in = i < D::Nfp;
                const int64_t off = nbBase + (ghost ? i : nt_[in ? i : 0]);
                const int offO = no_[in ? i : 0];
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    an_[f][fkt] = in ? __ldg(qn[f] + off) : 0.0;
                    am_[f][fkt] = in ? __ldg(p.qin[f] + eoff + offO) : 0.0;
                }
            }
        };
        auto faceInterp = [&](int fgt, const double (&am_)[4][D::FKT], const double (&an_)[4][D::FKT], double (&cm)[4][2], double (&cp)[4][2]) {
#pragma unroll
            for (int f = 0; f < 4; ++f) cm[f][0] = cm[f][1] = cp[f][0] = cp[f][1] = 0.0;
            const double* ti = tab + D::oIf + fgt * D::FKT * 32 + lane;
#pragma unroll
            for (int fkt = 0; fkt < D::FKT; ++fkt) {
                const double b = ti[fkt * 32];
#pragma unroll
                for (int f = 0; f < 4; ++f) dmma(cm[f], am_[f][fkt], b);
#pragma unroll
                for (int f = 0; f < 4; ++f) dmma(cp[f], an_[f][fkt], b);
            }
        };
        auto faceFlux = [&](int face, const double (&cm)[4][2], const double (&cp)[4][2], double (&fl)[2][4]) {
            const unsigned code = ((unsigned)cn.w >> (8 * face)) & 0xffu;
            const double2 nxy = __ldg(reinterpret_cast<const double2*>(geo + kGeoN) + face);
            const double nx = nxy.x, ny = nxy.y, fs = __ldg(geo + kGeoFs + face);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double qM[4] = {cm[0][h], cm[1][h], cm[2][h], cm[3][h]};
                double qP[4] = {cp[0][h], cp[1][h], cp[2][h], cp[3][h]};
                if (code & kCodeReflect) {      // transform(I - 2nn, trace) on the momentum (reflectiveDgPatchField.C:140-147)
                    const double d2 = 2.0 * (qP[1] * nx + qP[2] * ny);
                    qP[1] -= d2 * nx;
                    qP[2] -= d2 * ny;
                }
                // evaluate in the dgFace owner's orientation on both sides (one flux per face, flipped for the
                // neighbour: defaultConvectionScheme.C:114-127), branch-free
                const bool own = code & kCodeOwner;
                double qA[4], qB[4];
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    qA[f] = own ? qM[f] : qP[f];
                    qB[f] = own ? qP[f] : qM[f];
                }
                const double sg = own ? 1.0 : -1.0;
                eulerFaceFluxPoint(p.fluxKind, qA, qB, sg * nx, sg * ny, gm1, fl[h]);
                const double sc = sg * fs;
#pragma unroll
                for (int f = 0; f < 4; ++f) fl[h][f] *= sc;
            }
        };
        auto faceLift = [&](int face, int fgt, const double (&fl)[2][4]) {
            const double* tl = tab + D::oLift + (face * D::FGT + fgt) * 2 * D::NT * 32 + lane;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int nt = 0; nt < D::NT; ++nt) {
                    const double b = tl[(h * D::NT + nt) * 32];
#pragma unroll
                    for (int f = 0; f < 4; ++f) dmma(acc[f][nt], fl[h][f], b);
                }
        };
#ifdef HDG_FACE_PIPELINED
        if constexpr (D::FGT == 1) {
            // software pipeline over the faces: the Roe flux of face f+1 is independent of the lift DMMAs of face f and is emitted
            // in the same block; the gathers of face f+2 are issued one face ahead
            double am[4][D::FKT], an[4][D::FKT], cm[4][2], cp[4][2], fl[2][4];
            loadTraces(0, am, an);
            faceInterp(0, am, an, cm, cp);
            loadTraces(1, am, an);
            faceFlux(0, cm, cp, fl);
#pragma unroll 1
            for (int face = 0; face < 3; ++face) {
                if (face < 2) {
                    faceInterp(0, am, an, cm, cp);
                    if (face < 1) loadTraces(face + 2, am, an);
                    faceLift(face, 0, fl);
                    double fl2[2][4];
                    faceFlux(face + 1, cm, cp, fl2);
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int f = 0; f < 4; ++f) fl[h][f] = fl2[h][f];
                } else
                    faceLift(face, 0, fl);
            }
        } else
#endif
        {
            double amN[4][D::FKT], anN[4][D::FKT];
            loadTraces(0, amN, anN);
#pragma unroll 1
            for (int face = 0; face < 3; ++face) {
                double am[4][D::FKT], an[4][D::FKT];
#pragma unroll
                for (int f = 0; f < 4; ++f)
#pragma unroll
                    for (int fkt = 0; fkt < D::FKT; ++fkt) { am[f][fkt] = amN[f][fkt]; an[f][fkt] = anN[f][fkt]; }
                if (face < 2) loadTraces(face + 1, amN, anN);
#pragma unroll
                for (int fgt = 0; fgt < D::FGT; ++fgt) {
                    double cm[4][2], cp[4][2], fl[2][4];
                    faceInterp(fgt, am, an, cm, cp);
                    faceFlux(face, cm, cp, fl);
                    faceLift(face, fgt, fl);
                }
            }
        }

        // ---- explicit update (mass solve folded into Pr/Ps/LIFT) ---------------------------------------
        // per field: all loads first (independent requests in flight), then the arithmetic, then the stores
        if (valid) {
            const int64_t off0 = eoff + 2 * j;
            if (p.mode == 0) {
                const bool useAux = p.A != 0.0;
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    double2 qi[D::NT], qa[D::NT];
#pragma unroll
                    for (int nt = 0; nt < D::NT; ++nt) {
                        qi[nt] = __ldg(reinterpret_cast<const double2*>(p.qin[f] + off0 + nt * 8));
                        qa[nt] = useAux ? __ldg(reinterpret_cast<const double2*>(p.qaux[f] + off0 + nt * 8)) : make_double2(0.0, 0.0);
                    }
                    if (p.qout2[0]) {      // second result, from the same q_in and L
#pragma unroll
                        for (int nt = 0; nt < D::NT; ++nt) {
                            const double2 q2 = p.A2 != 0.0 ? __ldg(reinterpret_cast<const double2*>(p.qaux2[f] + off0 + nt * 8)) : make_double2(0.0, 0.0);
                            double2 o;
                            o.x = p.B2 * (qi[nt].x + p.dt * acc[f][nt][0]) + p.A2 * q2.x;
                            o.y = p.B2 * (qi[nt].y + p.dt * acc[f][nt][1]) + p.A2 * q2.y;
                            *reinterpret_cast<double2*>(p.qout2[f] + off0 + nt * 8) = o;
                        }
                    }
#pragma unroll
                    for (int nt = 0; nt < D::NT; ++nt) {
                        double2 o;
                        o.x = p.B * (qi[nt].x + p.dt * acc[f][nt][0]) + p.A * qa[nt].x;
                        o.y = p.B * (qi[nt].y + p.dt * acc[f][nt][1]) + p.A * qa[nt].y;
                        *reinterpret_cast<double2*>(p.qout[f] + off0 + nt * 8) = o;
                    }
                }
            } else {
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    double2 qi[D::NT], r[D::NT];
#pragma unroll
                    for (int nt = 0; nt < D::NT; ++nt) {
                        qi[nt] = __ldg(reinterpret_cast<const double2*>(p.qin[f] + off0 + nt * 8));
                        r[nt] = *reinterpret_cast<const double2*>(p.res[f] + off0 + nt * 8);
                    }
#pragma unroll
                    for (int nt = 0; nt < D::NT; ++nt) {
                        r[nt].x = p.A * r[nt].x + p.dt * acc[f][nt][0];
                        r[nt].y = p.A * r[nt].y + p.dt * acc[f][nt][1];
                        *reinterpret_cast<double2*>(p.res[f] + off0 + nt * 8) = r[nt];
                        *reinterpret_cast<double2*>(p.qout[f] + off0 + nt * 8) = make_double2(qi[nt].x + p.B * r[nt].x, qi[nt].y + p.B * r[nt].y);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Fused scalar-advection stage (nodal collapse of the quadrature form; exact for nodal U*T)
// ---------------------------------------------------------------------------------------------------------
#ifndef HDG_ADV_MB
#define HDG_ADV_MB 3
#endif
template <int N>
// N <= 2: 4 resident blocks (128 registers, no spills): 0.134 -> 0.119 ms and 0.136 -> 0.122 ms per 1 M-triangle stage; N = 5, 6 lose with 4
__global__ void __launch_bounds__(128, (N <= 2 ? 4 : (N <= 6 ? HDG_ADV_MB : 1))) advectStageKernel(const AdvectParams p)
{
    using D = Dims<N>;
    extern __shared__ __align__(128) double smem[];
    double* tab = smem;
    int* nodeTab = reinterpret_cast<int*>(smem + D::advTableDoubles);
    // per-warp staging tile of the octet's own nodal values [plane 0..2][element 0..7][NpPad]: the interior face traces are
    // read back from here (2 shared-memory wavefronts per request) instead of re-gathering them from L1 (8 lines per request)
    constexpr int OS = D::NpPad + 2;   // element stride of the tile: +2 doubles spreads the 8 elements over all banks, keeps 16-B alignment
    double* own = smem + D::advTableDoubles + (D::nodeTabInts + 1) / 2 + (threadIdx.x >> 5) * (3 * 8 * OS);
    __shared__ unsigned long long tableBar;
    for (int i = threadIdx.x; i < D::nodeTabInts; i += blockDim.x) nodeTab[i] = p.nodeTab[i];
    stageTables(tab, p.tables, D::advTableDoubles, &tableBar);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int e = lane >> 2, j = lane & 3;
    const int64_t warpsPerGrid = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t warpId = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nOct = (p.K + 7) >> 3;
    const int64_t PSU = p.planeStrideU;

    // This stage is HBM-bound by arithmetic (3 flop/B) but in practice limited by the L1 request rate: a request whose 32
    // lanes touch 8 different 128-B lines (one per element of the octet) costs 8 L1 wavefronts.  So: (i) the nodal values
    // are fetched as 16-B vectors (node pair 2j,2j+1 of each 8-node block - the SAME layout as the update's double2, so
    // they double as q_in of the update; the K index of the operator tables is permuted to match), (ii) interior traces
    // come from a shared-memory tile, (iii) all remaining loads of the octet are issued in one burst, with the
    // connectivity of the warp's next octet fetched one iteration ahead.
    // volume operator fragments stay in registers across octets when they are few (N <= 4: 16 doubles): 56 fewer shared-memory
    // wavefronts per octet on the L1 data pipe that bounds this kernel (0.2245 -> 0.2185 ms at N=4)
    constexpr bool kRegTab = 4 * D::NT * D::NT <= 16;
    double bwr[kRegTab ? 2 * D::NT : 1][kRegTab ? D::NT : 1], bws[kRegTab ? 2 * D::NT : 1][kRegTab ? D::NT : 1];
    if constexpr (kRegTab) {
#pragma unroll
        for (int kt = 0; kt < 2 * D::NT; ++kt)
#pragma unroll
            for (int nt = 0; nt < D::NT; ++nt) {
                bwr[kt][nt] = tab[D::oDwr + (kt * D::NT + nt) * 32 + lane];
                bws[kt][nt] = tab[D::oDws + (kt * D::NT + nt) * 32 + lane];
            }
    }
    int4 cT = make_int4(0, 0, 0, 0), cU = cT;
    if (warpId < nOct) {
        const int64_t el0 = min(warpId * 8 + e, p.K - 1);
        cT = __ldg(p.connT + el0);
        cU = __ldg(p.connU + el0);
    }
    for (int64_t oct = warpId; oct < nOct; oct += warpsPerGrid) {
        const int64_t elem = oct * 8 + e;
        const bool valid = elem < p.K;
        const int64_t el = valid ? elem : p.K - 1;
        const double* geo = p.geo + el * 16;
        const int64_t off0 = el * D::NpPad + 2 * j;

        // ---- burst of loads --------------------------------------------------------------------------------------
        double2 Tq[D::NT], Uxq[D::NT], Uyq[D::NT], qx[D::NT];
#pragma unroll
        for (int nt = 0; nt < D::NT; ++nt) {
            Tq[nt] = __ldg(reinterpret_cast<const double2*>(p.Tin + off0 + nt * 8));
            Uxq[nt] = __ldg(reinterpret_cast<const double2*>(p.U + off0 + nt * 8));
            Uyq[nt] = __ldg(reinterpret_cast<const double2*>(p.U + PSU + off0 + nt * 8));
        }
        double TN[3][D::FKT], uxN[3][D::FKT], uyN[3][D::FKT];
#pragma unroll
        for (int face = 0; face < 3; ++face) {
            const int nbT = face == 0 ? cT.x : (face == 1 ? cT.y : cT.z);
            const int nbU = face == 0 ? cU.x : (face == 1 ? cU.y : cU.z);
            const unsigned codeT = ((unsigned)cT.w >> (8 * face)) & 0xffu, codeU = ((unsigned)cU.w >> (8 * face)) & 0xffu;
            const bool ghT = codeT & kCodeGhost, ghU = codeU & kCodeGhost;
            const int64_t baseT = ghT ? p.ghostBase + (int64_t)nbT * D::NfpPad : (int64_t)nbT * D::NpPad;
            const int64_t baseU = ghU ? p.ghostBase + (int64_t)nbU * D::NfpPad : (int64_t)nbU * D::NpPad;
            const int* ntT = nodeTab + ((codeT & kCodeFaceMask) * 2 + ((codeT & kCodeRev) ? 1 : 0)) * D::NfpPad;
            const int* ntU = nodeTab + ((codeU & kCodeFaceMask) * 2 + ((codeU & kCodeRev) ? 1 : 0)) * D::NfpPad;
#pragma unroll
            for (int fkt = 0; fkt < D::FKT; ++fkt) {
                {
                    const int i = fkt * 4 + j;
                    const int ii = i < D::Nfp ? i : 0;
                    const int64_t oT = baseT + (ghT ? ii : ntT[ii]), oU = baseU + (ghU ? ii : ntU[ii]);
                    TN[face][fkt] = __ldg(p.Tin + oT);
                    uxN[face][fkt] = __ldg(p.U + oU);
                    uyN[face][fkt] = __ldg(p.U + PSU + oU);
                }
            }
        }
        const unsigned codesU = (unsigned)cU.w;
#pragma unroll
        for (int nt = 0; nt < D::NT; ++nt) {
            if (p.mode == 0) qx[nt] = p.A != 0.0 ? __ldg(reinterpret_cast<const double2*>(p.Taux + off0 + nt * 8)) : make_double2(0.0, 0.0);
            else             qx[nt] = *reinterpret_cast<const double2*>(p.res + off0 + nt * 8);
        }
        // geometry record (16 doubles): lane j fetches doubles 4j..4j+3, the 4 lanes of an element swap them by shuffles
        const double2 gq0 = __ldg(reinterpret_cast<const double2*>(geo) + 2 * j);
        const double2 gq1 = __ldg(reinterpret_cast<const double2*>(geo) + 2 * j + 1);
        {   // connectivity of the next octet, one iteration ahead
            const int64_t octn = oct + warpsPerGrid;
            if (octn < nOct) {
                const int64_t eln = min(octn * 8 + e, p.K - 1);
                cT = __ldg(p.connT + eln);
                cU = __ldg(p.connU + eln);
            }
        }
        double g[14];
#pragma unroll
        for (int i = 0; i < 14; ++i) {
            const double v = (i & 3) == 0 ? gq0.x : ((i & 3) == 1 ? gq0.y : ((i & 3) == 2 ? gq1.x : gq1.y));
            g[i] = __shfl_sync(0xffffffffu, v, (lane & ~3) | (i >> 2));
        }
        const double rx = g[0], ry = g[1], sx = g[2], sy = g[3];

        // stage the own nodal values for the trace reads
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < D::NT; ++nt) {
            *reinterpret_cast<double2*>(own + (0 * 8 + e) * OS + nt * 8 + 2 * j) = Tq[nt];
            *reinterpret_cast<double2*>(own + (1 * 8 + e) * OS + nt * 8 + 2 * j) = Uxq[nt];
            *reinterpret_cast<double2*>(own + (2 * 8 + e) * OS + nt * 8 + 2 * j) = Uyq[nt];
        }
        __syncwarp();

        double acc[D::NT][2];
#pragma unroll
        for (int nt = 0; nt < D::NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;

        // volume: rhs += Dwr (rx Ux T + ry Uy T) + Dws (sx Ux T + sy Uy T)   (defaultConvectionScheme.C:247-262)
        // k-tile (2*nt' + h), slot j  <->  node 8*nt' + 2*j + h   (tables built with the same permutation)
#pragma unroll
        for (int ntp = 0; ntp < D::NT; ++ntp)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kt = 2 * ntp + h;
                const double T = h ? Tq[ntp].y : Tq[ntp].x, ux = h ? Uxq[ntp].y : Uxq[ntp].x, uy = h ? Uyq[ntp].y : Uyq[ntp].x;
                const double fx = ux * T, fy = uy * T;
                const double ar = rx * fx + ry * fy, as = sx * fx + sy * fy;
#pragma unroll
                for (int nt = 0; nt < D::NT; ++nt) {
                    if constexpr (kRegTab) {
                        dmma(acc[nt], ar, bwr[kt][nt]);
                        dmma(acc[nt], as, bws[kt][nt]);
                    } else {
                        dmma(acc[nt], ar, tab[D::oDwr + (kt * D::NT + nt) * 32 + lane]);
                        dmma(acc[nt], as, tab[D::oDws + (kt * D::NT + nt) * 32 + lane]);
                    }
                }
            }

        // surface: nodal LF / average flux, lifted with LIFTn (LFFlux.C:147-206)
#pragma unroll
        for (int face = 0; face < 3; ++face) {
            const unsigned codeU = (codesU >> (8 * face)) & 0xffu;
            const double nx = g[kGeoN + 2 * face], ny = g[kGeoN + 2 * face + 1], fs = g[kGeoFs + face];
            const int* ntO = nodeTab + (face * 2) * D::NfpPad;
            double vO[D::FKT], vN[D::FKT], TO[D::FKT];
            double maxV = 0.0;
#pragma unroll
            for (int fkt = 0; fkt < D::FKT; ++fkt) {
                const int i = fkt * 4 + j;
                const int no = ntO[i < D::Nfp ? i : 0];
                TO[fkt] = own[(0 * 8 + e) * OS + no];
                const double uxo = own[(1 * 8 + e) * OS + no], uyo = own[(2 * 8 + e) * OS + no];
                double uxn = uxN[face][fkt], uyn = uyN[face][fkt];
                if (codeU & kCodeReflect) {
                    const double d2 = 2.0 * (uxn * nx + uyn * ny);
                    uxn -= d2 * nx;
                    uyn -= d2 * ny;
                }
                vO[fkt] = nx * uxo + ny * uyo;
                vN[fkt] = nx * uxn + ny * uyn;
                if (i < D::Nfp) maxV = fmax(maxV, fmax(fabs(vO[fkt]), fabs(vN[fkt])));
            }
            maxV = fmax(maxV, __shfl_xor_sync(0xffffffffu, maxV, 1));     // one maxV per face (LFFlux.C:189-196)
            maxV = fmax(maxV, __shfl_xor_sync(0xffffffffu, maxV, 2));
            const double diss = p.fluxKind == 1 ? maxV : 0.0;
#pragma unroll
            for (int fkt = 0; fkt < D::FKT; ++fkt) {
                const int i = fkt * 4 + j;
                double fl = (vO[fkt] * TO[fkt] + vN[fkt] * TN[face][fkt]) * 0.5 + diss * (TO[fkt] - TN[face][fkt]) * 0.5;
                fl = (i < D::Nfp && p.fluxKind != 3) ? fl * fs : 0.0;
#pragma unroll
                for (int nt = 0; nt < D::NT; ++nt) dmma(acc[nt], fl, tab[D::oLiftN + ((face * D::FKT + fkt) * D::NT + nt) * 32 + lane]);
            }
        }

        if (valid) {
#pragma unroll
            for (int nt = 0; nt < D::NT; ++nt) {
                double2 o;
                if (p.mode == 0) {
                    o.x = p.B * (Tq[nt].x + p.dt * acc[nt][0]) + p.A * qx[nt].x;
                    o.y = p.B * (Tq[nt].y + p.dt * acc[nt][1]) + p.A * qx[nt].y;
                } else {
                    double2 r;
                    r.x = p.A * qx[nt].x + p.dt * acc[nt][0];
                    r.y = p.A * qx[nt].y + p.dt * acc[nt][1];
                    *reinterpret_cast<double2*>(p.res + off0 + nt * 8) = r;
                    o.x = Tq[nt].x + p.B * r.x;
                    o.y = Tq[nt].y + p.B * r.y;
                }
                *reinterpret_cast<double2*>(p.Tout + off0 + nt * 8) = o;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Layout / utility kernels
// ---------------------------------------------------------------------------------------------------------

// host AoS (node-contiguous per element, stride hostStride) -> device plane [Kpad][NpPad]
__global__ void aosToPlaneKernel(const double* __restrict__ src, int hostStride, double* __restrict__ dst, int64_t K, int Np, int NpPad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * NpPad) return;
    const int64_t k = i / NpPad;
    const int n = (int)(i - k * NpPad);
    dst[i] = n < Np ? src[(k * Np + n) * hostStride] : 0.0;
}

__global__ void planeToAosKernel(const double* __restrict__ src, double* __restrict__ dst, int hostStride, int64_t K, int Np, int NpPad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * Np) return;
    const int64_t k = i / Np;
    const int n = (int)(i - k * Np);
    dst[i * hostStride] = src[k * NpPad + n];
}

// patch dof values (nFaces*Nfp, stride) -> ghost slots [ghostStart + face][NfpPad]
__global__ void patchToGhostKernel(const double* __restrict__ src, int hostStride, double* __restrict__ ghost, int64_t nFaces, int Nfp,
                                   int NfpPad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nFaces * NfpPad) return;
    const int64_t f = i / NfpPad;
    const int n = (int)(i - f * NfpPad);
    ghost[i] = n < Nfp ? src[(f * Nfp + n) * hostStride] : 0.0;
}

// the same for nPlanes planes and BOTH copies of a state in one launch: ghost0 / ghost1 = first ghost slot of the patch in plane 0 of the
// current / stage copy, planeStride between planes; component c of the source goes to plane c
__global__ void patchToGhostAllKernel(const double* __restrict__ src, int hostStride, double* __restrict__ ghost0, double* __restrict__ ghost1,
                                      int64_t planeStride, int nPlanes, int64_t nFaces, int Nfp, int NfpPad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = nFaces * NfpPad;
    if (i >= per * nPlanes) return;
    const int c = (int)(i / per);
    const int64_t r = i - c * per;
    const int64_t f = r / NfpPad;
    const int n = (int)(r - f * NfpPad);
    const double v = n < Nfp ? src[(f * Nfp + n) * hostStride + c] : 0.0;
    ghost0[c * planeStride + r] = v;
    ghost1[c * planeStride + r] = v;
}

// dst = a*x + b*y over whole planes (ghost slots included); dst may alias x or y
// (x[i], y[i]) pairs of two planes: the velocity layout of the TMA advection kernel
__global__ void zipPlanesKernel(const double* __restrict__ x, const double* __restrict__ y, double2* __restrict__ out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_double2(x[i], y[i]);
}

__global__ void axpbyKernel(double* __restrict__ dst, double a, const double* x, double b, const double* y, int64_t n)
{
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < n) {
        const double2 vx = *reinterpret_cast<const double2*>(x + i), vy = *reinterpret_cast<const double2*>(y + i);
        *reinterpret_cast<double2*>(dst + i) = make_double2(a * vx.x + b * vy.x, a * vx.y + b * vy.y);
    } else if (i < n)
        dst[i] = a * x[i] + b * y[i];
}

// sum |q - ref| over the real nodes: one block-level partial per block (deterministic two-pass reduction)
__global__ void l1DiffKernel(const double* __restrict__ q, const double* __restrict__ ref, int64_t K, int Np, int NpPad,
                             double* __restrict__ partial)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < K * Np; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i / Np;
        const int n = (int)(i - k * Np);
        s += fabs(q[k * NpPad + n] - ref[k * NpPad + n]);
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// halo pack: owner-side nodal trace of every plane on the faces of a processor patch, reversed per face
// (processorDgPatchField.C:240-262).  faceElem/faceLoc: per patch face the owner element and local face.
__global__ void haloPackKernel(const double* __restrict__ q, int64_t planeStride, int nPlanes, const int* __restrict__ faceElem,
                               const int* __restrict__ faceLoc, const int* __restrict__ nodeTab, int64_t nFaces, int Nfp, int NfpPad,
                               int NpPad, double* __restrict__ buf, int rev)
{
    // rev = 1: trace in the neighbour's traversal direction (processor halo); rev = 0: in the owner's own direction (frozen traces)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = nFaces * NfpPad;
    if (i >= per * nPlanes) return;
    const int pl = (int)(i / per);
    const int64_t r = i - pl * per;
    const int64_t f = r / NfpPad;
    const int n = (int)(r - f * NfpPad);
    double v = 0.0;
    if (n < Nfp) v = q[pl * planeStride + (int64_t)faceElem[f] * NpPad + nodeTab[(faceLoc[f] * 2 + rev) * NfpPad + n]];
    buf[i] = v;
}

__global__ void haloUnpackKernel(const double* __restrict__ buf, double* __restrict__ q, int64_t planeStride, int nPlanes,
                                 int64_t ghostOff, int64_t nFaces, int NfpPad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = nFaces * NfpPad;
    if (i >= per * nPlanes) return;
    const int pl = (int)(i / per);
    const int64_t r = i - pl * per;
    q[pl * planeStride + ghostOff + r] = buf[i];
}

// The same for ALL processor patches of a context in one launch, planes given by pointer (the four conserved planes may live in up to
// three states).  Message layout [face][plane][NfpPad]: the faces of one neighbour are contiguous (patch order), so the segment of a
// neighbour is one contiguous message.  faceGhost: ghost slot of every face (unpack side).
__global__ void haloPackAllKernel(HaloPlanes q, int nPlanes, const int* __restrict__ faceElem, const int* __restrict__ faceLoc,
                                  const int* __restrict__ nodeTab, int64_t nFaces, int Nfp, int NfpPad, int NpPad, double* __restrict__ buf)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nFaces * nPlanes * NfpPad) return;
    const int n = (int)(i % NfpPad);
    const int64_t r = i / NfpPad;
    const int pl = (int)(r % nPlanes);
    const int64_t f = r / nPlanes;
    double v = 0.0;
    if (n < Nfp) v = q.p[pl][(int64_t)faceElem[f] * NpPad + nodeTab[(faceLoc[f] * 2 + 1) * NfpPad + n]];
    buf[i] = v;
}

__global__ void haloUnpackAllKernel(const double* __restrict__ buf, HaloPlanes q, int nPlanes, const int* __restrict__ faceGhost,
                                    int64_t ghostBase, int64_t nFaces, int NfpPad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nFaces * nPlanes * NfpPad) return;
    const int n = (int)(i % NfpPad);
    const int64_t r = i / NfpPad;
    const int pl = (int)(r % nPlanes);
    const int64_t f = r / nPlanes;
    q.p[pl][ghostBase + (int64_t)faceGhost[f] * NfpPad + n] = buf[i];
}

// ---------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------
template <int N>
static void launchEulerT(const StageParams& p, int grid, cudaStream_t st)
{
    using D = Dims<N>;
    const size_t smem = sizeof(double) * D::eulerSmemDoubles + sizeof(int) * D::nodeTabInts;
    static bool configured[64] = {};          // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t err = cudaFuncSetAttribute(eulerStageKernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute(euler): ") + cudaGetErrorString(err));
        configured[dev & 63] = true;
    }
    eulerStageKernel<N><<<grid, HDG_EULER_THREADS(N), smem, st>>>(p);
}

template <int N>
static void launchAdvectT(const AdvectParams& p, int grid, cudaStream_t st)
{
    using D = Dims<N>;
    const size_t smem = sizeof(double) * (D::advTableDoubles + (D::nodeTabInts + 1) / 2 + 4 * 3 * 8 * (D::NpPad + 2));
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t err = cudaFuncSetAttribute(advectStageKernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute(advect): ") + cudaGetErrorString(err));
        configured[dev & 63] = true;
    }
    advectStageKernel<N><<<grid, 128, smem, st>>>(p);
}

void launchEulerStage(int N, const StageParams& p, int grid, cudaStream_t st)
{
    switch (N) {
        case 1: launchEulerT<1>(p, grid, st); break;
        case 2: launchEulerT<2>(p, grid, st); break;
        case 3: launchEulerT<3>(p, grid, st); break;
        case 4: launchEulerT<4>(p, grid, st); break;
        case 5: launchEulerT<5>(p, grid, st); break;
        case 6: launchEulerT<6>(p, grid, st); break;
        case 7: launchEulerT<7>(p, grid, st); break;
        case 8: launchEulerT<8>(p, grid, st); break;
        case 9: launchEulerT<9>(p, grid, st); break;
        case 10: launchEulerT<10>(p, grid, st); break;
        default: throw std::runtime_error("unsupported order");
    }
}

bool launchAdvectStageTma(int N, const AdvectParams& p, cudaStream_t st);      // dg_advect_tma.cu

void launchAdvectStage(int N, const AdvectParams& p, int grid, cudaStream_t st)
{
    if (launchAdvectStageTma(N, p, st)) return;      // N = 3, 4: TMA-pipelined kernel; other orders: advectStageKernel below
    switch (N) {
        case 1: launchAdvectT<1>(p, grid, st); break;
        case 2: launchAdvectT<2>(p, grid, st); break;
        case 3: launchAdvectT<3>(p, grid, st); break;
        case 4: launchAdvectT<4>(p, grid, st); break;
        case 5: launchAdvectT<5>(p, grid, st); break;
        case 6: launchAdvectT<6>(p, grid, st); break;
        case 7: launchAdvectT<7>(p, grid, st); break;
        case 8: launchAdvectT<8>(p, grid, st); break;
        case 9: launchAdvectT<9>(p, grid, st); break;
        case 10: launchAdvectT<10>(p, grid, st); break;
        default: throw std::runtime_error("unsupported order");
    }
}

template <int N>
static void occT(int* eulerBlocks, size_t* eulerSmem, int* advBlocks, size_t* advSmem)
{
    using D = Dims<N>;
    *eulerSmem = sizeof(double) * D::eulerSmemDoubles + sizeof(int) * D::nodeTabInts;
    *advSmem = sizeof(double) * (D::advTableDoubles + (D::nodeTabInts + 1) / 2 + 4 * 3 * 8 * (D::NpPad + 2));
    cudaFuncSetAttribute(eulerStageKernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*eulerSmem);
    cudaFuncSetAttribute(advectStageKernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*advSmem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(eulerBlocks, eulerStageKernel<N>, HDG_EULER_THREADS(N), *eulerSmem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(advBlocks, advectStageKernel<N>, 128, *advSmem);
}

int eulerWarpsPerBlock(int N) { return HDG_EULER_THREADS(N) / 32; }

void stageOccupancy(int N, int* eulerBlocks, size_t* eulerSmem, int* advBlocks, size_t* advSmem)
{
    switch (N) {
        case 1: occT<1>(eulerBlocks, eulerSmem, advBlocks, advSmem); break;
        case 2: occT<2>(eulerBlocks, eulerSmem, advBlocks, advSmem); break;
        case 3: occT<3>(eulerBlocks, eulerSmem, advBlocks, advSmem); break;
        case 4: occT<4>(eulerBlocks, eulerSmem, advBlocks, advSmem); break;
        case 5: occT<5>(eulerBlocks, eulerSmem, advBlocks, advSmem); break;
        case 6: occT<6>(eulerBlocks, eulerSmem, advBlocks, advSmem); break;
        case 7: occT<7>(eulerBlocks, eulerSmem, advBlocks, advSmem); break;
        case 8: occT<8>(eulerBlocks, eulerSmem, advBlocks, advSmem); break;
        case 9: occT<9>(eulerBlocks, eulerSmem, advBlocks, advSmem); break;
        case 10: occT<10>(eulerBlocks, eulerSmem, advBlocks, advSmem); break;
        default: throw std::runtime_error("unsupported order");
    }
}

void launchAosToPlane(const double* src, int hostStride, double* dst, int64_t K, int Np, int NpPad, cudaStream_t st)
{
    const int64_t n = K * NpPad;
    aosToPlaneKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, hostStride, dst, K, Np, NpPad);
}
void launchPlaneToAos(const double* src, double* dst, int hostStride, int64_t K, int Np, int NpPad, cudaStream_t st)
{
    const int64_t n = K * Np;
    planeToAosKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst, hostStride, K, Np, NpPad);
}
void launchPatchToGhost(const double* src, int hostStride, double* ghost, int64_t nFaces, int Nfp, int NfpPad, cudaStream_t st)
{
    const int64_t n = nFaces * NfpPad;
    if (n == 0) return;
    patchToGhostKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, hostStride, ghost, nFaces, Nfp, NfpPad);
}
void launchPatchToGhostAll(const double* src, int hostStride, double* ghost0, double* ghost1, int64_t planeStride, int nPlanes, int64_t nFaces, int Nfp,
                           int NfpPad, cudaStream_t st)
{
    const int64_t n = nFaces * NfpPad * nPlanes;
    if (n == 0) return;
    patchToGhostAllKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, hostStride, ghost0, ghost1, planeStride, nPlanes, nFaces, Nfp, NfpPad);
}
void launchZipPlanes(const double* x, const double* y, double* out, int64_t n, cudaStream_t st)
{
    zipPlanesKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, reinterpret_cast<double2*>(out), n);
}
void launchAxpby(double* dst, double a, const double* x, double b, const double* y, int64_t n, cudaStream_t st)
{
    axpbyKernel<<<(unsigned)((n / 2 + 256) / 256), 256, 0, st>>>(dst, a, x, b, y, n);
}
void launchL1Diff(const double* q, const double* ref, int64_t K, int Np, int NpPad, double* partial, int nBlocks, cudaStream_t st)
{
    l1DiffKernel<<<nBlocks, 256, 0, st>>>(q, ref, K, Np, NpPad, partial);
}
void launchHaloPack(const double* q, int64_t planeStride, int nPlanes, const int* faceElem, const int* faceLoc, const int* nodeTab,
                    int64_t nFaces, int Nfp, int NfpPad, int NpPad, double* buf, cudaStream_t st, int rev)
{
    const int64_t n = nFaces * NfpPad * nPlanes;
    if (n == 0) return;
    haloPackKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q, planeStride, nPlanes, faceElem, faceLoc, nodeTab, nFaces, Nfp, NfpPad,
                                                               NpPad, buf, rev);
}
void launchHaloUnpack(const double* buf, double* q, int64_t planeStride, int nPlanes, int64_t ghostOff, int64_t nFaces, int NfpPad,
                      cudaStream_t st)
{
    const int64_t n = nFaces * NfpPad * nPlanes;
    if (n == 0) return;
    haloUnpackKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(buf, q, planeStride, nPlanes, ghostOff, nFaces, NfpPad);
}

// FP64 pipe peak, measured live next to the stage kernel (bench.py): 8 independent DMMA.8x8x4 accumulator chains per warp, no memory
// traffic.  DMMA and DFMA share one pipe on B200 (tools/microbench/fp64_peak.cu): 64 FMA per clock per SM.
__global__ void __launch_bounds__(512) fp64PeakKernel(double* out, double a, double b, int iters)
{
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[2 * i]), "+d"(c[2 * i + 1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// returns the flops of one launch (2 * 256 FMA per DMMA)
double launchFp64Peak(double* out, int blocks, int iters, cudaStream_t st)
{
    fp64PeakKernel<<<blocks, 512, 0, st>>>(out, 1.0000001, 1e-9, iters);
    return 2.0 * 256.0 * 8.0 * iters * (double)blocks * 16.0;
}

void launchHaloPackAll(const HaloPlanes& q, int nPlanes, const int* faceElem, const int* faceLoc, const int* nodeTab, int64_t nFaces, int Nfp,
                       int NfpPad, int NpPad, double* buf, cudaStream_t st)
{
    const int64_t n = nFaces * nPlanes * NfpPad;
    if (n == 0) return;
    haloPackAllKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q, nPlanes, faceElem, faceLoc, nodeTab, nFaces, Nfp, NfpPad, NpPad, buf);
}
void launchHaloUnpackAll(const double* buf, const HaloPlanes& q, int nPlanes, const int* faceGhost, int64_t ghostBase, int64_t nFaces, int NfpPad,
                         cudaStream_t st)
{
    const int64_t n = nFaces * nPlanes * NfpPad;
    if (n == 0) return;
    haloUnpackAllKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(buf, q, nPlanes, faceGhost, ghostBase, nFaces, NfpPad);
}

}  // namespace hdg
