// Split Euler stage (sm_100a, FP64, DMMA.8x8x4): the surface flux of every dgFace is evaluated ONCE.
//
// The fused stage (dg_kernels.cu, eulerStageKernel) evaluates the Roe flux on both sides of every interior face, because an element
// lane cannot see the result its neighbour's warp computes: 60 % of its point-wise FP64 work, and - since those dependent chains run
// at a fifth of the pipe rate next to other warps' DMMAs - almost half of a warp's time.  The reference computes one flux per dgFace,
// at the owner's Gauss points and in the owner's orientation, and hands it to both cells (defaultConvectionScheme.C:114-127,
// RoeFlux.C:46-191).  This file does the same in two launches per stage:
//
//   eulerFaceFluxKernel<N>   one warp = 8 dgFaces on the DMMA M axis: owner / exterior traces (ghost slots, reflective mirror and the
//                            rotated neighbour map exactly as loadTraces of the fused kernel: the per-state connectivity codes decide),
//                            2*FKT*4 DMMA per face tile interpolate both sides to the face Gauss points, one Roe evaluation per point,
//                            flux[face][field][slot] <- F*.n_owner                              (192 B per face at N = 4)
//   eulerElemKernel<N>       the fused kernel's volume term and update; its surface term is 3 gathers of flux records (the neighbour
//                            reads the points in reverse order and flips the sign), scaled by Fscale, and the lift DMMAs.
//
// The extra HBM traffic (flux written once, read twice; the state read a second time by the face kernel) is paid from the 84 % of the
// HBM roofline the FP64-bound stage leaves unused.
#include <cuda_runtime.h>

#include <algorithm>
#include <stdexcept>
#include <string>

#include "dg_kernels.cuh"
#include "dg_device.cuh"

namespace hdg {

#ifndef HDG_SPLIT_EARLY_MIN
#define HDG_SPLIT_EARLY_MIN 5      // orders from here to 6 fetch the next octet's fragments and the flux records early (costs ~16 registers);
                                   // N <= 4 instead run 16 warps per SM at 128 registers (A/B on a B200, profiles/experiments_r02.md: +1.5..3 %)
#endif
// Block shapes.  Both kernels run ONE block per SM holding all of the SM's warps (element kernel: 16 warps at 128 registers for N <= 4,
// 8 warps at 255 for N >= 5; face kernel: 16 warps for N <= 6, 12 at 168 registers above).  Until the end of round 2 the same warps
// came as 4 (2, 3) blocks of 128 threads: co-resident blocks of a persistent grid are 148 block ids apart, so an SM worked on octets
// from four distant places of the mesh; the warps of one block walk CONSECUTIVE octets, and the rows / flux records that neighbouring
// octets share are L1 hits (N=4: 1.029 -> 0.973 ms per stage, 0.758 -> 0.800 of the FP64 peak; N=3: 0.753 -> 0.712 ms).  The *_LOW /
// *_56 / *_HIGH macros restore the small blocks for A/B runs (tools/build_variant.sh); threads per SM stay the same.
#ifndef HDG_SPLIT_THREADS_LOW
#define HDG_SPLIT_THREADS_LOW 512
#endif
#ifndef HDG_SPLIT_THREADS_56
#define HDG_SPLIT_THREADS_56 256
#endif
#define HDG_SPLIT_THREADS(N) ((N) <= 4 ? HDG_SPLIT_THREADS_LOW : ((N) <= 6 ? HDG_SPLIT_THREADS_56 : 256))
#define HDG_SPLIT_MINBLOCKS(N) ((N) <= 4 ? 512 / HDG_SPLIT_THREADS_LOW : ((N) <= 6 ? 256 / HDG_SPLIT_THREADS_56 : 1))
#ifndef HDG_FACE_THREADS_LOW
#define HDG_FACE_THREADS_LOW 512
#endif
#ifndef HDG_FACE_THREADS_HIGH
#define HDG_FACE_THREADS_HIGH 384
#endif
#define HDG_FACE_THREADS(N) ((N) <= 6 ? HDG_FACE_THREADS_LOW : HDG_FACE_THREADS_HIGH)
#define HDG_FACE_MINBLOCKS(N) ((N) <= 6 ? 512 / HDG_FACE_THREADS_LOW : 384 / HDG_FACE_THREADS_HIGH)

// -------------------------------------------------------------------------------------------------------------------------------
// Face kernel
// -------------------------------------------------------------------------------------------------------------------------------
// FLUX: 0 Roe, 1 point-wise local Lax-Friedrichs (compile time: a run-time switch between the two inlined fluxes costs the N >= 7
// kernels their register fit)
template <int N, int FLUX>
__global__ void __launch_bounds__(HDG_FACE_THREADS(N), HDG_FACE_MINBLOCKS(N)) eulerFaceFluxKernel(const StageParams p)
{
    using D = Dims<N>;
    constexpr int SL = D::fluxSlots;
    __shared__ int nodeTab[D::nodeTabInts];
    for (int i = threadIdx.x; i < D::nodeTabInts; i += blockDim.x) nodeTab[i] = p.nodeTab[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int e = lane >> 2, j = lane & 3;
    const double gm1 = p.gamma - 1.0;
    // trace-interpolation fragments of this lane: B[k=j][n=e] = If[point 8 fgt + e][trace node 4 fkt + j]
    double bIf[D::FGT][D::FKT];
#pragma unroll
    for (int fgt = 0; fgt < D::FGT; ++fgt)
#pragma unroll
        for (int fkt = 0; fkt < D::FKT; ++fkt) bIf[fgt][fkt] = __ldg(p.tables + D::oIf + (fgt * D::FKT + fkt) * 32 + lane);

    const int64_t warpsPerGrid = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t warpId = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nFaceOct = (p.F + 7) >> 3;
    // The chain faceOwner -> connectivity -> traces is three dependent trips to memory; un-pipelined, a warp spent 70 % of its time
    // waiting on two of them (ncu, profiles/ncu_split_r02.md).  So: the owner index is fetched three iterations ahead, the
    // connectivity row two ahead, and the traces (and normal) of the NEXT iteration are in flight while this iteration computes.
    // Faces are walked from the last to the first: the element kernel that produced q_in walks the octets upwards, so the rows of the
    // highest elements are the ones still in the 126 MB L2 when this kernel starts - and the lowest faces / rows, touched last here,
    // are what the element kernel that follows asks for first.
    auto ownerOf = [&](int64_t it_) -> int { return it_ < nFaceOct ? __ldg(p.faceOwner + min((nFaceOct - 1 - it_) * 8 + e, p.F - 1)) : 0; };
    auto prefetchTraces = [&](int fo_, const int4& cn_) {
        const int64_t el_ = fo_ >> 2;
        const int face_ = fo_ & 3;
        const int nb_ = face_ == 0 ? cn_.x : (face_ == 1 ? cn_.y : cn_.z);
        const unsigned code_ = ((unsigned)cn_.w >> (8 * face_)) & 0xffu;
        const bool ghost_ = code_ & kCodeGhost;
        const int64_t nbBase_ = ghost_ ? p.ghostBase + (int64_t)nb_ * D::NfpPad : (int64_t)nb_ * D::NpPad;
        const int* ntp = nodeTab + ((code_ & kCodeFaceMask) * 2 + ((code_ & kCodeRev) ? 1 : 0)) * D::NfpPad;
        const int* nop = nodeTab + (face_ * 2) * D::NfpPad;
        // lane j covers field j: first and last trace node of both sides (a trace spans at most two 128-B lines per 16 nodes of a row)
        const double* qo = j == 0 ? p.qin[0] : (j == 1 ? p.qin[1] : (j == 2 ? p.qin[2] : p.qin[3]));
        const double* qg = j == 0 ? p.qghost[0] : (j == 1 ? p.qghost[1] : (j == 2 ? p.qghost[2] : p.qghost[3]));
        prefetchL1(qo + el_ * D::NpPad + nop[0]);
        prefetchL1(qo + el_ * D::NpPad + nop[D::Nfp - 1]);
        prefetchL1((ghost_ ? qg : qo) + nbBase_ + (ghost_ ? 0 : ntp[0]));
        prefetchL1((ghost_ ? qg : qo) + nbBase_ + (ghost_ ? D::Nfp - 1 : ntp[D::Nfp - 1]));
        if (j == 0) prefetchL1(p.geo + el_ * 16 + kGeoN);
    };
    // owner (M) and exterior (P) traces of face-octet entry (fo_, cn_) as A fragments over the face nodes, both in the owner's
    // traversal direction; also the face's connectivity code and the owner's normal
    auto loadTraces = [&](int fo_, const int4& cn_, double (&am_)[4][D::FKT], double (&an_)[4][D::FKT], unsigned& code_, double2& nxy_) {
        const int64_t el_ = fo_ >> 2;
        const int face_ = fo_ & 3;
        const int nb_ = face_ == 0 ? cn_.x : (face_ == 1 ? cn_.y : cn_.z);
        code_ = ((unsigned)cn_.w >> (8 * face_)) & 0xffu;
        const bool ghost_ = code_ & kCodeGhost;
        const int64_t eoff_ = el_ * D::NpPad;
        const int64_t nbBase_ = ghost_ ? p.ghostBase + (int64_t)nb_ * D::NfpPad : (int64_t)nb_ * D::NpPad;
        const int* nt_ = nodeTab + ((code_ & kCodeFaceMask) * 2 + ((code_ & kCodeRev) ? 1 : 0)) * D::NfpPad;
        const int* no_ = nodeTab + (face_ * 2) * D::NfpPad;
        nxy_ = __ldg(reinterpret_cast<const double2*>(p.geo + el_ * 16 + kGeoN) + face_);
#pragma unroll
        for (int fkt = 0; fkt < D::FKT; ++fkt) {
            const int i = fkt * 4 + j;
            const bool in = i < D::Nfp;
            const int64_t off = nbBase_ + (ghost_ ? i : nt_[in ? i : 0]);
            const int offO = no_[in ? i : 0];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                an_[f][fkt] = in ? __ldg((ghost_ ? p.qghost[f] : p.qin[f]) + off) : 0.0;
                am_[f][fkt] = in ? __ldg(p.qin[f] + eoff_ + offO) : 0.0;
            }
        }
    };
    int fo0 = ownerOf(warpId), fo1 = ownerOf(warpId + warpsPerGrid), fo2 = ownerOf(warpId + 2 * warpsPerGrid);
    int4 cn0 = __ldg(p.conn + (fo0 >> 2)), cn1 = __ldg(p.conn + (fo1 >> 2));
#ifndef HDG_FACE_L1PF
    // the traces of the next iteration are fetched into a second register set while this iteration computes (A/B on a B200: as fast
    // as or faster than pulling them into L1 with prefetches, which cost L1 tag look-ups of their own; HDG_FACE_L1PF selects that variant)
    double amN[4][D::FKT], anN[4][D::FKT];
    unsigned codeN = 0;
    double2 nxyN = make_double2(0.0, 0.0);
    if (warpId < nFaceOct) loadTraces(fo0, cn0, amN, anN, codeN, nxyN);
#endif
    for (int64_t it = warpId; it < nFaceOct; it += warpsPerGrid) {
        const int64_t fid = (nFaceOct - 1 - it) * 8 + e;
        const bool valid = fid < p.F;
        double am[4][D::FKT], an[4][D::FKT];
        unsigned code;
        double2 nxy;
#ifndef HDG_FACE_L1PF
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
            for (int fkt = 0; fkt < D::FKT; ++fkt) { am[f][fkt] = amN[f][fkt]; an[f][fkt] = anN[f][fkt]; }
        code = codeN;
        nxy = nxyN;
        if (it + warpsPerGrid < nFaceOct) loadTraces(fo1, cn1, amN, anN, codeN, nxyN);
#else
        loadTraces(fo0, cn0, am, an, code, nxy);
#ifndef HDG_FACE_NOPF
        if (it + warpsPerGrid < nFaceOct) prefetchTraces(fo1, cn1);      // next iteration's lines towards L1
#endif
#endif
        {   // the index pipeline one step on
            const int4 cn2 = __ldg(p.conn + (fo2 >> 2));
            const int fo3 = ownerOf(it + 3 * warpsPerGrid);
            fo0 = fo1; cn0 = cn1;
            fo1 = fo2; cn1 = cn2;
            fo2 = fo3;
        }
        double* fb = p.flux + fid * (4 * SL);
#pragma unroll
        for (int fgt = 0; fgt < D::FGT; ++fgt) {
            double cm[4][2], cp[4][2];
#pragma unroll
            for (int f = 0; f < 4; ++f) cm[f][0] = cm[f][1] = cp[f][0] = cp[f][1] = 0.0;
#pragma unroll
            for (int fkt = 0; fkt < D::FKT; ++fkt) {
#pragma unroll
                for (int f = 0; f < 4; ++f) dmma(cm[f], am[f][fkt], bIf[fgt][fkt]);
#pragma unroll
                for (int f = 0; f < 4; ++f) dmma(cp[f], an[f][fkt], bIf[fgt][fkt]);
            }
            double fl[2][4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double qM[4] = {cm[0][h], cm[1][h], cm[2][h], cm[3][h]};
                double qP[4] = {cp[0][h], cp[1][h], cp[2][h], cp[3][h]};
                if (code & kCodeReflect) {      // transform(I - 2nn, trace) on the momentum (reflectiveDgPatchField.C:140-147)
                    const double d2 = 2.0 * (qP[1] * nxy.x + qP[2] * nxy.y);
                    qP[1] -= d2 * nxy.x;
                    qP[2] -= d2 * nxy.y;
                }
                eulerFaceFluxPoint(FLUX, qM, qP, nxy.x, nxy.y, gm1, fl[h]);
            }
            const int s0 = fgt * 8 + 2 * j;
            if (valid && s0 < SL) {
#pragma unroll
                for (int f = 0; f < 4; ++f) *reinterpret_cast<double2*>(fb + f * SL + s0) = make_double2(fl[0][f], fl[1][f]);
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------------------------------------
// Face kernel for N = 1, 2: TWO faces per DMMA row.  With Nfg <= 4 Gauss points a face fills half of an 8-column tile, so the plain
// kernel would evaluate the Roe flux on 8 slots for 3-4 real points.  Here row e carries faces A = 2e and B = 2e + 1 of a 16-face
// group: k-tile 0 holds A's trace nodes (Nfp <= 3 < 4) against an operator that is non-zero in columns 0-3 only, k-tile 1 B's against
// columns 4-7, so lanes j = 0, 1 end up with A's four points and lanes j = 2, 3 with B's.  Same DMMA count per iteration, twice the faces.
// -------------------------------------------------------------------------------------------------------------------------------
template <int N, int FLUX>
__global__ void __launch_bounds__(HDG_FACE_THREADS(N), HDG_FACE_MINBLOCKS(N)) eulerFacePairFluxKernel(const StageParams p)
{
    using D = Dims<N>;
    constexpr int SL = D::fluxSlots;
    static_assert(D::Nfg <= 4 && D::Nfp <= 4 && D::FGT == 1 && D::FKT == 1 && SL == 4, "two faces per tile need Nfg <= 4");
    __shared__ int nodeTab[D::nodeTabInts];
    for (int i = threadIdx.x; i < D::nodeTabInts; i += blockDim.x) nodeTab[i] = p.nodeTab[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int e = lane >> 2, j = lane & 3;
    const double gm1 = p.gamma - 1.0;
    // B[k=j][n=e] = If[point e & 3][node j] in this face's half of the columns (the table entry of lane 4 (e & 3) + j), zero in the other half
    const double bIf = __ldg(p.tables + D::oIf + (e & 3) * 4 + j);
    const double bT[2] = {e < 4 ? bIf : 0.0, e >= 4 ? bIf : 0.0};
    const int mine = j >> 1;      // which of the row's two faces this lane's two points belong to

    const int64_t warpsPerGrid = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t warpId = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nGroup = (p.F + 15) >> 4;
    auto faceOf = [&](int64_t it_, int t) -> int64_t { return ((nGroup - 1 - it_) * 8 + e) * 2 + t; };      // walked downwards, see eulerFaceFluxKernel
    auto ownerOf = [&](int64_t it_, int t) -> int { return it_ < nGroup ? __ldg(p.faceOwner + min(faceOf(it_, t), p.F - 1)) : 0; };
    auto loadTraces = [&](int fo_, const int4& cn_, double (&am_)[4], double (&an_)[4], unsigned& code_, double2& nxy_) {
        const int64_t el_ = fo_ >> 2;
        const int face_ = fo_ & 3;
        const int nb_ = face_ == 0 ? cn_.x : (face_ == 1 ? cn_.y : cn_.z);
        code_ = ((unsigned)cn_.w >> (8 * face_)) & 0xffu;
        const bool ghost_ = code_ & kCodeGhost;
        const int64_t nbBase_ = ghost_ ? p.ghostBase + (int64_t)nb_ * D::NfpPad : (int64_t)nb_ * D::NpPad;
        const int* nt_ = nodeTab + ((code_ & kCodeFaceMask) * 2 + ((code_ & kCodeRev) ? 1 : 0)) * D::NfpPad;
        const int* no_ = nodeTab + (face_ * 2) * D::NfpPad;
        nxy_ = __ldg(reinterpret_cast<const double2*>(p.geo + el_ * 16 + kGeoN) + face_);
        const bool in = j < D::Nfp;
        const int64_t off = nbBase_ + (ghost_ ? j : nt_[in ? j : 0]);
        const int offO = no_[in ? j : 0];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            an_[f] = in ? __ldg((ghost_ ? p.qghost[f] : p.qin[f]) + off) : 0.0;
            am_[f] = in ? __ldg(p.qin[f] + el_ * D::NpPad + offO) : 0.0;
        }
    };
    int fo0[2], fo1[2], fo2[2];
    int4 cn0[2], cn1[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        fo0[t] = ownerOf(warpId, t);
        fo1[t] = ownerOf(warpId + warpsPerGrid, t);
        fo2[t] = ownerOf(warpId + 2 * warpsPerGrid, t);
        cn0[t] = __ldg(p.conn + (fo0[t] >> 2));
        cn1[t] = __ldg(p.conn + (fo1[t] >> 2));
    }
    double amN[2][4], anN[2][4];
    unsigned codeN[2] = {0, 0};
    double2 nxyN[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
    if (warpId < nGroup) {
#pragma unroll
        for (int t = 0; t < 2; ++t) loadTraces(fo0[t], cn0[t], amN[t], anN[t], codeN[t], nxyN[t]);
    }
    for (int64_t it = warpId; it < nGroup; it += warpsPerGrid) {
        double am[2][4], an[2][4];
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int f = 0; f < 4; ++f) { am[t][f] = amN[t][f]; an[t][f] = anN[t][f]; }
        const unsigned code = mine ? codeN[1] : codeN[0];
        const double2 nxy = mine ? nxyN[1] : nxyN[0];
        const int64_t fid = faceOf(it, mine);
        if (it + warpsPerGrid < nGroup) {
#pragma unroll
            for (int t = 0; t < 2; ++t) loadTraces(fo1[t], cn1[t], amN[t], anN[t], codeN[t], nxyN[t]);
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {      // the index pipeline one step on
            const int4 cn2 = __ldg(p.conn + (fo2[t] >> 2));
            const int fo3 = ownerOf(it + 3 * warpsPerGrid, t);
            fo0[t] = fo1[t]; cn0[t] = cn1[t];
            fo1[t] = fo2[t]; cn1[t] = cn2;
            fo2[t] = fo3;
        }
        double cm[4][2], cp[4][2];
#pragma unroll
        for (int f = 0; f < 4; ++f) cm[f][0] = cm[f][1] = cp[f][0] = cp[f][1] = 0.0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
#pragma unroll
            for (int f = 0; f < 4; ++f) dmma(cm[f], am[t][f], bT[t]);
#pragma unroll
            for (int f = 0; f < 4; ++f) dmma(cp[f], an[t][f], bT[t]);
        }
        double fl[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double qM[4] = {cm[0][h], cm[1][h], cm[2][h], cm[3][h]};
            double qP[4] = {cp[0][h], cp[1][h], cp[2][h], cp[3][h]};
            if (code & kCodeReflect) {      // transform(I - 2nn, trace) on the momentum (reflectiveDgPatchField.C:140-147)
                const double d2 = 2.0 * (qP[1] * nxy.x + qP[2] * nxy.y);
                qP[1] -= d2 * nxy.x;
                qP[2] -= d2 * nxy.y;
            }
            eulerFaceFluxPoint(FLUX, qM, qP, nxy.x, nxy.y, gm1, fl[h]);
        }
        if (fid < p.F) {
            double* fb = p.flux + fid * (4 * SL) + ((2 * j) & 3);
#pragma unroll
            for (int f = 0; f < 4; ++f) *reinterpret_cast<double2*>(fb + f * SL) = make_double2(fl[0][f], fl[1][f]);
        }
    }
}

// -------------------------------------------------------------------------------------------------------------------------------
// Element kernel: volume term, lift of the stored fluxes, explicit update
// -------------------------------------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(HDG_SPLIT_THREADS(N), HDG_SPLIT_MINBLOCKS(N)) eulerElemKernel(const StageParams p)
{
    using D = Dims<N>;
    constexpr int SL = D::fluxSlots;
#ifdef HDG_SPLIT_RHO_QUAD
    constexpr int kF0 = 0;      // A/B: density equation through the cubature like the others
#else
    constexpr int kF0 = 1;
#endif
    extern __shared__ __align__(128) double smem[];
    // N >= 9 (own cubature, beyond the reference's table): the fragments (410 / 600 KB) do not fit in shared memory and are read through
    // L1 from global memory; the nodal A fragments no longer fit in registers next to the accumulators and are parked in shared memory
    // ([f][kt][lane] per warp: each lane reads back its own slots) - as in the fused kernel
    const double* tab = D::big ? p.splitTables : smem;
    __shared__ unsigned long long tableBar;
    if constexpr (!D::big) stageTables(smem, p.splitTables, D::splitTableDoubles, &tableBar);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int e = lane >> 2, j = lane & 3;
    const int64_t warpsPerGrid = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t warpId = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const double gm1 = p.gamma - 1.0;
    double* aS = smem + (D::big ? (threadIdx.x >> 5) * (4 * D::KT * 32) + lane : 0);

    const int64_t n1 = p.octEnd - p.octBegin, nTot = p.octList ? p.nList : n1 + (p.octEnd2 - p.octBegin2);
    auto octOf = [&](int64_t i) -> int64_t { return p.octList ? (int64_t)__ldg(p.octList + i) : (i < n1 ? p.octBegin + i : p.octBegin2 + (i - n1)); };
    // A fragments of an element's nodal state: a[f][kt] = q_f[node 4*kt + j].  They are loop-carried: the fragments of the warp's NEXT
    // octet are fetched as soon as the last interpolation of this one has been issued (the registers are free from there on), so
    // the fetch is covered by the last projection, the lift and the update instead of stalling the start of the next octet
    double a[D::big ? 1 : 4][D::big ? 1 : D::KT];
    auto loadA = [&](int64_t it_) {
        const int64_t el_ = min(octOf(it_) * 8 + e, p.K - 1);
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
            for (int kt = 0; kt < D::KT; ++kt) {
                const double v = __ldg(p.qin[f] + el_ * D::NpPad + kt * 4 + j);
                if constexpr (D::big) aS[(f * D::KT + kt) * 32] = v;
                else a[f][kt] = v;
            }
    };
    auto aFrag = [&](int f, int kt) -> double {
        if constexpr (D::big) return aS[(f * D::KT + kt) * 32];
        else return a[f][kt];
    };
    // (N >= 7: the fragments alone are 72-96 registers; there the fetch stays at the top of the octet, behind an L1 prefetch)
    constexpr bool kEarly = N >= HDG_SPLIT_EARLY_MIN && N <= 6;
    if (kEarly && warpId < nTot) loadA(warpId);
    for (int64_t it = warpId; it < nTot; it += warpsPerGrid) {
        if constexpr (!kEarly) loadA(it);
        const int64_t oct = octOf(it);
        const int64_t elem = oct * 8 + e;
        const bool valid = elem < p.K;
        const int64_t el = valid ? elem : p.K - 1;
        const double* geo = p.geo + el * 16;
        const int64_t eoff = el * D::NpPad;

        double acc[4][D::NT][2];
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
            for (int nt = 0; nt < D::NT; ++nt) acc[f][nt][0] = acc[f][nt][1] = 0.0;

        // connectivity codes (owner bit, reversal) and the dgFace ids; pull the three flux records and q_aux towards L1 now
        const int4 cn = __ldg(p.conn + el);
        const int4 ef = __ldg(p.elemFace + el);
        {
            const int fidj = j == 0 ? ef.x : (j == 1 ? ef.y : ef.z);
            const char* rec = reinterpret_cast<const char*>(p.flux + (int64_t)fidj * (4 * SL));
            if (j < 3) {
#pragma unroll
                for (int b = 0; b < 4 * SL * 8; b += 128) prefetchL1(rec + b);
                if ((4 * SL * 8) % 128) prefetchL1(rec + 4 * SL * 8 - 8);
            }
#ifndef HDG_SPLIT_NO_UPD_PF
            if constexpr (N >= 7)
#endif
            if (p.mode == 0 && p.A != 0.0) prefetchL1((j == 0 ? p.qaux[0] : j == 1 ? p.qaux[1] : j == 2 ? p.qaux[2] : p.qaux[3]) + eoff);
        }

        // face fluxes over ONE K axis of the Gauss points of all three faces (slot s = face * Nfg + point, KTL k-tiles of 4 slots:
        // lane j supplies slot 4 kt + j).  The dgFace owner reads its points as stored; the neighbour traverses the face the other
        // way round (kCodeRev) and sees -F*.n: stored point Nfg-1-point, sign flipped.  Slots beyond 3 Nfg read a finite value;
        // their lift entries are 0.
        double fl[D::big ? 1 : D::KTL][4];
        auto loadFluxTile = [&](int kt, double (&flt)[4]) {
            {
                const int s = kt * 4 + j;
                const int face = s >= 3 * D::Nfg ? 0 : (s >= 2 * D::Nfg ? 2 : (s >= D::Nfg ? 1 : 0));
                const int pt = s >= 3 * D::Nfg ? 0 : s - face * D::Nfg;
                const int fid = face == 0 ? ef.x : (face == 1 ? ef.y : ef.z);
                const unsigned code = ((unsigned)cn.w >> (8 * face)) & 0xffu;
                const bool own = code & kCodeOwner;
                const bool rev = !own && (code & kCodeRev);
                const double fs = __ldg(geo + kGeoFs + face);
                const double sc = own ? fs : -fs;
                const double* fb = p.flux + (int64_t)fid * (4 * SL) + (rev ? D::Nfg - 1 - pt : pt);
#pragma unroll
                for (int f = 0; f < 4; ++f) flt[f] = sc * __ldg(fb + f * SL);
            }
        };
        auto loadFlux = [&]() {
            if constexpr (!D::big) {
#pragma unroll
                for (int kt = 0; kt < D::KTL; ++kt) loadFluxTile(kt, fl[kt]);
            }
        };

        // ---- volume term (as eulerStageKernel) ---------------------------------------------------------------------------------
        {
            const double2 g01 = __ldg(reinterpret_cast<const double2*>(geo));
            const double2 g23 = __ldg(reinterpret_cast<const double2*>(geo) + 1);
            const double rx = g01.x, ry = g01.y, sx = g23.x, sy = g23.y;
            auto interp = [&](int gt, double (&c)[4][2]) {
#pragma unroll
                for (int f = 0; f < 4; ++f) c[f][0] = c[f][1] = 0.0;
                const double* tv = tab + D::sVg + gt * D::KT * 32 + lane;
#pragma unroll
                for (int kt = 0; kt < D::KT; ++kt) {
                    const double b = tv[kt * 32];
#pragma unroll
                    for (int f = 0; f < 4; ++f) dmma(c[f], aFrag(f, kt), b);
                }
            };
            auto project = [&](int gt, const double (&Gr)[2][4], const double (&Gs)[2][4]) {
                const double* tr = tab + D::sPr + gt * 2 * D::NT * 32 + lane;
                const double* ts = tab + D::sPs + gt * 2 * D::NT * 32 + lane;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int nt = 0; nt < D::NT; ++nt) {
                        const double br = tr[(h * D::NT + nt) * 32];
#pragma unroll
                        for (int f = kF0; f < 4; ++f) dmma(acc[f][nt], Gr[h][f], br);
                    }
#pragma unroll
                    for (int nt = 0; nt < D::NT; ++nt) {
                        const double bs = ts[(h * D::NT + nt) * 32];
#pragma unroll
                        for (int f = kF0; f < 4; ++f) dmma(acc[f][nt], Gs[h][f], bs);
                    }
                }
            };
            if constexpr (kF0 == 1) {
                // density: its flux rhoU is linear in the nodal data, so the cubature sum Pr diag(..) Vg collapses exactly to the weak
                // nodal derivative (the 3(N+1) rule integrates the degree 2N-1 integrand exactly):  acc_rho += Dwr (rx q1 + ry q2) + Dws (sx q1 + sy q2)
#pragma unroll
                for (int kt = 0; kt < D::KT; ++kt) {
                    const double a1 = aFrag(1, kt), a2 = aFrag(2, kt);
                    const double ar = rx * a1 + ry * a2, as = sx * a1 + sy * a2;
                    // (all n-tiles of one operator before the other: two DMMAs into the same accumulator are never back to back)
#pragma unroll
                    for (int nt = 0; nt < D::NT; ++nt) dmma(acc[0][nt], ar, tab[D::sDwr + (kt * D::NT + nt) * 32 + lane]);
#pragma unroll
                    for (int nt = 0; nt < D::NT; ++nt) dmma(acc[0][nt], as, tab[D::sDws + (kt * D::NT + nt) * 32 + lane]);
                }
            }
            double c[4][2], Gr[2][4], Gs[2][4];
#ifdef HDG_SPLIT_NOPIPE
            constexpr bool kPipe = false;      // A/B: plain loop for every order
#else
            constexpr bool kPipe = !D::big;    // N >= 9: plain loop (interpolate, point-wise fluxes, project): fewer live registers
#endif
            if constexpr (!kPipe) {
#pragma unroll 1
            for (int gt = 0; gt + 1 < D::GT; ++gt) {
                interp(gt, c);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double q[4] = {c[0][h], c[1][h], c[2][h], c[3][h]};
                    eulerVolumeFlux(q, rx, ry, sx, sy, gm1, Gr[h], Gs[h]);
                }
                project(gt, Gr, Gs);
            }
            interp(D::GT - 1, c);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double q[4] = {c[0][h], c[1][h], c[2][h], c[3][h]};
                eulerVolumeFlux(q, rx, ry, sx, sy, gm1, Gr[h], Gs[h]);
            }
            } else {
            // software pipeline: the point-wise fluxes of tile gt+1 are emitted with the projection DMMAs of tile gt
            interp(0, c);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double q[4] = {c[0][h], c[1][h], c[2][h], c[3][h]};
                eulerVolumeFlux(q, rx, ry, sx, sy, gm1, Gr[h], Gs[h]);
            }
#ifdef HDG_SPLIT_VOL_UNROLL2
#pragma unroll 2
#else
#pragma unroll 1
#endif
            for (int gt = 0; gt + 1 < D::GT; ++gt) {
                double Gr2[2][4], Gs2[2][4];
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
            }
            }
            // the last interpolation has been issued: gather the face fluxes and the next octet's nodal fragments now, under the
            // projection DMMAs of the last tile
            if constexpr (kEarly) loadFlux();
#ifndef HDG_SPLIT_NO_UPD_PF
            if constexpr (kEarly) {   // what the update will read (q_in in its double2 layout, q_aux / the residual): towards L1 under the last projection and the lift
                const double* const* up = p.mode == 0 ? p.qaux : p.res;
                if (p.mode != 0 || p.A != 0.0) {
#pragma unroll
                    for (int b = 0; b < D::NpPad; b += 16) prefetchL1((j == 0 ? up[0] : j == 1 ? up[1] : j == 2 ? up[2] : up[3]) + eoff + b);
                }
#pragma unroll
                for (int b = 0; b < D::NpPad; b += 16) prefetchL1((j == 0 ? p.qin[0] : j == 1 ? p.qin[1] : j == 2 ? p.qin[2] : p.qin[3]) + eoff + b);
            }
#endif
            {
                const int64_t itn = it + warpsPerGrid;
                if (itn < nTot) {
                    const int64_t eln = min(octOf(itn) * 8 + e, p.K - 1);
                    if constexpr (kEarly) loadA(itn);
                    else prefetchL1((j == 0 ? p.qin[0] : j == 1 ? p.qin[1] : j == 2 ? p.qin[2] : p.qin[3]) + eln * D::NpPad);
                    if (j == 0) prefetchL1(p.geo + eln * 16);
                    if (j == 1) prefetchL1(p.conn + eln);
                    if (j == 2) prefetchL1(p.elemFace + eln);
                }
            }
            project(D::GT - 1, Gr, Gs);
            if constexpr (!kEarly) loadFlux();
        }

        // ---- surface term: lift of the stored face fluxes (gathered above) ----------------------------------------------------------
#pragma unroll
        for (int kt = 0; kt < D::KTL; ++kt) {
            if constexpr (D::big) loadFluxTile(kt, fl[0]);      // N >= 9: one k-tile at a time (the accumulators hold 112 / 144 registers)
#pragma unroll
            for (int nt = 0; nt < D::NT; ++nt) {
                const double b = tab[D::sLiftC + (kt * D::NT + nt) * 32 + lane];
#pragma unroll
                for (int f = 0; f < 4; ++f) dmma(acc[f][nt], fl[D::big ? 0 : kt][f], b);
            }
        }

        // ---- explicit update (mass solve folded into Pr/Ps/LIFT), as eulerStageKernel -----------------------------------------------
        if (valid && D::big) {
            // N >= 9: one node pair at a time (no per-field arrays next to 112 / 144 accumulator registers)
            const int64_t off0 = eoff + 2 * j;
            const bool useAux = p.mode == 0 && p.A != 0.0, two = p.mode == 0 && p.qout2[0] != nullptr;
#pragma unroll
            for (int f = 0; f < 4; ++f)
#pragma unroll
                for (int nt = 0; nt < D::NT; ++nt) {
                    const double2 qi = __ldg(reinterpret_cast<const double2*>(p.qin[f] + off0 + nt * 8));
                    if (p.mode == 0) {
                        const double2 qa = useAux ? __ldg(reinterpret_cast<const double2*>(p.qaux[f] + off0 + nt * 8)) : make_double2(0.0, 0.0);
                        if (two) {
                            const double2 q2 = p.A2 != 0.0 ? __ldg(reinterpret_cast<const double2*>(p.qaux2[f] + off0 + nt * 8)) : make_double2(0.0, 0.0);
                            *reinterpret_cast<double2*>(p.qout2[f] + off0 + nt * 8) =
                                make_double2(p.B2 * (qi.x + p.dt * acc[f][nt][0]) + p.A2 * q2.x, p.B2 * (qi.y + p.dt * acc[f][nt][1]) + p.A2 * q2.y);
                        }
                        *reinterpret_cast<double2*>(p.qout[f] + off0 + nt * 8) =
                            make_double2(p.B * (qi.x + p.dt * acc[f][nt][0]) + p.A * qa.x, p.B * (qi.y + p.dt * acc[f][nt][1]) + p.A * qa.y);
                    } else {
                        double2 r = *reinterpret_cast<const double2*>(p.res[f] + off0 + nt * 8);
                        r.x = p.A * r.x + p.dt * acc[f][nt][0];
                        r.y = p.A * r.y + p.dt * acc[f][nt][1];
                        *reinterpret_cast<double2*>(p.res[f] + off0 + nt * 8) = r;
                        *reinterpret_cast<double2*>(p.qout[f] + off0 + nt * 8) = make_double2(qi.x + p.B * r.x, qi.y + p.B * r.y);
                    }
                }
        } else if (valid) {
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

// -------------------------------------------------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------------------------------------------------
namespace {
// N = 1, 2 (Nfg <= 4): two faces per DMMA row (eulerFacePairFluxKernel); HDG_FACE_NO_PAIRS keeps the plain face kernel (A/B)
#ifdef HDG_FACE_NO_PAIRS
template <int N> constexpr bool kFacePairs = false;
#else
template <int N> constexpr bool kFacePairs = N <= 2;
#endif
struct SplitCfg { int elemBlocks = 0, faceBlocks = 0; size_t smem = 0; bool ok = false; };

template <int N>
SplitCfg& splitCfgT()
{
    static SplitCfg cfg[64];
    int dev = 0;
    cudaGetDevice(&dev);
    SplitCfg& c = cfg[dev & 63];
    if (!c.ok) {
        using D = Dims<N>;
        c.smem = sizeof(double) * (D::big ? (HDG_SPLIT_THREADS(N) / 32) * 4 * D::KT * 32 : D::splitTableDoubles);
        cudaError_t err = cudaFuncSetAttribute(eulerElemKernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
        if (err != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute(eulerElemKernel): ") + cudaGetErrorString(err));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.elemBlocks, eulerElemKernel<N>, HDG_SPLIT_THREADS(N), c.smem);
        if constexpr (kFacePairs<N>) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.faceBlocks, eulerFacePairFluxKernel<N, 0>, HDG_FACE_THREADS(N), 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.faceBlocks, eulerFaceFluxKernel<N, 0>, HDG_FACE_THREADS(N), 0);
        if (c.elemBlocks < 1 || c.faceBlocks < 1) throw std::runtime_error("split Euler stage: kernel does not fit on this device");
        c.ok = true;
    }
    return c;
}

template <int N>
void launchSplitT(const StageParams& p, bool faces, int smCount, cudaStream_t st)
{
    SplitCfg& c = splitCfgT<N>();
    if (faces) {
        const int64_t nFaceOct = kFacePairs<N> ? (p.F + 15) >> 4 : (p.F + 7) >> 3;
        const int grid = (int)std::min<int64_t>((int64_t)smCount * c.faceBlocks, (nFaceOct + HDG_FACE_THREADS(N) / 32 - 1) / (HDG_FACE_THREADS(N) / 32));
        if (grid > 0) {
            if constexpr (kFacePairs<N>) {
                if (p.fluxKind == 1) eulerFacePairFluxKernel<N, 1><<<grid, HDG_FACE_THREADS(N), 0, st>>>(p);
                else eulerFacePairFluxKernel<N, 0><<<grid, HDG_FACE_THREADS(N), 0, st>>>(p);
            } else {
                if (p.fluxKind == 1) eulerFaceFluxKernel<N, 1><<<grid, HDG_FACE_THREADS(N), 0, st>>>(p);
                else eulerFaceFluxKernel<N, 0><<<grid, HDG_FACE_THREADS(N), 0, st>>>(p);
            }
        }
    }
    const int64_t nOct = p.octList ? p.nList : (p.octEnd - p.octBegin) + (p.octEnd2 - p.octBegin2);
    const int wpb = HDG_SPLIT_THREADS(N) / 32;
    const int grid = (int)std::min<int64_t>((int64_t)smCount * c.elemBlocks, (nOct + wpb - 1) / wpb);
    if (grid > 0) eulerElemKernel<N><<<grid, HDG_SPLIT_THREADS(N), c.smem, st>>>(p);
}
}  // namespace

bool eulerSplitAvailable(int N) { return N >= 1 && N <= 10; }

// faces: also launch the face kernel (all dgFaces) before the element kernel
void launchEulerSplit(int N, const StageParams& p, bool faces, int smCount, cudaStream_t st)
{
    switch (N) {
        case 1: launchSplitT<1>(p, faces, smCount, st); break;
        case 2: launchSplitT<2>(p, faces, smCount, st); break;
        case 3: launchSplitT<3>(p, faces, smCount, st); break;
        case 4: launchSplitT<4>(p, faces, smCount, st); break;
        case 5: launchSplitT<5>(p, faces, smCount, st); break;
        case 6: launchSplitT<6>(p, faces, smCount, st); break;
        case 7: launchSplitT<7>(p, faces, smCount, st); break;
        case 8: launchSplitT<8>(p, faces, smCount, st); break;
        case 9: launchSplitT<9>(p, faces, smCount, st); break;
        case 10: launchSplitT<10>(p, faces, smCount, st); break;
        default: throw std::runtime_error("split Euler stage: orders 1..10");
    }
}

}  // namespace hdg
