// TMA-pipelined fused scalar-advection stage (sm_100a, FP64).  Same arithmetic as advectStageKernel (dg_kernels.cu) - the nodal
// collapse of defaultConvectionScheme.C:216-303 + LFFlux.C:105-211 - with a different data path:
//
//   * every contiguous stream of an octet (T_in, T_aux | residual, geometry: 1 KB each at NpPad = 16; velocity pairs: 2 KB) is fetched by a
//     TMA tensor copy (cp.async.bulk.tensor.2d, SASS UTMALDG) and its connectivity rows by a bulk copy (UBLKCP) into a per-warp
//     ring of S stages, completion on one mbarrier per stage; the 128B swizzle of the tensor maps spreads the per-element-row
//     fragment reads (lane = 4*row + j) over the banks, and DMMA row g carries element 4*(g&1) + (g>>1) so that the two rows of a
//     quarter warp never collide;
//   * the velocity is read from a derived copy that holds (x,y) pairs (hopedg.cu keeps it current): one 16-B gather per trace
//     slot instead of two 8-B ones;
//   * the result leaves with 16-B stores from the accumulator registers (DS) - the alternative, a swizzled shared-memory tile
//     and a TMA tensor store (UTMASTG), is kept behind HDG_ADV_CFG=2 and measures 2 % slower;
//   * only the neighbour traces are gathered with ordinary loads (L2 hits), issued as soon as the stage has landed and consumed
//     after the volume term; the three faces share one K axis (slot = face*Nfp + i): 4 k-tiles instead of 6 at N=4;
//   * all operator fragments live in registers (16 + 2*KTC double2 per lane), face geometry and connectivity are read with
//     per-element broadcast 16-B loads and selected per slot in registers.
//
// Each warp owns its ring and its barriers: there is no block-level synchronisation after the prologue.  ONE block of 12 warps per SM:
// the warps of a block walk consecutive octets, so their neighbour-trace gathers share L1 lines (3 blocks of 4 warps, 148 block ids
// apart, measure 4 % slower at N=4 and 25 % slower at N=5).
// advectStageTmaKernel is built for orders whose padded element row is one 128-B line (NpPad = 16: N = 3, 4);
// advectStageTmaWideKernel (second half of this file) carries the same pipeline to rows of NT x 64 B (N = 1, 2, 5, 6, 7); N >= 8 use
// advectStageKernel.  What was measured on the way (profiles/experiments_r01.md, experiments_r02.md) is the reason for every choice.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "dg_kernels.cuh"
#include "dg_advect_tiles.hpp"

namespace hdg {

namespace {

__device__ __forceinline__ void dmmaT(double (&d)[2], double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbarExpectTx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned bar, unsigned parity)
{
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
    }
}
__device__ __forceinline__ void tmaLoadRows(unsigned dst, const CUtensorMap* tm, int row, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(reinterpret_cast<unsigned long long>(tm)), "r"(0), "r"(row), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmaStoreRows(const CUtensorMap* tm, int row, unsigned src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<unsigned long long>(tm)),
                 "r"(0), "r"(row), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void tmaCommit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int Pending>
__device__ __forceinline__ void tmaWaitRead() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(Pending) : "memory"); }
__device__ __forceinline__ void tmaWaitAll() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// max of two non-NaN doubles (fmax() costs ~7 instructions for its NaN rules); ldgD keeps the gathers where they are written
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double ldgD(const double* p)
{
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ldgD2(const double* p)
{
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
// one lane of the (converged) warp, chosen by the hardware: the compiler knows that exactly one thread runs the guarded block and
// issues the uniform-datapath instructions (UTMALDG, UBLKCP, SYNCS) directly instead of wrapping each one in an election loop
__device__ __forceinline__ bool electOne()
{
    unsigned pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void pin(int& x) { asm volatile("" : "+r"(x)); }      // keep a loop-invariant in its register (no rematerialisation)

constexpr int kStageTiles = 5;     // T_in (1 KB), velocity pairs (2 KB), T_aux | residual, geometry
constexpr int kConnBytes = 256;    // per stage: connectivity of the octet for T and for U (8 x int4 each)
constexpr int kWarps = 4;

__device__ __forceinline__ void bulkLoad(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

template <int N, int S, bool DS, int NW = kWarps>
struct TmaLayout {
    using D = Dims<N>;
    static constexpr int warpBytes = (S * kStageTiles + (DS ? 0 : 1)) * kTile; // ring (+ one result tile when the result leaves by TMA)
    static constexpr int oConn = NW * warpBytes;                            // [warp][stage][256 B]
    static constexpr int oTab = oConn + NW * S * kConnBytes;                // operator fragments, nt pairs as double2
    static constexpr int tabBytes = (8 + D::KTC) * 32 * 16;
    static constexpr int oNode = oTab + tabBytes;                               // faceToCellIndex as [rev*4 + face][NfpPad]
    static constexpr int oBars = oNode + 8 * D::NfpPad * 4;
    static constexpr int total = oBars + NW * S * 8;
};

}  // namespace

// DS = direct stores: the updated values leave with ordinary 16-B stores from the accumulator registers instead of a TMA store
// from shared memory: no proxy fence (MEMBAR.ALL.CTA), no result tile, and the stage is refilled before the stores are issued.
template <int N, int S, int MB, bool DS, int NW = kWarps>
__global__ void __launch_bounds__(32 * NW, MB)
    advectStageTmaKernel(const AdvectParams p, const __grid_constant__ CUtensorMap tmTin, const __grid_constant__ CUtensorMap tmUZ,
                         const __grid_constant__ CUtensorMap tmAux,
                         const __grid_constant__ CUtensorMap tmGeo, const __grid_constant__ CUtensorMap tmTout,
                         const __grid_constant__ CUtensorMap tmRes)
{
    using D = Dims<N>;
    using L = TmaLayout<N, S, DS, NW>;
    static_assert(D::NT == 2, "one 128-B line per element row");
    static_assert(D::Nfp >= 4, "a k-tile of 4 trace slots spans at most two faces");
    constexpr int KTC = D::KTC;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    unsigned char* base = smemRaw + ((1024u - (smemAddr(smemRaw) & 1023u)) & 1023u);      // swizzle atoms are 1 KB aligned
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);                  // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    // 16-B shared-memory reads are served a quarter warp (two DMMA rows g = lane >> 2) at a time; with e = g, rows 2q and 2q+1 would
    // map the four chunks 4nt+j to the same four swizzled positions (2-way bank conflict).  Row g carries element 4*(g & 1) + (g >> 1):
    // the two rows then differ in bit 2 of the swizzle XOR
    const int e = elemOfRow128(lane >> 2), j = lane & 3;
    unsigned char* ring = base + warp * L::warpBytes;
    unsigned char* outT = ring + S * kStageTiles * kTile;
    unsigned char* connS = base + L::oConn + warp * S * kConnBytes;
    const double2* tabS = reinterpret_cast<const double2*>(base + L::oTab) + lane;
    const unsigned char* nodeK = base + L::oNode;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + L::oBars) + warp * S;

    // ---- prologue: barriers, operator fragments (nt = 0,1 of one (k-tile, operator) side by side), face-node table ------------
    for (int i = threadIdx.x; i < (8 + KTC) * 32; i += blockDim.x) {
        const int t = i >> 5, ln = i & 31;      // t: 0..3 Dwr k-tiles, 4..7 Dws k-tiles, 8.. combined lift k-tiles
        const int off = t < 4 ? D::oDwr + t * 64 : (t < 8 ? D::oDws + (t - 4) * 64 : D::oLiftC + (t - 8) * 64);
        reinterpret_cast<double2*>(base + L::oTab)[i] = make_double2(__ldg(p.tables + off + ln), __ldg(p.tables + off + 32 + ln));
    }
    for (int i = threadIdx.x; i < 8 * D::NfpPad; i += blockDim.x) {
        const int row = i / D::NfpPad, c = i % D::NfpPad, face = row & 3, rev = row >> 2;      // row = code & 7
        reinterpret_cast<int*>(base + L::oNode)[i] = face < 3 ? p.nodeTab[(face * 2 + rev) * D::NfpPad + c] : 0;
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbarInit(smemAddr(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fenceProxyAsync();
    }
    __syncthreads();

    const int64_t nOct = (p.K + 7) >> 3;
    const int64_t W = (int64_t)gridDim.x * NW, w0 = (int64_t)blockIdx.x * NW + warp;
    const bool useAux = p.mode == 1 || p.A != 0.0;
    const bool sameConn = p.sameConn != 0;

    auto issueLoads = [&](int64_t oct, int s) {      // lane 0 only
        const unsigned bar = smemAddr(bars + s), dst = smemAddr(ring + s * kStageTiles * kTile), cdst = smemAddr(connS + s * kConnBytes);
        const int row = (int)(oct * 8);
        mbarExpectTx(bar, (useAux ? 5u : 4u) * kTile + (sameConn ? 128u : 256u));
        bulkLoad(cdst, p.connT + oct * 8, 128u, bar);
        if (!sameConn) bulkLoad(cdst + 128u, p.connU + oct * 8, 128u, bar);
        tmaLoadRows(dst, &tmTin, row, bar);
        tmaLoadRows(dst + kTile, &tmUZ, 2 * row, bar);      // 16 rows of 128 B: the (x,y) pairs of the 8 elements
        if (useAux) tmaLoadRows(dst + 3 * kTile, &tmAux, row, bar);
        tmaLoadRows(dst + 4 * kTile, &tmGeo, row, bar);
    };
    if (electOne()) {
#pragma unroll
        for (int s = 0; s < S; ++s)
            if (w0 + s * W < nOct) issueLoads(w0 + s * W, s);
    }

    // per-lane constants of the trace slots: slot = 4*kt + j = face*Nfp + i.  A k-tile spans faces fLo(kt) <= fHi(kt) (compile
    // time); `hi` tells whether this lane's slot belongs to the upper one.
    int slotI[KTC], slotI4[KTC], offOwn[KTC], offOwnU[KTC];
    bool hi[KTC];
#pragma unroll
    for (int kt = 0; kt < KTC; ++kt) {
        const int slot = 4 * kt + j;
        const bool slotValid = slot < 3 * D::Nfp;
        const int f = slotValid ? (slot >= D::Nfp) + (slot >= 2 * D::Nfp) : 2;
        hi[kt] = f != (4 * kt) / D::Nfp;
        slotI[kt] = slotValid ? slot - f * D::Nfp : 0;
        const int ownNode = reinterpret_cast<const int*>(nodeK)[f * D::NfpPad + slotI[kt]];
        offOwn[kt] = swz(e, ownNode);
        offOwnU[kt] = kTile + swzU(e, ownNode);
        slotI4[kt] = slotI[kt] * 4;
        pin(slotI[kt]); pin(slotI4[kt]); pin(offOwn[kt]); pin(offOwnU[kt]);
    }
    int offQ[2];      // the lane's node pair (8nt + 2j, +1) of its element row
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) offQ[nt] = swz(e, 8 * nt + 2 * j);
    // velocity pairs of the lane's four volume nodes 8nt + 2j + h: the two rows of a quarter warp share the swizzle XOR of the
    // velocity tile, so odd rows fetch h = 1 first (disjoint banks) and swap afterwards
    const int oddU = (lane >> 2) & 1;
    int offQU[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) offQU[nt][hh] = kTile + swzU(e, 8 * nt + 2 * j + (hh ^ oddU));
    // the operator fragments stay in registers (48 at N=4) instead of being re-read from shared memory for every octet: the 12
    // fragment loads were 48 of ~330 L1 data-pipe wavefronts per octet (ncu: that pipe was 89 % busy with them, 70 % without)
    double2 tabR[8 + KTC];
#pragma unroll
    for (int t = 0; t < 8 + KTC; ++t) tabR[t] = tabS[t * 32];
    auto frag = [&](int t) -> double2 { return tabR[t]; };
    const int ghostBase = (int)p.ghostBase;

    // element offset (in doubles) of the exterior trace value of slot kt: ghost slot or the neighbour's (rotated) face node
    // cn = the element's connectivity row (one broadcast 16-B read per element); the slot's face is fLo(kt) or fHi(kt)
    auto faceCode = [&](const int4& cn, int kt) -> unsigned {
        const int fLo = (4 * kt) / D::Nfp, fHi = (4 * kt + 3) / D::Nfp > 2 ? 2 : (4 * kt + 3) / D::Nfp;
        const unsigned w = (unsigned)cn.w;
        return (fLo == fHi ? w >> (8 * fLo) : (hi[kt] ? w >> (8 * fHi) : w >> (8 * fLo))) & 0xffu;
    };
    auto traceOffset = [&](const int4& cn, int kt) -> int {
        const int fLo = (4 * kt) / D::Nfp, fHi = (4 * kt + 3) / D::Nfp > 2 ? 2 : (4 * kt + 3) / D::Nfp;
        const int nbLo = fLo == 0 ? cn.x : (fLo == 1 ? cn.y : cn.z), nbHi = fHi == 0 ? cn.x : (fHi == 1 ? cn.y : cn.z);
        const int nb = fLo == fHi ? nbLo : (hi[kt] ? nbHi : nbLo);
        const unsigned code = faceCode(cn, kt);
        const int node = *reinterpret_cast<const int*>(nodeK + (code & 7u) * (D::NfpPad * 4) + slotI4[kt]);
        const bool gh = code & kCodeGhost;
        return nb * (gh ? D::NfpPad : D::NpPad) + (gh ? ghostBase + slotI[kt] : node);
    };

    double TN[KTC], uxN[KTC], uyN[KTC];
    unsigned codesU = 0;
    auto gather = [&](const unsigned char* cs) {
        const int4 cT = *reinterpret_cast<const int4*>(cs + e * 16);
        int4 cU = cT;
        if (!sameConn) cU = *reinterpret_cast<const int4*>(cs + 128 + e * 16);
        codesU = (unsigned)cU.w;
#pragma unroll
        for (int kt = 0; kt < KTC; ++kt) {
            const int oT = traceOffset(cT, kt);
            const int oU = sameConn ? oT : traceOffset(cU, kt);
            TN[kt] = ldgD(p.Tin + oT);
            const double2 u = ldgD2(p.UZ + 2 * (int64_t)oU);
            uxN[kt] = u.x;
            uyN[kt] = u.y;
        }
    };

    int it = 0;
    for (int64_t oct = w0; oct < nOct; oct += W, ++it) {
        const int s = it % S;
        const unsigned parity = (unsigned)(it / S) & 1u;
        const unsigned char* st = ring + s * kStageTiles * kTile;
        const unsigned char* cs = connS + s * kConnBytes;
        const bool valid = oct * 8 + e < p.K;

        mbarWait(smemAddr(bars + s), parity);

        // ---- neighbour-trace gathers (the only non-TMA loads; L1/L2 hits), consumed after the volume term.  (Issuing them one
        // iteration ahead was measured slower: 0.166 -> 0.19 ms; the neighbours' rows are then not yet in L2.)
        gather(cs);

        // ---- volume: rhs += Dwr (rx Ux T + ry Uy T) + Dws (sx Ux T + sy Uy T)   (defaultConvectionScheme.C:247-262) ------------
        // k-tile (2*nt'+h), slot j <-> node 8*nt' + 2*j + h: the double2 at chunk 4*nt'+j of the element row
        double2 Tq[2];
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        {
            const double2 g01 = *reinterpret_cast<const double2*>(st + 4 * kTile + swz(e, 0));
            const double2 g23 = *reinterpret_cast<const double2*>(st + 4 * kTile + swz(e, 2));
            double2 u[2][2];      // [nt][h] = (Ux, Uy) at node 8nt + 2j + h
#pragma unroll
            for (int ntp = 0; ntp < 2; ++ntp) {
                Tq[ntp] = *reinterpret_cast<const double2*>(st + offQ[ntp]);
                const double2 a0 = *reinterpret_cast<const double2*>(st + offQU[ntp][0]);
                const double2 a1 = *reinterpret_cast<const double2*>(st + offQU[ntp][1]);
                u[ntp][0].x = oddU ? a1.x : a0.x; u[ntp][0].y = oddU ? a1.y : a0.y;
                u[ntp][1].x = oddU ? a0.x : a1.x; u[ntp][1].y = oddU ? a0.y : a1.y;
            }
#pragma unroll
            for (int ntp = 0; ntp < 2; ++ntp) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int kt = 2 * ntp + h;
                    const double T = h ? Tq[ntp].y : Tq[ntp].x;
                    const double fx = u[ntp][h].x * T, fy = u[ntp][h].y * T;
                    const double ar = g01.x * fx + g01.y * fy, as = g23.x * fx + g23.y * fy;
                    const double2 br = frag(kt), bs = frag(4 + kt);
                    dmmaT(acc[0], ar, br.x);
                    dmmaT(acc[1], ar, br.y);
                    dmmaT(acc[0], as, bs.x);
                    dmmaT(acc[1], as, bs.y);
                }
            }
        }

        // ---- surface: nodal LF / average flux over the 3*Nfp trace slots, lifted with the combined LIFTn (LFFlux.C:147-206) ----
        double vO[KTC], vN[KTC], TO[KTC], fsK[KTC];
        double pm[3] = {0.0, 0.0, 0.0};
        // face geometry of the element: (nx,ny) of the three faces and the three Fscale, 5 broadcast 16-B reads (1 wavefront each)
        double2 nF[3];
#pragma unroll
        for (int f = 0; f < 3; ++f) nF[f] = *reinterpret_cast<const double2*>(st + 4 * kTile + swz(e, kGeoN + 2 * f));
        const double2 fs01 = *reinterpret_cast<const double2*>(st + 4 * kTile + swz(e, kGeoFs));
        const double2 fs2J = *reinterpret_cast<const double2*>(st + 4 * kTile + swz(e, kGeoFs + 2));
        const double fsF[3] = {fs01.x, fs01.y, fs2J.x};
#pragma unroll
        for (int kt = 0; kt < KTC; ++kt) {
            const int fLo = (4 * kt) / D::Nfp, fHi = (4 * kt + 3) / D::Nfp > 2 ? 2 : (4 * kt + 3) / D::Nfp;
            TO[kt] = *reinterpret_cast<const double*>(st + offOwn[kt]);
            const double2 uo = *reinterpret_cast<const double2*>(st + offOwnU[kt]);
            const double uxo = uo.x, uyo = uo.y;
            double2 nxy = nF[fLo];
            fsK[kt] = fsF[fLo];
            if (fLo != fHi) {
                nxy.x = hi[kt] ? nF[fHi].x : nxy.x;
                nxy.y = hi[kt] ? nF[fHi].y : nxy.y;
                fsK[kt] = hi[kt] ? fsF[fHi] : fsK[kt];
            }
            double uxn = uxN[kt], uyn = uyN[kt];
            if (p.anyReflect) {      // reflective U patch somewhere in the mesh (uniform branch): mirror the exterior velocity
                int4 cw;
                cw.w = (int)codesU;
                if (faceCode(cw, kt) & kCodeReflect) {
                    const double d2 = 2.0 * (uxn * nxy.x + uyn * nxy.y);
                    uxn -= d2 * nxy.x;
                    uyn -= d2 * nxy.y;
                }
            }
            vO[kt] = nxy.x * uxo + nxy.y * uyo;
            vN[kt] = nxy.x * uxn + nxy.y * uyn;
            // (a padding slot, 4kt+j >= 3Nfp, aliases node 0 of face 2: its values are finite members of that face's set, so it
            // changes neither the face maximum nor - its lift fragments being zero - the result)
            const double m = dmax(fabs(vO[kt]), fabs(vN[kt]));
            if (fLo == fHi) pm[fLo] = dmax(pm[fLo], m);
            else {
                pm[fLo] = (!hi[kt] && m > pm[fLo]) ? m : pm[fLo];
                pm[fHi] = (hi[kt] && m > pm[fHi]) ? m : pm[fHi];
            }
        }
        // one maxV per face (LFFlux.C:189-196): the slots of a face are spread over the 4 lanes of the element
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            pm[f] = dmax(pm[f], __shfl_xor_sync(0xffffffffu, pm[f], 1));
            pm[f] = dmax(pm[f], __shfl_xor_sync(0xffffffffu, pm[f], 2));
        }
        const double dissOn = p.fluxKind == 1 ? 0.5 : 0.0, fluxOn = p.fluxKind != 3 ? 1.0 : 0.0;
#pragma unroll
        for (int kt = 0; kt < KTC; ++kt) {
            const int fLo = (4 * kt) / D::Nfp, fHi = (4 * kt + 3) / D::Nfp > 2 ? 2 : (4 * kt + 3) / D::Nfp;
            const double maxV = fLo == fHi ? pm[fLo] : (hi[kt] ? pm[fHi] : pm[fLo]);
            double fl = (vO[kt] * TO[kt] + vN[kt] * TN[kt]) * 0.5 + (dissOn * maxV) * (TO[kt] - TN[kt]);
            fl *= fsK[kt] * fluxOn;
            const double2 bl = frag(8 + kt);
            dmmaT(acc[0], fl, bl.x);
            dmmaT(acc[1], fl, bl.y);
        }

        // ---- explicit update ------------------------------------------------------------------------------------------------
        double2 qx[2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) qx[nt] = useAux ? *reinterpret_cast<const double2*>(st + 3 * kTile + offQ[nt]) : make_double2(0.0, 0.0);
        double2 o[2], r[2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            if (p.mode == 0) {
                o[nt].x = p.B * (Tq[nt].x + p.dt * acc[nt][0]) + p.A * qx[nt].x;
                o[nt].y = p.B * (Tq[nt].y + p.dt * acc[nt][1]) + p.A * qx[nt].y;
            } else {
                r[nt].x = p.A * qx[nt].x + p.dt * acc[nt][0];
                r[nt].y = p.A * qx[nt].y + p.dt * acc[nt][1];
                o[nt].x = Tq[nt].x + p.B * r[nt].x;
                o[nt].y = Tq[nt].y + p.B * r[nt].y;
            }
        }
        if (oct * 8 + 8 > p.K) {      // warp-uniform: the last, ragged octet - its padding rows stay zero
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
                if (!valid) o[nt] = r[nt] = make_double2(0.0, 0.0);
        }
        if constexpr (DS) {
            __syncwarp();      // every lane has read what it needs from the stage: refill it
            if (electOne()) {
                const int64_t octr = oct + (int64_t)S * W;
                if (octr < nOct) issueLoads(octr, s);
            }
            const int64_t g0 = (oct * 8 + e) * D::NpPad + 2 * j;      // node pair (8nt+2j, +1) of element row e
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                *reinterpret_cast<double2*>(p.Tout + g0 + 8 * nt) = o[nt];
                if (p.mode == 1) *reinterpret_cast<double2*>(p.res + g0 + 8 * nt) = r[nt];
            }
        } else {
            // result tile in the swizzled layout, TMA store; mode 1 returns the residual through the stage's own T_in tile
            if (lane == 0) tmaWaitRead<0>();      // the previous store has finished reading the result tile (it had a whole iteration)
            __syncwarp();
            unsigned char* resT = const_cast<unsigned char*>(st);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                if (p.mode == 1) *reinterpret_cast<double2*>(resT + offQ[nt]) = r[nt];
                *reinterpret_cast<double2*>(outT + offQ[nt]) = o[nt];
            }
            fenceProxyAsync();
            __syncwarp();
            if (lane == 0) {
                const int row = (int)(oct * 8);
                tmaStoreRows(&tmTout, row, smemAddr(outT));
                if (p.mode == 1) tmaStoreRows(&tmRes, row, smemAddr(resT));
                tmaCommit();
                const int64_t octr = oct + (int64_t)S * W;      // refill the stage just consumed
                if (octr < nOct) {
                    if (p.mode == 1) tmaWaitRead<0>();           // the residual store still reads this stage
                    issueLoads(octr, s);
                }
            }
        }
    }
    if (lane == 0) tmaWaitAll();
}

// ---------------------------------------------------------------------------------------------------------
// Other row widths (NT = NpPad / 8 = 1, 3, 4, 5: N = 1, 2 | 5 | 6 | 7): the same per-warp TMA pipeline for element rows of NT * 64 B.
//
//   * NT odd (64-, 192-, 320-B rows): tensor maps without swizzle, box = 8 element rows; consecutive element rows are an odd multiple
//     of four 16-B chunks apart, so the two DMMA rows of a quarter warp (elements 2q, 2q+1) read disjoint bank halves as they are;
//   * NT even (256-B rows): the plane is viewed as rows of 128 B (NT/2 per element) under the 128B swizzle; the two DMMA rows of a
//     quarter warp carry elements e and e ^ 3, which differ in bit 2 of the swizzle XOR (WideTile::elemOfRow);
//   * velocity pairs: NT swizzled 128-B rows per element for every NT (see WideTile::offU);
//   * operator fragments ((4 NT + KTC) NT doubles per lane: 51 at N=5, 88 at N=6) do not fit in registers next to the octet's
//     data: they are read from shared memory as they are needed (8-B loads, 2 wavefronts each).  The bytes per octet grow
//     faster than these reads (N=5: ~300 L1 wavefronts per 7 KB octet against ~280 per 4.6 KB at N=4).
// ---------------------------------------------------------------------------------------------------------
namespace {

template <int N, int S, int NW>
struct WideLayout {
    using D = Dims<N>;
    using G = WideTile<D::NT>;
    static constexpr int nFrag = (4 * D::NT + D::KTC) * D::NT;                 // Dwr [2NT][NT], Dws [2NT][NT], combined lift [KTC][NT]
    static constexpr int warpBytes = S * G::stageBytes;
    static constexpr int oConn = NW * warpBytes;                               // [warp][stage][256 B]
    static constexpr int oTab = oConn + NW * S * kConnBytes;                   // [nFrag][32] doubles
    static constexpr int oNode = oTab + nFrag * 32 * 8;                        // faceToCellIndex as [rev*4 + face][NfpPad]
    static constexpr int oBars = oNode + 8 * D::NfpPad * 4;
    static constexpr int total = oBars + NW * S * 8;
};

}  // namespace

// FRL: the combined-lift fragments (KTC * NT doubles per lane) are kept in registers
template <int N, int S, int NW, int MB, bool FRL>
__global__ void __launch_bounds__(32 * NW, MB)
    advectStageTmaWideKernel(const AdvectParams p, const __grid_constant__ CUtensorMap tmTin, const __grid_constant__ CUtensorMap tmUZ,
                             const __grid_constant__ CUtensorMap tmAux, const __grid_constant__ CUtensorMap tmGeo)
{
    using D = Dims<N>;
    using L = WideLayout<N, S, NW>;
    using G = WideTile<D::NT>;
    constexpr int NT = D::NT, KTC = D::KTC;
    static_assert(NT != 2, "128-B rows are served by advectStageTmaKernel");
    static_assert(D::Nfp >= 2, "a k-tile of 4 trace slots spans at most two faces");
    extern __shared__ __align__(16) unsigned char smemRaw[];
    unsigned char* base = smemRaw + ((1024u - (smemAddr(smemRaw) & 1023u)) & 1023u);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int e = G::elemOfRow(lane >> 2), j = lane & 3;
    unsigned char* ring = base + warp * L::warpBytes;
    unsigned char* connS = base + L::oConn + warp * S * kConnBytes;
    const double* tabS = reinterpret_cast<const double*>(base + L::oTab) + lane;
    const unsigned char* nodeK = base + L::oNode;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + L::oBars) + warp * S;

    // ---- prologue: barriers, operator fragments [Dwr | Dws | combined lift], face-node table -----------------------------------
    constexpr int nV = 2 * NT * NT * 32;      // doubles of one volume operator
    for (int i = threadIdx.x; i < L::nFrag * 32; i += blockDim.x) {
        const int off = i < nV ? D::oDwr + i : (i < 2 * nV ? D::oDws + (i - nV) : D::oLiftC + (i - 2 * nV));
        reinterpret_cast<double*>(base + L::oTab)[i] = __ldg(p.tables + off);
    }
    for (int i = threadIdx.x; i < 8 * D::NfpPad; i += blockDim.x) {
        const int row = i / D::NfpPad, c = i % D::NfpPad, face = row & 3, rev = row >> 2;
        reinterpret_cast<int*>(base + L::oNode)[i] = face < 3 ? p.nodeTab[(face * 2 + rev) * D::NfpPad + c] : 0;
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbarInit(smemAddr(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fenceProxyAsync();
    }
    __syncthreads();

    const int64_t nOct = (p.K + 7) >> 3;
    const int64_t W = (int64_t)gridDim.x * NW, w0 = (int64_t)blockIdx.x * NW + warp;
    const bool useAux = p.mode == 1 || p.A != 0.0;
    const bool sameConn = p.sameConn != 0;
    constexpr int rowsT = G::swizzled ? 4 * NT : 8;      // tensor-map rows of one octet (128-B rows under the swizzle, element rows without)
    constexpr int rowsU = 8 * NT;

    auto issueLoads = [&](int64_t oct, int s) {      // one lane
        const unsigned bar = smemAddr(bars + s), dst = smemAddr(ring + s * G::stageBytes), cdst = smemAddr(connS + s * kConnBytes);
        mbarExpectTx(bar, (useAux ? 4u : 3u) * G::tBytes + kTile + (sameConn ? 128u : 256u));
        bulkLoad(cdst, p.connT + oct * 8, 128u, bar);
        if (!sameConn) bulkLoad(cdst + 128u, p.connU + oct * 8, 128u, bar);
        tmaLoadRows(dst + G::oTin, &tmTin, (int)(oct * rowsT), bar);
        tmaLoadRows(dst + G::oU, &tmUZ, (int)(oct * rowsU), bar);
        if (useAux) tmaLoadRows(dst + G::oAux, &tmAux, (int)(oct * rowsT), bar);
        tmaLoadRows(dst + G::oGeo, &tmGeo, (int)(oct * 8), bar);
    };
    if (electOne()) {
#pragma unroll
        for (int s = 0; s < S; ++s)
            if (w0 + s * W < nOct) issueLoads(w0 + s * W, s);
    }

    // per-lane constants of the trace slots: slot = 4*kt + j = face*Nfp + i (a k-tile spans faces fLo(kt) <= fHi(kt))
    int slotI[KTC], slotI4[KTC], offOwn[KTC], offOwnU[KTC];
    bool hi[KTC];
#pragma unroll
    for (int kt = 0; kt < KTC; ++kt) {
        const int slot = 4 * kt + j;
        const bool slotValid = slot < 3 * D::Nfp;
        const int f = slotValid ? (slot >= D::Nfp) + (slot >= 2 * D::Nfp) : 2;
        hi[kt] = f != (4 * kt) / D::Nfp;
        slotI[kt] = slotValid ? slot - f * D::Nfp : 0;
        const int ownNode = reinterpret_cast<const int*>(nodeK)[f * D::NfpPad + slotI[kt]];
        offOwn[kt] = G::oTin + G::offT(e, ownNode);
        offOwnU[kt] = G::oU + G::offU(e, ownNode);
        slotI4[kt] = slotI[kt] * 4;
        pin(slotI[kt]); pin(slotI4[kt]); pin(offOwn[kt]); pin(offOwnU[kt]);
    }
    int offQ[NT];      // the lane's node pair (8nt + 2j, +1) of its element row
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) offQ[nt] = G::offT(e, 8 * nt + 2 * j);
    // velocity pairs of the volume nodes 8nt + 2j + h.  NT even: the rows of a quarter warp share the swizzle XOR's bit 0, so odd
    // rows fetch h = 1 first (disjoint banks) and swap afterwards; NT odd: consecutive elements differ in that bit, no swap needed
    const int oddU = G::swizzled ? (lane >> 2) & 1 : 0;
    int offQU[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) offQU[nt][hh] = G::oU + G::offU(e, 8 * nt + 2 * j + (hh ^ oddU));
    const int ghostBase = (int)p.ghostBase;
    auto fragR = [&](int kt, int nt) -> double { return tabS[(kt * NT + nt) * 32]; };
    auto fragS = [&](int kt, int nt) -> double { return tabS[nV + (kt * NT + nt) * 32]; };
    double tabL[FRL ? KTC * NT : 1];
    if constexpr (FRL) {
#pragma unroll
        for (int t = 0; t < KTC * NT; ++t) tabL[t] = tabS[2 * nV + t * 32];
    }
    auto fragL = [&](int kt, int nt) -> double {
        if constexpr (FRL) return tabL[kt * NT + nt];
        else return tabS[2 * nV + (kt * NT + nt) * 32];
    };

    auto faceCode = [&](const int4& cn, int kt) -> unsigned {
        const int fLo = (4 * kt) / D::Nfp, fHi = (4 * kt + 3) / D::Nfp > 2 ? 2 : (4 * kt + 3) / D::Nfp;
        const unsigned w = (unsigned)cn.w;
        return (fLo == fHi ? w >> (8 * fLo) : (hi[kt] ? w >> (8 * fHi) : w >> (8 * fLo))) & 0xffu;
    };
    auto traceOffset = [&](const int4& cn, int kt) -> int {
        const int fLo = (4 * kt) / D::Nfp, fHi = (4 * kt + 3) / D::Nfp > 2 ? 2 : (4 * kt + 3) / D::Nfp;
        const int nbLo = fLo == 0 ? cn.x : (fLo == 1 ? cn.y : cn.z), nbHi = fHi == 0 ? cn.x : (fHi == 1 ? cn.y : cn.z);
        const int nb = fLo == fHi ? nbLo : (hi[kt] ? nbHi : nbLo);
        const unsigned code = faceCode(cn, kt);
        const int node = *reinterpret_cast<const int*>(nodeK + (code & 7u) * (D::NfpPad * 4) + slotI4[kt]);
        const bool gh = code & kCodeGhost;
        return nb * (gh ? D::NfpPad : D::NpPad) + (gh ? ghostBase + slotI[kt] : node);
    };

    double TN[KTC], uxN[KTC], uyN[KTC];
    unsigned codesU = 0;
    auto gather = [&](const unsigned char* cs) {
        const int4 cT = *reinterpret_cast<const int4*>(cs + e * 16);
        int4 cU = cT;
        if (!sameConn) cU = *reinterpret_cast<const int4*>(cs + 128 + e * 16);
        codesU = (unsigned)cU.w;
#pragma unroll
        for (int kt = 0; kt < KTC; ++kt) {
            const int oT = traceOffset(cT, kt);
            const int oU = sameConn ? oT : traceOffset(cU, kt);
            TN[kt] = ldgD(p.Tin + oT);
            const double2 u = ldgD2(p.UZ + 2 * (int64_t)oU);
            uxN[kt] = u.x;
            uyN[kt] = u.y;
        }
    };

    int it = 0;
    for (int64_t oct = w0; oct < nOct; oct += W, ++it) {
        const int s = it % S;
        const unsigned parity = (unsigned)(it / S) & 1u;
        const unsigned char* st = ring + s * G::stageBytes;
        const unsigned char* cs = connS + s * kConnBytes;
        const bool valid = oct * 8 + e < p.K;

        mbarWait(smemAddr(bars + s), parity);
        gather(cs);      // neighbour traces (L2 hits), consumed after the volume term

        // ---- volume: rhs += Dwr (rx Ux T + ry Uy T) + Dws (sx Ux T + sy Uy T)   (defaultConvectionScheme.C:247-262) ------------
        double2 Tq[NT];
        double acc[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
        {
            const double2 g01 = *reinterpret_cast<const double2*>(st + G::oGeo + swz(e, 0));
            const double2 g23 = *reinterpret_cast<const double2*>(st + G::oGeo + swz(e, 2));
#pragma unroll
            for (int ntp = 0; ntp < NT; ++ntp) {
                Tq[ntp] = *reinterpret_cast<const double2*>(st + G::oTin + offQ[ntp]);
                const double2 a0 = *reinterpret_cast<const double2*>(st + offQU[ntp][0]);
                const double2 a1 = *reinterpret_cast<const double2*>(st + offQU[ntp][1]);
                double2 u[2];
                u[0].x = oddU ? a1.x : a0.x; u[0].y = oddU ? a1.y : a0.y;
                u[1].x = oddU ? a0.x : a1.x; u[1].y = oddU ? a0.y : a1.y;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int kt = 2 * ntp + h;
                    const double T = h ? Tq[ntp].y : Tq[ntp].x;
                    const double fx = u[h].x * T, fy = u[h].y * T;
                    const double ar = g01.x * fx + g01.y * fy, as = g23.x * fx + g23.y * fy;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        dmmaT(acc[nt], ar, fragR(kt, nt));
                        dmmaT(acc[nt], as, fragS(kt, nt));
                    }
                }
            }
        }

        // ---- surface: nodal LF / average flux over the 3*Nfp trace slots, lifted with the combined LIFTn (LFFlux.C:147-206) ----
        double vO[KTC], vN[KTC], TO[KTC], fsK[KTC];
        double pm[3] = {0.0, 0.0, 0.0};
        double2 nF[3];
#pragma unroll
        for (int f = 0; f < 3; ++f) nF[f] = *reinterpret_cast<const double2*>(st + G::oGeo + swz(e, kGeoN + 2 * f));
        const double2 fs01 = *reinterpret_cast<const double2*>(st + G::oGeo + swz(e, kGeoFs));
        const double2 fs2J = *reinterpret_cast<const double2*>(st + G::oGeo + swz(e, kGeoFs + 2));
        const double fsF[3] = {fs01.x, fs01.y, fs2J.x};
#pragma unroll
        for (int kt = 0; kt < KTC; ++kt) {
            const int fLo = (4 * kt) / D::Nfp, fHi = (4 * kt + 3) / D::Nfp > 2 ? 2 : (4 * kt + 3) / D::Nfp;
            TO[kt] = *reinterpret_cast<const double*>(st + offOwn[kt]);
            const double2 uo = *reinterpret_cast<const double2*>(st + offOwnU[kt]);
            const double uxo = uo.x, uyo = uo.y;
            double2 nxy = nF[fLo];
            fsK[kt] = fsF[fLo];
            if (fLo != fHi) {
                nxy.x = hi[kt] ? nF[fHi].x : nxy.x;
                nxy.y = hi[kt] ? nF[fHi].y : nxy.y;
                fsK[kt] = hi[kt] ? fsF[fHi] : fsK[kt];
            }
            double uxn = uxN[kt], uyn = uyN[kt];
            if (p.anyReflect) {      // reflective U patch somewhere in the mesh (uniform branch): mirror the exterior velocity
                int4 cw;
                cw.w = (int)codesU;
                if (faceCode(cw, kt) & kCodeReflect) {
                    const double d2 = 2.0 * (uxn * nxy.x + uyn * nxy.y);
                    uxn -= d2 * nxy.x;
                    uyn -= d2 * nxy.y;
                }
            }
            vO[kt] = nxy.x * uxo + nxy.y * uyo;
            vN[kt] = nxy.x * uxn + nxy.y * uyn;
            // (a padding slot, 4kt+j >= 3Nfp, aliases node 0 of face 2: finite members of that face's set, zero lift fragments)
            const double m = dmax(fabs(vO[kt]), fabs(vN[kt]));
            if (fLo == fHi) pm[fLo] = dmax(pm[fLo], m);
            else {
                pm[fLo] = (!hi[kt] && m > pm[fLo]) ? m : pm[fLo];
                pm[fHi] = (hi[kt] && m > pm[fHi]) ? m : pm[fHi];
            }
        }
        // one maxV per face (LFFlux.C:189-196): the slots of a face are spread over the 4 lanes of the element
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            pm[f] = dmax(pm[f], __shfl_xor_sync(0xffffffffu, pm[f], 1));
            pm[f] = dmax(pm[f], __shfl_xor_sync(0xffffffffu, pm[f], 2));
        }
        const double dissOn = p.fluxKind == 1 ? 0.5 : 0.0, fluxOn = p.fluxKind != 3 ? 1.0 : 0.0;
#pragma unroll
        for (int kt = 0; kt < KTC; ++kt) {
            const int fLo = (4 * kt) / D::Nfp, fHi = (4 * kt + 3) / D::Nfp > 2 ? 2 : (4 * kt + 3) / D::Nfp;
            const double maxV = fLo == fHi ? pm[fLo] : (hi[kt] ? pm[fHi] : pm[fLo]);
            double fl = (vO[kt] * TO[kt] + vN[kt] * TN[kt]) * 0.5 + (dissOn * maxV) * (TO[kt] - TN[kt]);
            fl *= fsK[kt] * fluxOn;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) dmmaT(acc[nt], fl, fragL(kt, nt));
        }

        // ---- explicit update ------------------------------------------------------------------------------------------------
        double2 o[NT], r[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const double2 qx = useAux ? *reinterpret_cast<const double2*>(st + G::oAux + offQ[nt]) : make_double2(0.0, 0.0);
            if (p.mode == 0) {
                o[nt].x = p.B * (Tq[nt].x + p.dt * acc[nt][0]) + p.A * qx.x;
                o[nt].y = p.B * (Tq[nt].y + p.dt * acc[nt][1]) + p.A * qx.y;
                r[nt] = o[nt];
            } else {
                r[nt].x = p.A * qx.x + p.dt * acc[nt][0];
                r[nt].y = p.A * qx.y + p.dt * acc[nt][1];
                o[nt].x = Tq[nt].x + p.B * r[nt].x;
                o[nt].y = Tq[nt].y + p.B * r[nt].y;
            }
        }
        if (oct * 8 + 8 > p.K) {      // warp-uniform: the last, ragged octet - its padding rows stay zero
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                if (!valid) o[nt] = r[nt] = make_double2(0.0, 0.0);
        }
        __syncwarp();      // every lane has read what it needs from the stage: refill it
        if (electOne()) {
            const int64_t octr = oct + (int64_t)S * W;
            if (octr < nOct) issueLoads(octr, s);
        }
        const int64_t g0 = (oct * 8 + e) * D::NpPad + 2 * j;      // node pair (8nt+2j, +1) of element row e
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            *reinterpret_cast<double2*>(p.Tout + g0 + 8 * nt) = o[nt];
            if (p.mode == 1) *reinterpret_cast<double2*>(p.res + g0 + 8 * nt) = r[nt];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encodeTiled()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        const cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (err != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr)
            throw std::runtime_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// [rows][16] doubles, box = one octet (8 rows) or two, 128B swizzle
CUtensorMap rowsMap(const double* ptr, int64_t rows, unsigned boxRows = 8)
{
    CUtensorMap m;
    const cuuint64_t gdim[2] = {16, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {16, boxRows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encodeTiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), gdim, gstride, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

// plane of `rows` rows of `rowDoubles` doubles, box = `boxRows` rows; 128B swizzle for 128-B rows, none for wider ones
CUtensorMap planeMap(const double* ptr, int64_t rows, unsigned rowDoubles, unsigned boxRows)
{
    CUtensorMap m;
    const cuuint64_t gdim[2] = {rowDoubles, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {rowDoubles * 8ull};
    const cuuint32_t box[2] = {rowDoubles, boxRows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encodeTiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), gdim, gstride, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, rowDoubles == 16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

template <int N, int S, int NW, int MB, bool FRL = false>
void launchTmaWide(const AdvectParams& p, cudaStream_t st)
{
    using D = Dims<N>;
    using G = WideTile<D::NT>;
    constexpr int NT = D::NT;
    const size_t smem = 1024 + (size_t)WideLayout<N, S, NW>::total;
    if (p.ghostBase + (p.planeStrideT - p.ghostBase) >= (int64_t)1 << 31) throw std::runtime_error("advect tma kernel: plane too large for 32-bit trace offsets");
    static int gridFor[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!gridFor[dev & 63]) {
        cudaError_t err = cudaFuncSetAttribute(advectStageTmaWideKernel<N, S, NW, MB, FRL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute(advect tma wide): ") + cudaGetErrorString(err));
        int blocks = 0, sms = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, advectStageTmaWideKernel<N, S, NW, MB, FRL>, 32 * NW, smem);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (blocks < 1) throw std::runtime_error("advect tma wide kernel does not fit on an SM");
        gridFor[dev & 63] = blocks * sms;
    }
    const int64_t Kpad = (p.K + 7) / 8 * 8, nOct = Kpad / 8;
    const int grid = (int)std::min<int64_t>(gridFor[dev & 63], (nOct + NW - 1) / NW);
    const bool useAux = p.mode == 1 || p.A != 0.0;
    if (!p.UZ) throw std::runtime_error("advect tma kernel: the interleaved velocity copy is missing");
    const double* auxP = p.mode == 1 ? p.res : (useAux ? p.Taux : p.Tin);
    CUtensorMap tin, uz, aux;
    if (G::swizzled) {      // rows of 128 B: NT/2 per element (T), NT per element (velocity pairs)
        tin = planeMap(p.Tin, Kpad * (NT / 2), 16, 4 * NT);
        aux = planeMap(auxP, Kpad * (NT / 2), 16, 4 * NT);
    } else {                // element rows as they are
        tin = planeMap(p.Tin, Kpad, 8 * NT, 8);
        aux = planeMap(auxP, Kpad, 8 * NT, 8);
    }
    uz = planeMap(p.UZ, Kpad * NT, 16, 8 * NT);
    const CUtensorMap geo = planeMap(p.geo, Kpad, 16, 8);
    advectStageTmaWideKernel<N, S, NW, MB, FRL><<<grid, 32 * NW, smem, st>>>(p, tin, uz, aux, geo);
}

// A/B aid for the wide kernel: HDG_ADVW_CFG selects (stages, warps per block, resident blocks)
int wideConfig()
{
    static int cfg = -1;
    if (cfg < 0) {
        const char* v = std::getenv("HDG_ADVW_CFG");
        cfg = v ? std::atoi(v) : 0;
    }
    return cfg;
}

// shared memory per block = NW * S * (stage + 256 B) + fragments: N=5: 7 KB stages, 13 KB of fragments; N=6: 9 KB, 22.5 KB
template <int N>
void launchTmaWideCfg(const AdvectParams& p, cudaStream_t st);
template <>
void launchTmaWideCfg<5>(const AdvectParams& p, cudaStream_t st)
{
    switch (wideConfig()) {
        case 1: launchTmaWide<5, 2, 6, 2>(p, st); break;       // 12 warps per SM in two blocks (0.53 of the HBM peak)
        case 2: launchTmaWide<5, 3, 4, 2>(p, st); break;       // 8 warps per SM, 3 stages (0.62 of the HBM peak)
        case 3: launchTmaWide<5, 3, 8, 1, true>(p, st); break; // 8 warps, lift fragments in registers (0.64)
        default: launchTmaWide<5, 2, 12, 1>(p, st); break;     // 12 warps in one block, one fragment table (0.70)
    }
}
template <>
void launchTmaWideCfg<6>(const AdvectParams& p, cudaStream_t st)
{
    switch (wideConfig()) {
        case 1: launchTmaWide<6, 2, 8, 1, true>(p, st); break; // lift fragments in registers (0.52)
        case 2: launchTmaWide<6, 2, 9, 1>(p, st); break;       // 9 warps per SM (0.46)
        case 3: launchTmaWide<6, 2, 9, 1, true>(p, st); break; // (0.38, spills)
        default: launchTmaWide<6, 2, 8, 1>(p, st); break;      // 8 warps in one block, 2 stages (0.56)
    }
}

template <>
void launchTmaWideCfg<7>(const AdvectParams& p, cudaStream_t st)
{
    switch (wideConfig()) {
        case 1: launchTmaWide<7, 2, 7, 1>(p, st); break;
        case 2: launchTmaWide<7, 2, 6, 1>(p, st); break;
        default: launchTmaWide<7, 2, 8, 1>(p, st); break;      // 11 KB stages, 32.5 KB of fragments: 8 warps x 2 stages fill the SM
    }
}
template <>
void launchTmaWideCfg<2>(const AdvectParams& p, cudaStream_t st)
{
    switch (wideConfig()) {
        case 1: launchTmaWide<2, 4, 16, 1>(p, st); break;      // (0.53)
        case 2: launchTmaWide<2, 3, 20, 1>(p, st); break;      // (0.52)
        case 3: launchTmaWide<2, 2, 24, 1>(p, st); break;      // (0.60)
        default: launchTmaWide<2, 3, 16, 1>(p, st); break;     // 64-B rows: 3 KB stages, 16 warps in one block (0.60; 4 blocks of 4 warps: 0.55)
    }
}
template <>
void launchTmaWideCfg<1>(const AdvectParams& p, cudaStream_t st)
{
    switch (wideConfig()) {
        case 1: launchTmaWide<1, 4, 16, 1>(p, st); break;      // (0.39)
        case 2: launchTmaWide<1, 3, 20, 1>(p, st); break;      // (0.39)
        case 3: launchTmaWide<1, 3, 16, 1>(p, st); break;      // (0.44; 4 blocks of 4 warps: 0.39)
        default: launchTmaWide<1, 2, 24, 1>(p, st); break;     // 24 warps in one block, 2 stages (0.49)
    }
}

template <int N, int S, int MB, bool DS, int NW = kWarps>
void launchTmaCfg(const AdvectParams& p, cudaStream_t st)
{
    using D = Dims<N>;
    (void)sizeof(D);
    const size_t smem = 1024 + (size_t)TmaLayout<N, S, DS, NW>::total;
    if (p.ghostBase + (p.planeStrideT - p.ghostBase) >= (int64_t)1 << 31) throw std::runtime_error("advect tma kernel: plane too large for 32-bit trace offsets");
    static int gridFor[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!gridFor[dev & 63]) {
        cudaError_t err = cudaFuncSetAttribute(advectStageTmaKernel<N, S, MB, DS, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute(advect tma): ") + cudaGetErrorString(err));
        int blocks = 0, sms = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, advectStageTmaKernel<N, S, MB, DS, NW>, 32 * NW, smem);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (blocks < 1) throw std::runtime_error("advect tma kernel does not fit on an SM");
        gridFor[dev & 63] = blocks * sms;
    }
    const int64_t Kpad = (p.K + 7) / 8 * 8, nOct = Kpad / 8;
    const int grid = (int)std::min<int64_t>(gridFor[dev & 63], (nOct + NW - 1) / NW);
    const bool useAux = p.mode == 1 || p.A != 0.0;
    if (!p.UZ) throw std::runtime_error("advect tma kernel: the interleaved velocity copy is missing");
    const CUtensorMap tin = rowsMap(p.Tin, Kpad), uz = rowsMap(p.UZ, 2 * Kpad, 16);
    const CUtensorMap aux = rowsMap(p.mode == 1 ? p.res : (useAux ? p.Taux : p.Tin), Kpad);
    const CUtensorMap geo = rowsMap(p.geo, Kpad), tout = rowsMap(p.Tout, Kpad), res = rowsMap(p.mode == 1 ? p.res : p.Tout, Kpad);
    advectStageTmaKernel<N, S, MB, DS, NW><<<grid, 32 * NW, smem, st>>>(p, tin, uz, aux, geo, tout, res);
}

int tmaConfig()
{
    static int cfg = -1;
    if (cfg < 0) {
        // A/B aid: HDG_ADV_CFG=0 legacy advectStageKernel, 2 = result through a shared-memory tile + TMA store, 3 = three blocks of 4
        // warps per SM (round 1: N=4 0.70 of the HBM peak), 5 = 10 warps x 4 stages (0.63); default: ONE block of 12 warps per SM with
        // direct stores (0.73) - the warps of an SM then work on consecutive octets and their neighbour-trace gathers share L1 lines
        const char* v = std::getenv("HDG_ADV_CFG");
        cfg = v ? std::atoi(v) : 1;
    }
    return cfg;
}

}  // namespace

// HDG_ADVW_ORDERS (A/B aid): bit N set = order N runs the wide-row kernel
unsigned wideOrders()
{
    static int m = -1;
    if (m < 0) {
        const char* v = std::getenv("HDG_ADVW_ORDERS");
        m = v ? std::atoi(v) : ((1 << 1) | (1 << 2) | (1 << 5) | (1 << 6) | (1 << 7));
    }
    return (unsigned)m;
}

bool advectUsesTma(int N) { return tmaConfig() != 0 && (N == 3 || N == 4 || (N >= 1 && N <= 7 && (wideOrders() >> N & 1u))); }

// returns false when this order / configuration is served by the legacy kernel
bool launchAdvectStageTma(int N, const AdvectParams& p, cudaStream_t st)
{
    const int cfg = tmaConfig();
    if (cfg == 0 || !advectUsesTma(N)) return false;
    switch (N) {
        case 1: launchTmaWideCfg<1>(p, st); return true;
        case 2: launchTmaWideCfg<2>(p, st); return true;
        case 5: launchTmaWideCfg<5>(p, st); return true;
        case 6: launchTmaWideCfg<6>(p, st); return true;
        case 7: launchTmaWideCfg<7>(p, st); return true;
        default: break;
    }
#define HDG_TMA_CASE(NN)                                              \
    case NN:                                                          \
        if (cfg == 2) launchTmaCfg<NN, 3, 3, false>(p, st);        \
        else if (cfg == 3) launchTmaCfg<NN, 3, 3, true>(p, st);    \
        else if (cfg == 5) launchTmaCfg<NN, 4, 1, true, 10>(p, st);\
        else launchTmaCfg<NN, 3, 1, true, 12>(p, st);              \
        break;
    switch (N) {
        HDG_TMA_CASE(3)
        HDG_TMA_CASE(4)
        default: return false;
    }
#undef HDG_TMA_CASE
    return true;
}

}  // namespace hdg
