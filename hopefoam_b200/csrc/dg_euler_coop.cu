// Co-scheduled split Euler stage (sm_100a, FP64, DMMA.8x8x4): the two kernels of dg_euler_split.cu as two warp ROLES of one launch.
//
// The split stage runs the face-flux pass and the element pass back to back.  The face pass is bound by memory latency, the L1 tag
// stage and DRAM, with the FP64 pipe 60 % busy; the element pass keeps the pipe 80 % busy and leaves DRAM idle - and the flux array
// and the state make a round trip through HBM in between.  Here every block holds EW element warps and ONE face warp:
//
//   face warp      walks the dgFaces in the order in which the element warps will need them (sorted by the lower of the two adjacent
//                  octets), one Roe flux per face exactly as eulerFaceFluxKernel, and publishes its progress: a counter per chunk of
//                  kChunk face-octets, and the number of leading complete chunks (`progress[0]`), advanced by whichever warp
//                  completes the lowest open chunk;
//   element warps  volume term as eulerElemKernel; before the lift they wait until the chunks that hold their octet's faces are
//                  complete (`octNeed[octet]`, acquire load + nanosleep), then gather the flux records with L2 loads (ld.global.cg:
//                  the lines were written by other SMs during this launch).
//
// The face warps' point-wise chains and gathers run in the issue slots and pipe cycles the element warps leave idle, the flux
// records and the state rows are read out of L2 shortly after they were written / first touched, and the stage is ONE launch.
// No block-level synchronisation after the prologue; the face warps never wait, so the element warps' waits always end.
// A per-SM arrival counter rotates the position of the face warp inside the block, so that every scheduler of an SM gets one.
#include <cuda_runtime.h>

#include <algorithm>
#include <stdexcept>
#include <string>

#include "dg_kernels.cuh"
#include "dg_device.cuh"

namespace hdg {

#define HDG_COOP_EW(N) ((N) <= 4 ? 3 : ((N) <= 6 ? 4 : 7))          // element warps per block (+ 1 face warp)
#define HDG_COOP_MB(N) ((N) <= 4 ? 4 : ((N) <= 6 ? 2 : 1))          // resident blocks per SM
#define HDG_COOP_THREADS(N) ((HDG_COOP_EW(N) + 1) * 32)

__device__ __forceinline__ int ldAcquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned smId()
{
    unsigned v;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
    return v;
}

template <int N>
__global__ void __launch_bounds__(HDG_COOP_THREADS(N), HDG_COOP_MB(N)) eulerCoopStageKernel(const StageParams p)
{
    using D = Dims<N>;
    constexpr int SL = D::fluxSlots;
    constexpr int EW = HDG_COOP_EW(N);
    extern __shared__ __align__(128) double smem[];
    const double* tab = smem;
    int* nodeTab = reinterpret_cast<int*>(smem + D::splitTableDoubles);
    __shared__ unsigned long long tableBar;
    __shared__ int faceWarpOfBlock;
    for (int i = threadIdx.x; i < D::nodeTabInts; i += blockDim.x) nodeTab[i] = p.nodeTab[i];
    if (threadIdx.x == 0) faceWarpOfBlock = atomicAdd(p.coopSmSlots + (smId() & 255u), 1) % (EW + 1);
    stageTables(smem, p.splitTables, D::splitTableDoubles, &tableBar);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int e = lane >> 2, j = lane & 3;
    const int warp = threadIdx.x >> 5;
    const double gm1 = p.gamma - 1.0;

    if (warp == faceWarpOfBlock) {
        // =========================================================================================================================
        // face role (body of eulerFaceFluxKernel; entries in need order, progress published)
        // =========================================================================================================================
        double bIf[D::FGT][D::FKT];
#pragma unroll
        for (int fgt = 0; fgt < D::FGT; ++fgt)
#pragma unroll
            for (int fkt = 0; fkt < D::FKT; ++fkt) bIf[fgt][fkt] = __ldg(p.tables + D::oIf + (fgt * D::FKT + fkt) * 32 + lane);
        const int64_t nFaceOct = p.coopFaceOct;
        // face-octets are handed out by a ticket counter, not by a fixed stride: whatever part of the grid is resident makes progress on
        // the LOWEST open face-octets, so element warps can never wait for a block that has not been scheduled yet
        auto ticket = [&]() -> int64_t {
            int t = 0;
            if (lane == 0) t = atomicAdd(p.coopTicket, 1);
            return (int64_t)__shfl_sync(0xffffffffu, t, 0);
        };
        auto entryOf = [&](int64_t it_) -> int2 { return it_ < nFaceOct ? __ldg(p.faceSorted + it_ * 8 + e) : make_int2(0, -1); };
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
        int64_t t0 = ticket(), t1 = ticket(), t2 = ticket();
        int2 en0 = entryOf(t0), en1 = entryOf(t1), en2 = entryOf(t2);
        int4 cn0 = __ldg(p.conn + (en0.x >> 2)), cn1 = __ldg(p.conn + (en1.x >> 2));
        double amN[4][D::FKT], anN[4][D::FKT];
        unsigned codeN = 0;
        double2 nxyN = make_double2(0.0, 0.0);
        if (t0 < nFaceOct) loadTraces(en0.x, cn0, amN, anN, codeN, nxyN);
        while (t0 < nFaceOct) {
            const int64_t it = t0;
            const int fid = en0.y;
            double am[4][D::FKT], an[4][D::FKT];
#pragma unroll
            for (int f = 0; f < 4; ++f)
#pragma unroll
                for (int fkt = 0; fkt < D::FKT; ++fkt) { am[f][fkt] = amN[f][fkt]; an[f][fkt] = anN[f][fkt]; }
            const unsigned code = codeN;
            const double2 nxy = nxyN;
            if (t1 < nFaceOct) loadTraces(en1.x, cn1, amN, anN, codeN, nxyN);
            {   // the index pipeline one step on
                const int64_t t3 = ticket();
                const int4 cn2 = __ldg(p.conn + (en2.x >> 2));
                const int2 en3 = entryOf(t3);
                t0 = t1; en0 = en1; cn0 = cn1;
                t1 = t2; en1 = en2; cn1 = cn2;
                t2 = t3; en2 = en3;
            }
            double* fb = p.flux + (int64_t)(fid < 0 ? 0 : fid) * (4 * SL);
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
                    roeFlux(qM, qP, nxy.x, nxy.y, gm1, fl[h]);
                }
                const int s0 = fgt * 8 + 2 * j;
                if (fid >= 0 && s0 < SL) {
#pragma unroll
                    for (int f = 0; f < 4; ++f) __stcg(reinterpret_cast<double2*>(fb + f * SL + s0), make_double2(fl[0][f], fl[1][f]));
                }
            }
            // publish: every lane's stores are visible device-wide before lane 0 counts this face-octet in its chunk; the warp that
            // completes a chunk moves `progress[0]` over every leading complete chunk
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                const int chunk = (int)(it / kCoopChunk);
                const int nChunks = p.coopChunks;
                auto fullOf = [&](int c) -> int { return (int)min((int64_t)kCoopChunk, nFaceOct - (int64_t)c * kCoopChunk); };
                const int old = atomicAdd(p.coopProgress + 1 + chunk, 1);
                if (old + 1 == fullOf(chunk)) {
                    __threadfence();
                    int pfx = ldAcquire(p.coopProgress);
                    while (pfx < nChunks && ldAcquire(p.coopProgress + 1 + pfx) == fullOf(pfx)) ++pfx;
                    atomicMax(p.coopProgress, pfx);
                }
            }
        }
        return;
    }

    // =============================================================================================================================
    // element role (body of eulerElemKernel; waits for its faces, flux gathers from L2)
    // =============================================================================================================================
    const int ew = warp - (warp > faceWarpOfBlock ? 1 : 0);
    const int64_t warpsPerGrid = (int64_t)gridDim.x * EW;
    const int64_t warpId = (int64_t)blockIdx.x * EW + ew;
    const int64_t n1 = p.octEnd - p.octBegin, nTot = p.octList ? p.nList : n1 + (p.octEnd2 - p.octBegin2);
    auto octOf = [&](int64_t i) -> int64_t { return p.octList ? (int64_t)__ldg(p.octList + i) : (i < n1 ? p.octBegin + i : p.octBegin2 + (i - n1)); };
    for (int64_t it = warpId; it < nTot; it += warpsPerGrid) {
        const int64_t oct = octOf(it);
        const int64_t elem = oct * 8 + e;
        const bool valid = elem < p.K;
        const int64_t el = valid ? elem : p.K - 1;
        const double* geo = p.geo + el * 16;
        const int64_t eoff = el * D::NpPad;

        double a[4][D::KT];
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
            for (int kt = 0; kt < D::KT; ++kt) a[f][kt] = __ldg(p.qin[f] + eoff + kt * 4 + j);

        double acc[4][D::NT][2];
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
            for (int nt = 0; nt < D::NT; ++nt) acc[f][nt][0] = acc[f][nt][1] = 0.0;

        const int4 cn = __ldg(p.conn + el);
        const int4 ef = __ldg(p.elemFace + el);
        const int need = __ldg(p.octNeed + oct);
        if (p.mode == 0 && p.A != 0.0) prefetchL1((j == 0 ? p.qaux[0] : j == 1 ? p.qaux[1] : j == 2 ? p.qaux[2] : p.qaux[3]) + eoff);

        // ---- volume term (as eulerElemKernel) ----------------------------------------------------------------------------------
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
                    for (int f = 0; f < 4; ++f) dmma(c[f], a[f][kt], b);
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
                        for (int f = 1; f < 4; ++f) dmma(acc[f][nt], Gr[h][f], br);
                    }
#pragma unroll
                    for (int nt = 0; nt < D::NT; ++nt) {
                        const double bs = ts[(h * D::NT + nt) * 32];
#pragma unroll
                        for (int f = 1; f < 4; ++f) dmma(acc[f][nt], Gs[h][f], bs);
                    }
                }
            };
            // density on the weak nodal derivative (see eulerElemKernel)
#pragma unroll
            for (int kt = 0; kt < D::KT; ++kt) {
                const double ar = rx * a[1][kt] + ry * a[2][kt], as = sx * a[1][kt] + sy * a[2][kt];
#pragma unroll
                for (int nt = 0; nt < D::NT; ++nt) {
                    dmma(acc[0][nt], ar, tab[D::sDwr + (kt * D::NT + nt) * 32 + lane]);
                    dmma(acc[0][nt], as, tab[D::sDws + (kt * D::NT + nt) * 32 + lane]);
                }
            }
            double c[4][2], Gr[2][4], Gs[2][4];
            interp(0, c);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double q[4] = {c[0][h], c[1][h], c[2][h], c[3][h]};
                eulerVolumeFlux(q, rx, ry, sx, sy, gm1, Gr[h], Gs[h]);
            }
#pragma unroll 1
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
            {   // next octet of this warp towards L1
                const int64_t itn = it + warpsPerGrid;
                if (itn < nTot) {
                    const int64_t eln = min(octOf(itn) * 8 + e, p.K - 1);
                    prefetchL1((j == 0 ? p.qin[0] : j == 1 ? p.qin[1] : j == 2 ? p.qin[2] : p.qin[3]) + eln * D::NpPad);
                    if (j == 0) prefetchL1(p.geo + eln * 16);
                    if (j == 1) prefetchL1(p.conn + eln);
                    if (j == 2) prefetchL1(p.elemFace + eln);
                }
            }
            project(D::GT - 1, Gr, Gs);
        }

        // ---- wait for this octet's faces, gather their fluxes out of L2 -------------------------------------------------------------
        if (lane == 0) {
            while (ldAcquire(p.coopProgress) < need) __nanosleep(128);
        }
        __syncwarp();
        double fl[D::KTL][4];
#pragma unroll
        for (int kt = 0; kt < D::KTL; ++kt) {
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
            for (int f = 0; f < 4; ++f) fl[kt][f] = sc * __ldcg(fb + f * SL);
        }
#pragma unroll
        for (int kt = 0; kt < D::KTL; ++kt)
#pragma unroll
            for (int nt = 0; nt < D::NT; ++nt) {
                const double b = tab[D::sLiftC + (kt * D::NT + nt) * 32 + lane];
#pragma unroll
                for (int f = 0; f < 4; ++f) dmma(acc[f][nt], fl[kt][f], b);
            }

        // ---- explicit update (as eulerElemKernel) -----------------------------------------------------------------------------------
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

// -------------------------------------------------------------------------------------------------------------------------------
// launcher
// -------------------------------------------------------------------------------------------------------------------------------
namespace {
struct CoopCfg { int blocks = 0; size_t smem = 0; bool ok = false; };

template <int N>
CoopCfg& coopCfgT()
{
    static CoopCfg cfg[64];
    int dev = 0;
    cudaGetDevice(&dev);
    CoopCfg& c = cfg[dev & 63];
    if (!c.ok) {
        using D = Dims<N>;
        c.smem = sizeof(double) * D::splitTableDoubles + sizeof(int) * D::nodeTabInts;
        cudaError_t err = cudaFuncSetAttribute(eulerCoopStageKernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
        if (err != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute(eulerCoopStageKernel): ") + cudaGetErrorString(err));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.blocks, eulerCoopStageKernel<N>, HDG_COOP_THREADS(N), c.smem);
        if (c.blocks < 1) throw std::runtime_error("co-scheduled Euler stage: kernel does not fit on this device");
        c.ok = true;
    }
    return c;
}

template <int N>
void launchCoopT(const StageParams& p, int smCount, cudaStream_t st)
{
    CoopCfg& c = coopCfgT<N>();
    // every block must be resident at once (the element warps wait for face warps): the grid never exceeds the resident capacity
    const int64_t nOct = p.octList ? p.nList : (p.octEnd - p.octBegin) + (p.octEnd2 - p.octBegin2);
    const int64_t want = std::max<int64_t>((nOct + HDG_COOP_EW(N) - 1) / HDG_COOP_EW(N), 1);
    const int grid = (int)std::min<int64_t>((int64_t)smCount * c.blocks, want);
    eulerCoopStageKernel<N><<<grid, HDG_COOP_THREADS(N), c.smem, st>>>(p);
}
}  // namespace

bool eulerCoopAvailable(int N) { return N >= 3 && N <= 8; }

void launchEulerCoop(int N, const StageParams& p, int smCount, cudaStream_t st)
{
    switch (N) {
        case 3: launchCoopT<3>(p, smCount, st); break;
        case 4: launchCoopT<4>(p, smCount, st); break;
        case 5: launchCoopT<5>(p, smCount, st); break;
        case 6: launchCoopT<6>(p, smCount, st); break;
        case 7: launchCoopT<7>(p, smCount, st); break;
        case 8: launchCoopT<8>(p, smCount, st); break;
        default: throw std::runtime_error("co-scheduled Euler stage: orders 3..8");
    }
}

}  // namespace hdg
