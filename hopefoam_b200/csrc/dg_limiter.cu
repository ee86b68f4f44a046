// `Triangle` slope limiter on the device (Trianglelimite.C:61-864): three launches, each needing the previous one complete for ALL cells.
//
//   A  limAveragesKernel     passes 1+2: cell averages / centroid / A0, the three vertex values of every field -> vtx, ghost cells
//   B  limGradientsKernel    passes 3+4: every element evaluates the gradients of its three faces itself (the formula is symmetric in
//                            the two elements: bit-identical on both sides, nothing stored per face), A_2-weighted mean -> CV; ghost cells
//   C  limReconstructKernel  pass 5: limited gradients from the three neighbour cells, P1 reconstruction, back to conserved variables
//
// Work distribution = the stage kernels' octet layout: a warp owns 8 consecutive elements, lane = 4*e + j.  In A and C lane j moves the
// node pairs (8*nt + 2j, +1) of its element as 16-B vectors, so a warp request covers the 8 element rows completely (whole 128-B lines
// at N = 3, 4) - the first version walked one row per THREAD, 32 lines per request, and ran at 10 % of the HBM roofline.  In B lane
// j < 3 owns face j; the per-element reductions are 4-lane shuffles in the reference's summation order.
// HBM traffic per element (N = 4): A reads the four planes (512 B) and writes 21 doubles; B gathers compact per-cell arrays (L2 hits
// for the neighbours) and writes 8; C reads 3 x 8 gradient entries + 8 constants and writes the four planes (512 B).
//
// Parity: tests/test_gpu_limiter.py (device, through hdg_euler_limit) and tests/test_limiter_core_host.py (the same inline functions of
// dg_limiter_core.hpp in host loops, five-pass and fused forms).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "dg_limiter_core.hpp"

namespace hdg {

namespace {

#ifndef HDG_LIM_MB
#define HDG_LIM_MB 4
#endif
constexpr int kLimThreads = 256;      // 8 warps = 64 elements per block
constexpr int kMaxNp = 72;            // NpPad <= 72 (N <= 10)

__device__ __forceinline__ double shflD(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// The planes (A reads them, C writes them: 2 x 256 MB at 500 k triangles) stream through the 126 MB L2 once; the work records between
// the three launches (cell, vertex, gradient: 112 MB) are what the next launch gathers from.  The plane accesses carry an evict-first
// policy so that they do not push the records out before they are read again.
__device__ __forceinline__ unsigned long long streamPolicy()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double2 ldStream(const double* p, unsigned long long pol)
{
    double2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void stStream(double* p, double2 v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}

// A: NT = NpPad / 8 node tiles per element row, fully unrolled so that all 4*NT vector loads of a lane are in flight together
template <int NT>
__global__ void __launch_bounds__(kLimThreads) limAveragesKernel(const LimiterView v)
{
    __shared__ double sw[kMaxNp];      // mpp (0 beyond Np)
    for (int i = threadIdx.x; i < kMaxNp; i += blockDim.x) sw[i] = i < v.Np ? v.mpp[i] : 0.0;
    // sum_j w_j (a_j, b_j, c_j) and sum_j w_j come with the view (v.cabc, summed once on the host): the first version had thread 0 of
    // every block sum them while the other 255 waited at the barrier (ncu: 20 % of the warp time)
    const double sabc[4] = {v.cabc[0], v.cabc[1], v.cabc[2], v.cabc[3]};
    __syncthreads();
    const int lane = threadIdx.x & 31, e = lane >> 2, j = lane & 3;
    const unsigned long long pol = streamPolicy();
    const bool stream = v.streamPlanes != 0;
    const int64_t nOct = (v.K + 7) >> 3, W = (int64_t)gridDim.x * (blockDim.x >> 5);
    // the grid may be persistent (launchT): a warp walks the octets w, w + W, ... and the block prologue is paid once
    for (int64_t oct = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); oct < nOct; oct += W) {
    const int64_t k = oct * 8 + e;
    const bool valid = k < v.K;
    const int64_t kk = valid ? k : v.K - 1;
    double2 q[4][NT];
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) q[f][nt] = stream ? ldStream(v.q[f] + kk * (NT * 8) + nt * 8 + 2 * j, pol) : *reinterpret_cast<const double2*>(v.q[f] + kk * (NT * 8) + nt * 8 + 2 * j);
    const int nv[3] = {limVertexNode(v, 0), limVertexNode(v, 1), limVertexNode(v, 2)};
    double a[4] = {0, 0, 0, 0};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int n0 = nt * 8 + 2 * j;
        const double w0 = sw[n0], w1 = sw[n0 + 1];
#pragma unroll
        for (int f = 0; f < 4; ++f) a[f] += q[f][nt].x * w0 + q[f][nt].y * w1;
        if (valid) {      // the lane that holds a vertex node drops its four fields into the element's vertex record (one 32-B sector)
#pragma unroll
            for (int vert = 0; vert < 3; ++vert)
                if (n0 == (nv[vert] & ~1)) {
                    const bool hi = nv[vert] & 1;
                    double* d = v.vtx + k * 12 + vert * 4;
                    *reinterpret_cast<double2*>(d) = make_double2(hi ? q[0][nt].y : q[0][nt].x, hi ? q[1][nt].y : q[1][nt].x);
                    *reinterpret_cast<double2*>(d + 2) = make_double2(hi ? q[2][nt].y : q[2][nt].x, hi ? q[3][nt].y : q[3][nt].x);
                }
        }
    }
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        a[f] += __shfl_xor_sync(0xffffffffu, a[f], 1);
        a[f] += __shfl_xor_sync(0xffffffffu, a[f], 2);
    }
    if (valid && j == 0) {
        const double* p = v.verts + 6 * k;
        const double J = 0.25 * ((p[2] - p[0]) * (p[5] - p[1]) - (p[3] - p[1]) * (p[4] - p[0]));
        const double rec[8] = {a[0], a[1], a[2], a[3], sabc[0] * p[0] + sabc[1] * p[2] + sabc[2] * p[4],
                               sabc[0] * p[1] + sabc[1] * p[3] + sabc[2] * p[5], sabc[3] * J * (2.0 / 3.0), 0.0};
        limStore(v.cell + 8 * k, 8, rec);
    }
    __syncwarp();      // the record of element k is visible to the lanes that build its ghost cells
    if (valid && j < 3) limGhostCell(v, k, j);
    }
}

// B.  The first fused version let every lane fetch its records itself: 19 requests per warp, each touching ~24 different 128-B lines -
// the kernel ran at the speed of the L1 tag stage (one line per cycle: ~450 cycles per octet, 21 % of the DRAM bandwidth).  Now the warp
// moves the data cooperatively: the octet's own cell / vertex / coordinate records are three contiguous ranges (coalesced 16-B loads),
// and a neighbour's cell record and its two end-point states are fetched by the four lanes of the element together (one request per
// face round covers the record: 8 lines instead of 24).  Everything is parked in a per-warp shared-memory tile and read back by the
// lane that owns the face; a neighbour record's 16-B chunks are XORed with (e & 1 | face << 1), so the six active lanes of a quarter
// warp read six different bank groups.  Boundary faces (rare) keep the generic path of dg_limiter_core.hpp.
constexpr int kOwnCell = 0, kOwnVtx = 64, kOwnVerts = 160, kNb = 208, kWarpTile = kNb + 3 * 8 * 16;      // doubles per warp (4.7 KB)

__device__ __forceinline__ void cpAsync16(void* smem, const void* gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpAsyncWaitAll() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// OCT octets per warp: the copies of all of them (cp.async, no registers held) are in flight together, so the two dependent trips
// to memory (connectivity -> neighbour records) are paid once per OCT octets instead of once per octet
template <int MB, int OCT>
__global__ void __launch_bounds__(kLimThreads, MB) limGradientsKernel(const LimiterView v)
{
    extern __shared__ __align__(16) double tileRaw[];      // [warp][OCT][kWarpTile]
    const int lane = threadIdx.x & 31, e = lane >> 2, j = lane & 3;
    const int64_t kw = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (8 * OCT);
    if (kw >= v.K) return;      // warp-uniform
    double* wtile = tileRaw + (size_t)(threadIdx.x >> 5) * (OCT * kWarpTile);
    int4 cn[OCT];
#pragma unroll
    for (int o = 0; o < OCT; ++o) {
        const int64_t k = kw + 8 * o + e;
        cn[o] = make_int4(0, 0, 0, 0);
        if (k < v.K) cn[o] = *reinterpret_cast<const int4*>(v.connS + 4 * k);
    }
#pragma unroll
    for (int o = 0; o < OCT; ++o) {
        const int64_t k0 = kw + 8 * o;
        if (k0 >= v.K) break;
        double2* ws2 = reinterpret_cast<double2*>(wtile + o * kWarpTile);
        const int nEl = (int)(v.K - k0 < 8 ? v.K - k0 : 8);
        // ---- own records of the octet: contiguous ranges ---------------------------------------------------------------------
        if ((lane >> 2) < nEl) cpAsync16(ws2 + kOwnCell / 2 + lane, v.cell + 8 * k0 + 2 * lane);
        if (lane < 6 * nEl) cpAsync16(ws2 + kOwnVtx / 2 + lane, v.vtx + 12 * k0 + 2 * lane);
        if (lane + 32 < 6 * nEl) cpAsync16(ws2 + kOwnVtx / 2 + lane + 32, v.vtx + 12 * k0 + 2 * (lane + 32));
        if (lane < 3 * nEl) cpAsync16(ws2 + kOwnVerts / 2 + lane, v.verts + 6 * k0 + 2 * lane);
    }
#pragma unroll
    for (int o = 0; o < OCT; ++o) {
        const int64_t k = kw + 8 * o + e;
        const bool valid = k < v.K;
        double2* ws2 = reinterpret_cast<double2*>(wtile + o * kWarpTile);
        // ---- neighbour records: round r = face r of every element, the element's four lanes fetch one 16-B chunk each ------------
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const unsigned code = ((unsigned)cn[o].w >> (8 * r)) & 0xffu;
            const int64_t nb = r == 0 ? cn[o].x : (r == 1 ? cn[o].y : cn[o].z);
            const bool boundary = (code & kLimGhost) || nb == k;
            if (valid && !boundary) {
                const int nf = (int)(code & kLimFaceMask);
                const bool rev = code & kLimRev;
                const int vS = rev ? (nf + 1) % 3 : nf, vE = rev ? nf : (nf + 1) % 3;      // the neighbour's trace in this element's direction
                const int key = (e & 1) | (r << 1);
                double2* rec = ws2 + kNb / 2 + (r * 8 + e) * 8;
                cpAsync16(rec + (j ^ key), v.cell + 8 * nb + 2 * j);
                cpAsync16(rec + ((4 + j) ^ key), v.vtx + 12 * nb + 4 * (j < 2 ? vS : vE) + 2 * (j & 1));
            }
        }
    }
    cpAsyncWaitAll();
    __syncwarp();
#pragma unroll 1
    for (int o = 0; o < OCT; ++o) {
        const int64_t k = kw + 8 * o + e;
        if (kw + 8 * o >= v.K) break;      // warp-uniform
        const bool valid = k < v.K;
        const double* ws = wtile + o * kWarpTile;
        const double2* ws2 = reinterpret_cast<const double2*>(ws);
        const int4 cno = o == 0 ? cn[0] : (o == 1 ? cn[OCT > 1 ? 1 : 0] : (o == 2 ? cn[OCT > 2 ? 2 : 0] : cn[OCT > 3 ? 3 : 0]));
        double V[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a2 = 0;
        if (valid && j < 3) {
            const unsigned code = ((unsigned)cno.w >> (8 * j)) & 0xffu;
            const int64_t nb = j == 0 ? cno.x : (j == 1 ? cno.y : cno.z);
            const bool boundary = (code & kLimGhost) || nb == k;
            if (boundary) {
                const int slot = limFaceGradientOwnSide(v, k, j, V, a2);
                if (slot >= 0) limStore(v.CV + 8 * (v.K + slot), 8, V);      // a ghost cell takes the gradient of its face (:606-637)
            } else {
                const int vE = j == 2 ? 0 : j + 1;      // the face runs from vertex j to vertex j + 1
                double S[4], E[4], oS[4], oE[4], ck[8], cn8[8];
                limLoad(ws + kOwnVtx + e * 12 + j * 4, 4, S);
                limLoad(ws + kOwnVtx + e * 12 + vE * 4, 4, E);
                limLoad(ws + kOwnCell + e * 8, 8, ck);
                const int key = (e & 1) | (j << 1);
                const double2* rec = ws2 + kNb / 2 + (j * 8 + e) * 8;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double2 t = rec[c ^ key];
                    cn8[2 * c] = t.x;
                    cn8[2 * c + 1] = t.y;
                }
                const double2 s0 = rec[4 ^ key], s1 = rec[5 ^ key], e0 = rec[6 ^ key], e1 = rec[7 ^ key];
                oS[0] = s0.x; oS[1] = s0.y; oS[2] = s1.x; oS[3] = s1.y;
                oE[0] = e0.x; oE[1] = e0.y; oE[2] = e1.x; oE[3] = e1.y;
                const double2 p0 = ws2[kOwnVerts / 2 + e * 3 + j], p1 = ws2[kOwnVerts / 2 + e * 3 + vE];
                a2 = ck[6] + cn8[6];
                limFaceGradientCompute(v, S, E, oS, oE, ck, cn8, p0.x, p0.y, p1.x, p1.y, V);
            }
        }
        // cellA2 = (A2_0 + A2_1) + A2_2 and s = ((t_0 + t_1) + t_2) with t_f = A2_f V_f / cellA2: limCellGradientFused's order
        const int base = lane & ~3;
        const double cellA2 = (shflD(a2, base) + shflD(a2, base + 1)) + shflD(a2, base + 2);
        const double iA = limRcp(valid ? cellA2 : 1.0);
        double s[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double t = a2 * V[c] * iA;
            s[c] = (shflD(t, base) + shflD(t, base + 1)) + shflD(t, base + 2);
        }
        // lane j stores the pair (2j, 2j+1): the four lanes of the element write its 64-B record together
        if (valid) *reinterpret_cast<double2*>(v.CV + 8 * k + 2 * j) = make_double2(j == 0 ? s[0] : j == 1 ? s[2] : j == 2 ? s[4] : s[6],
                                                                                    j == 0 ? s[1] : j == 1 ? s[3] : j == 2 ? s[5] : s[7]);
    }
}

template <int NT, int MB>
__global__ void __launch_bounds__(kLimThreads, MB) limReconstructKernel(const LimiterView v)
{
    __shared__ double sabc[3][kMaxNp];      // affine node map: x_i = a_i p0 + b_i p1 + c_i p2 (limNode)
    for (int i = threadIdx.x; i < kMaxNp; i += blockDim.x) {
        const bool in = i < v.Np;
        sabc[0][i] = in ? -(v.r[i] + v.s[i]) * 0.5 : 0.0;
        sabc[1][i] = in ? (v.r[i] + 1.0) * 0.5 : 0.0;
        sabc[2][i] = in ? (v.s[i] + 1.0) * 0.5 : 0.0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, e = lane >> 2, j = lane & 3;
    const unsigned long long pol = streamPolicy();
    const bool stream = v.streamPlanes != 0;
    const double igm1 = 1.0 / (v.gamma - 1.0);
    const int64_t nOct = (v.K + 7) >> 3, W = (int64_t)gridDim.x * (blockDim.x >> 5);
    int64_t oct = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (oct >= nOct) return;      // warp-uniform
    auto loadConn = [&](int64_t oct_) -> int4 {
        const int64_t k_ = oct_ * 8 + e;
        return *reinterpret_cast<const int4*>(v.connS + 4 * (k_ < v.K ? k_ : v.K - 1));
    };
    int4 cnw = loadConn(oct), cnwNext = cnw;
    // persistent grid (launchT): the block prologue is paid once, and the connectivity of the warp's next octet is in flight while this
    // one is processed (connectivity -> neighbour gradients are two dependent trips to memory)
    for (; oct < nOct; oct += W) {
    if (oct + W < nOct) cnwNext = loadConn(oct + W);
    const int64_t k = oct * 8 + e;
    const bool valid = k < v.K;
    const int64_t kk = valid ? k : v.K - 1;
    // lane j computes the limited gradient of primitive j (its 16-B slice of the three neighbour records); the four lanes exchange them
    int64_t cn[3];      // limNeighbourCells on the connectivity record already in registers
#pragma unroll
    for (int lf = 0; lf < 3; ++lf) {
        const unsigned code = ((unsigned)cnw.w >> (8 * lf)) & 0xffu;
        const int64_t nb = lf == 0 ? cnw.x : (lf == 1 ? cnw.y : cnw.z);
        const bool boundary = (code & kLimGhost) || nb == kk;
        const int slot = boundary ? v.bslot[3 * kk + lf] : -1;
        cn[lf] = slot >= 0 ? v.K + slot : nb;
    }
    double gx[3], gy[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double2 g = *reinterpret_cast<const double2*>(v.CV + 8 * cn[i] + 2 * j);
        gx[i] = g.x;
        gy[i] = g.y;
    }
    double c[8];
    limCellConstants(v, kk, c);
    const double* p = v.verts + 6 * kk;
    const double p0x = p[0], p0y = p[1], p1x = p[2], p1y = p[3], p2x = p[4], p2y = p[5];
    double Lx, Ly;
    limLimitedGradientField(v, gx, gy, Lx, Ly);
    double L[8];
    const int base = lane & ~3;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        L[2 * f] = shflD(Lx, base + f);
        L[2 * f + 1] = shflD(Ly, base + f);
    }
    cnw = cnwNext;
    if (!valid) continue;      // (lanes of a ragged last octet; no warp-level operation follows in this iteration)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int n0 = nt * 8 + 2 * j;
        double o0[4] = {0, 0, 0, 0}, o1[4] = {0, 0, 0, 0};      // padding nodes stay zero
        if (n0 < v.Np)
            limReconstructAt(v, sabc[0][n0] * p0x + sabc[1][n0] * p1x + sabc[2][n0] * p2x, sabc[0][n0] * p0y + sabc[1][n0] * p1y + sabc[2][n0] * p2y, L, c, igm1, o0);
        if (n0 + 1 < v.Np)
            limReconstructAt(v, sabc[0][n0 + 1] * p0x + sabc[1][n0 + 1] * p1x + sabc[2][n0 + 1] * p2x,
                             sabc[0][n0 + 1] * p0y + sabc[1][n0 + 1] * p1y + sabc[2][n0 + 1] * p2y, L, c, igm1, o1);
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            if (stream) stStream(v.qout[f] + k * (NT * 8) + n0, make_double2(o0[f], o1[f]), pol);
            else *reinterpret_cast<double2*>(v.qout[f] + k * (NT * 8) + n0) = make_double2(o0[f], o1[f]);
        }
    }
    }
}

// resident blocks per SM of kernels B and C (register budget 64 / 85).  Measured at 500 k triangles, N=4 (ms per call, B with plain
// loads): B4 C4 0.239, B3 C4 0.244, B4 C3 0.220, B3 C3 0.225 - C spills at 64 registers and is a streaming writer, B lives on occupancy.
// HDG_LIM_CFG (A/B aid): bits 0 and 2 = octets per warp of B (see launchT), bit 1 = C at 4 blocks
int limConfig()
{
    static int cfg = -1;
    if (cfg < 0) {
        const char* e = std::getenv("HDG_LIM_CFG");
        cfg = e ? std::atoi(e) : 0;
    }
    return cfg;
}

// grid of kernels A and C: all octets, or (default) a persistent grid of the resident blocks - the block prologue is paid once and C
// keeps the connectivity of its next octet in flight (A 62.9 -> 60.9 us, C 79.4 -> 75.4 us at 500 k triangles).  HDG_LIM_CFG bit 3 =
// one block per 64 elements as in the first version
template <class Kernel>
unsigned limGrid(Kernel kern, int smem, int64_t blocksNeeded)
{
    if (limConfig() & 8) return (unsigned)blocksNeeded;
    int dev = 0, sms = 0, blocks = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, kLimThreads, smem);
    if (blocks < 1 || sms < 1) return (unsigned)blocksNeeded;
    return (unsigned)std::min<int64_t>(blocksNeeded, (int64_t)blocks * sms);
}

template <int MB, int OCT>
void launchGradients(const LimiterView& v, cudaStream_t st)
{
    constexpr int smem = (kLimThreads / 32) * OCT * kWarpTile * (int)sizeof(double);
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncSetAttribute(limGradientsKernel<MB, OCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured[dev & 63] = true;
    }
    // one block per 64 * OCT elements: as a persistent loop with the next item's connectivity in flight this kernel needs more than its
    // 64 registers and measured slower (67.6 vs 60.5 us at 500 k triangles)
    const int perBlock = (kLimThreads / 32) * 8 * OCT;
    limGradientsKernel<MB, OCT><<<(unsigned)((v.K + perBlock - 1) / perBlock), kLimThreads, smem, st>>>(v);
}

template <int NT>
void launchT(const LimiterView& v, unsigned grid, cudaStream_t st)
{
    const int cfg = limConfig();
    static unsigned gridA[64] = {}, gridC3[64] = {}, gridC4[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!gridA[dev & 63]) {
        gridA[dev & 63] = limGrid(limAveragesKernel<NT>, 0, (int64_t)1 << 30);
        gridC3[dev & 63] = limGrid(limReconstructKernel<NT, 3>, 0, (int64_t)1 << 30);
        gridC4[dev & 63] = limGrid(limReconstructKernel<NT, HDG_LIM_MB>, 0, (int64_t)1 << 30);
    }
    limAveragesKernel<NT><<<std::min(grid, gridA[dev & 63]), kLimThreads, 0, st>>>(v);
    switch (cfg & 5) {      // bit 0 / bit 2: octets per warp of kernel B (ms per call at 500 k triangles, N=4, C at 3 blocks)
        case 1: launchGradients<3, 2>(v, st); break;      // 0.209
        case 4: launchGradients<2, 3>(v, st); break;      // 0.213
        case 5: launchGradients<2, 4>(v, st); break;      // 0.242
        default: launchGradients<4, 1>(v, st); break;     // 0.203 (the same kernel with plain loads through registers: 0.218)
    }
    if (cfg & 2) limReconstructKernel<NT, HDG_LIM_MB><<<std::min(grid, gridC4[dev & 63]), kLimThreads, 0, st>>>(v);
    else limReconstructKernel<NT, 3><<<std::min(grid, gridC3[dev & 63]), kLimThreads, 0, st>>>(v);
}

}  // namespace

// returns the number of kernels launched
int launchTriangleLimiter(const LimiterView& v, cudaStream_t st)
{
    if (v.K <= 0) return 0;
    const int perBlock = (kLimThreads / 32) * 8;
    const unsigned grid = (unsigned)((v.K + perBlock - 1) / perBlock);
    switch (v.NpPad >> 3) {
        case 1: launchT<1>(v, grid, st); break;
        case 2: launchT<2>(v, grid, st); break;
        case 3: launchT<3>(v, grid, st); break;
        case 4: launchT<4>(v, grid, st); break;
        case 5: launchT<5>(v, grid, st); break;
        case 6: launchT<6>(v, grid, st); break;
        case 7: launchT<7>(v, grid, st); break;
        case 9: launchT<9>(v, grid, st); break;
        default: return 0;
    }
    return 3;
}

}  // namespace hdg
