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

#include "dg_limiter_core.hpp"

namespace hdg {

namespace {

#ifndef HDG_LIM_MB
#define HDG_LIM_MB 4
#endif
constexpr int kLimThreads = 256;      // 8 warps = 64 elements per block
constexpr int kMaxNp = 72;            // NpPad <= 72 (N <= 10)

__device__ __forceinline__ double shflD(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// A: NT = NpPad / 8 node tiles per element row, fully unrolled so that all 4*NT vector loads of a lane are in flight together
template <int NT>
__global__ void __launch_bounds__(kLimThreads) limAveragesKernel(const LimiterView v)
{
    __shared__ double sw[kMaxNp];      // mpp (0 beyond Np)
    __shared__ double sabc[4];         // sum_j w_j (a_j, b_j, c_j): the centroid is affine in the vertices (limNode); sum_j w_j
    for (int i = threadIdx.x; i < kMaxNp; i += blockDim.x) sw[i] = i < v.Np ? v.mpp[i] : 0.0;
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0, w = 0;
        for (int i = 0; i < v.Np; ++i) {
            const double wi = v.mpp[i];
            a += -(v.r[i] + v.s[i]) * 0.5 * wi;
            b += (v.r[i] + 1.0) * 0.5 * wi;
            c += (v.s[i] + 1.0) * 0.5 * wi;
            w += wi;
        }
        sabc[0] = a; sabc[1] = b; sabc[2] = c; sabc[3] = w;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, e = lane >> 2, j = lane & 3;
    const int64_t k = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8 + e;
    const bool valid = k < v.K;
    const int64_t kk = valid ? k : v.K - 1;
    double2 q[4][NT];
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) q[f][nt] = *reinterpret_cast<const double2*>(v.q[f] + kk * (NT * 8) + nt * 8 + 2 * j);
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

__global__ void __launch_bounds__(kLimThreads, HDG_LIM_MB) limGradientsKernel(const LimiterView v)
{
    const int lane = threadIdx.x & 31, e = lane >> 2, j = lane & 3;
    const int64_t k = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8 + e;
    const bool valid = k < v.K;
    double V[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a2 = 0;
    if (valid && j < 3) {
        const int slot = limFaceGradientOwnSide(v, k, j, V, a2);
        if (slot >= 0) limStore(v.CV + 8 * (v.K + slot), 8, V);      // a ghost cell takes the gradient of its face (:606-637)
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

template <int NT>
__global__ void __launch_bounds__(kLimThreads, HDG_LIM_MB) limReconstructKernel(const LimiterView v)
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
    const int64_t k = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8 + e;
    const bool valid = k < v.K;
    const int64_t kk = valid ? k : v.K - 1;
    // lane j computes the limited gradient of primitive j (its 16-B slice of the three neighbour records); the four lanes exchange them
    int64_t cn[3];
    limNeighbourCells(v, kk, cn);
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
    if (!valid) return;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int n0 = nt * 8 + 2 * j;
        double o0[4] = {0, 0, 0, 0}, o1[4] = {0, 0, 0, 0};      // padding nodes stay zero
        if (n0 < v.Np)
            limReconstructAt(v, sabc[0][n0] * p0x + sabc[1][n0] * p1x + sabc[2][n0] * p2x, sabc[0][n0] * p0y + sabc[1][n0] * p1y + sabc[2][n0] * p2y, L, c, o0);
        if (n0 + 1 < v.Np)
            limReconstructAt(v, sabc[0][n0 + 1] * p0x + sabc[1][n0 + 1] * p1x + sabc[2][n0 + 1] * p2x,
                             sabc[0][n0 + 1] * p0y + sabc[1][n0 + 1] * p1y + sabc[2][n0 + 1] * p2y, L, c, o1);
#pragma unroll
        for (int f = 0; f < 4; ++f) *reinterpret_cast<double2*>(v.qout[f] + k * (NT * 8) + n0) = make_double2(o0[f], o1[f]);
    }
}

template <int NT>
void launchT(const LimiterView& v, unsigned grid, cudaStream_t st)
{
    limAveragesKernel<NT><<<grid, kLimThreads, 0, st>>>(v);
    limGradientsKernel<<<grid, kLimThreads, 0, st>>>(v);
    limReconstructKernel<NT><<<grid, kLimThreads, 0, st>>>(v);
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
