// `Triangle` slope limiter on the device: the five passes of dg_limiter_core.hpp, one thread per element (element-faces are the
// thread's three faces), launched back to back on the context's stream (each pass needs the previous one complete for all cells).
// HBM-bound gather/scatter work, small beside the stage (one read + one write of the four planes per call, plus O(K) work arrays).
// First version: plain kernels, each thread walks its own element row (NpPad doubles; the sectors of a row are reused through L1 over
// the node loop, the O(K) work arrays are coalesced) - no shared-memory staging yet.
//
// Parity: tests/test_gpu_limiter.py (device, through hdg_euler_limit) and tests/test_limiter_core_host.py (the same inline functions in
// host loops).
#include <cuda_runtime.h>

#include "dg_limiter_core.hpp"

namespace hdg {

namespace {

constexpr int kLimThreads = 128;

__global__ void __launch_bounds__(kLimThreads) limAveragesKernel(const LimiterView v)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < v.K) limCellAverages(v, k);
}

__global__ void __launch_bounds__(kLimThreads) limGhostKernel(const LimiterView v)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= v.K) return;
    for (int lf = 0; lf < 3; ++lf) limGhostCell(v, k, lf);
}

__global__ void __launch_bounds__(kLimThreads) limFaceGradKernel(const LimiterView v)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= v.K) return;
    for (int lf = 0; lf < 3; ++lf) limFaceGradient(v, k, lf);
}

__global__ void __launch_bounds__(kLimThreads) limCellGradKernel(const LimiterView v)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < v.K) limCellGradient(v, k);
}

__global__ void __launch_bounds__(kLimThreads) limReconstructKernel(const LimiterView v)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < v.K) limReconstruct(v, k);
}

// split form of pass 5 (v.L != nullptr): limited gradients per cell, then one thread per node slot so that the four planes are written
// with coalesced stores.  Opt-in (HDG_LIMITER_CFG=1) until it has been timed and run on a GPU; bit-identical on the host
// (tests/test_limiter_core_host.py::test_split_reconstruction_is_bit_identical)
__global__ void __launch_bounds__(kLimThreads) limStoreGradientKernel(const LimiterView v)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < v.K) limStoreGradient(v, k);
}

__global__ void __launch_bounds__(kLimThreads) limReconstructSlotKernel(const LimiterView v)
{
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    limReconstructSlot(v, slot);      // guards k < K and i < Np itself
}

}  // namespace

// returns the number of kernels launched
int launchTriangleLimiter(const LimiterView& v, cudaStream_t st)
{
    if (v.K <= 0) return 0;
    const unsigned grid = (unsigned)((v.K + kLimThreads - 1) / kLimThreads);
    limAveragesKernel<<<grid, kLimThreads, 0, st>>>(v);
    limGhostKernel<<<grid, kLimThreads, 0, st>>>(v);
    limFaceGradKernel<<<grid, kLimThreads, 0, st>>>(v);
    limCellGradKernel<<<grid, kLimThreads, 0, st>>>(v);
    if (!v.L) {
        limReconstructKernel<<<grid, kLimThreads, 0, st>>>(v);
        return 5;
    }
    limStoreGradientKernel<<<grid, kLimThreads, 0, st>>>(v);
    const int64_t slots = v.K * v.NpPad;
    limReconstructSlotKernel<<<(unsigned)((slots + kLimThreads - 1) / kLimThreads), kLimThreads, 0, st>>>(v);
    return 6;
}

}  // namespace hdg
