// `Triangle` slope limiter on the device: the five passes of dg_limiter_core.hpp, one thread per element (element-faces are the
// thread's three faces), launched back to back on the context's stream (each pass needs the previous one complete for all cells).
// HBM-bound gather/scatter work, tiny beside the stage (one read + one write of the four planes per call, plus O(K) work arrays):
// plain coalesced kernels, no tensor-core or shared-memory staging.
//
// STATUS: compiled for sm_100a, arithmetic verified on the host through the same inline functions (tests/test_limiter_core_host.py);
// not yet run on a GPU (see the header of dg_limiter_core.hpp).
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
    limReconstructKernel<<<grid, kLimThreads, 0, st>>>(v);
    return 5;
}

}  // namespace hdg
