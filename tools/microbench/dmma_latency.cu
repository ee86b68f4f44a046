// DMMA.8x8x4 dependent-chain latency and throughput vs number of independent accumulator chains per warp (sm_100a).
// One warp per SMSP (4 warps per block, 1 block per SM) so that no other warp hides the latency.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void k(double* out, long long* cyc, double a, double b, int iters)
{
    double c[2 * CH];
#pragma unroll
    for (int i = 0; i < 2 * CH; ++i) c[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 2 * CH; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
void run(int warps)
{
    double* out; long long* cyc; cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<CH><<<148, warps * 32>>>(out, cyc, 1.0000001, 1e-9, iters);
    k<CH><<<148, warps * 32>>>(out, cyc, 1.0000001, 1e-9, iters);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM %2d chains/warp %d : %.1f cycles per DMMA per warp, %.1f cycles per DMMA per SMSP\n", warps, CH, (double)h / (iters * CH),
           (double)h / (iters * CH) / ((warps + 3) / 4));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int w : {4, 8, 16}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
