// How much FP64-pipe time does a DFMA cost when it is issued among DMMA.8x8x4 (B200, sm_100a)?
// Pattern per warp and iteration: GD back-to-back DMMAs (8 independent accumulator tiles) followed by GF independent DFMAs, ratio and
// group size varied; time per iteration vs the pure-DMMA time gives the effective cycles per DFMA, for 1..4 warps per scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix_fp64 mix_fp64.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__);exit(1);}}while(0)
constexpr int ITERS = 2048;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int GD, int GF>
__global__ void __launch_bounds__(512) k_mix(double* out, double a, double b, long long* cyc)
{
    double c[16], d[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { c[i] = threadIdx.x + i; d[i] = threadIdx.x * 0.5 + i; }
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < GD; i++) dmma884(c[2 * (i & 7)], c[2 * (i & 7) + 1], a, b);
#pragma unroll
        for (int i = 0; i < GF; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i & 15]) : "d"(a), "d"(b));
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i] + d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int GD, int GF>
void run(int warpsPerSM, double* out, long long* dcyc)
{
    int dev; cudaDeviceProp pr; CK(cudaGetDevice(&dev)); CK(cudaGetDeviceProperties(&pr, dev));
    const int blocks = pr.multiProcessorCount * (warpsPerSM / 4);
    // one block of warpsPerSM warps per SM (launch bounds 1024 threads): every scheduler holds exactly warpsPerSM/4 warps
    const int threads = warpsPerSM * 32;
    (void)blocks;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_mix<GD, GF><<<pr.multiProcessorCount, threads>>>(out, 1.0000001, 1e-9, dcyc);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_mix<GD, GF><<<pr.multiProcessorCount, threads>>>(out, 1.0000001, 1e-9, dcyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    int khz; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    long long cyc; CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    const double perIter = (double)cyc / ITERS;                 // cycles per iteration of ONE warp (its scheduler is shared by warpsPerSM/4 warps)
    const double perSched = perIter / (warpsPerSM / 4);         // scheduler cycles per warp-iteration
    printf("[kernel %.3f ms = %.1f cycles/iter/warp at the nominal clock] ", ms, ms * 1e-3 * khz * 1e3 / ITERS);
    const double dfma = GF ? (perSched - 16.0 * GD) / GF : 0.0;
    printf("warps/SM %2d  GD %3d GF %3d : %8.1f cycles/iter/warp  %7.1f per scheduler  (pure DMMA %5d)  => %5.2f cycles per DFMA\n", warpsPerSM, GD, GF,
           perIter, perSched, 16 * GD, dfma);
}

int main()
{
    double* out; long long* dcyc;
    CK(cudaMalloc(&out, 148 * 8 * 512 * 8)); CK(cudaMalloc(&dcyc, 8));
    for (int w : {4, 12, 16}) {
        run<8, 0>(w, out, dcyc);
        run<0, 16>(w, out, dcyc);
        run<1, 1>(w, out, dcyc);
        run<2, 2>(w, out, dcyc);
        run<4, 4>(w, out, dcyc);
        run<8, 8>(w, out, dcyc);
        run<16, 16>(w, out, dcyc);
        run<32, 32>(w, out, dcyc);
        run<8, 2>(w, out, dcyc);
        run<32, 8>(w, out, dcyc);
        run<4, 6>(w, out, dcyc);
        run<24, 36>(w, out, dcyc);
    }
    return 0;
}
