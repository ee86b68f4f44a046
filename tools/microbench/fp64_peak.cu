// FP64 pipe micro-benchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 f64)
// and DFMA+DMMA mixed, to decide which contraction engine the DG stage kernel uses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__);exit(1);}}while(0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double* out, double a, double b)
{
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* c, const double* a, double b)
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// 8 independent accumulator tiles per warp, m8n8k4: 256 FMA per instruction
__global__ void k_dmma884(double* out, double a, double b)
{
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma1684(double* out, double a, double b)
{
    double c[16]; double av[2] = {a, a + 1};
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) dmma1684(c + 4 * i, av, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma1688(double* out, double a, double b)
{
    double c[16]; double av[4] = {a, a + 1, a + 2, a + 3}; double bv[2] = {b, b + 1};
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) dmma1688(c + 4 * i, av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma16816(double* out, double a, double b)
{
    double c[16]; double av[8]; double bv[4];
#pragma unroll
    for (int i = 0; i < 8; i++) av[i] = a + i;
#pragma unroll
    for (int i = 0; i < 4; i++) bv[i] = b + i;
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) dmma16816(c + 4 * i, av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: per iteration 8 DMMA.884 (2048 FMA/warp) + 16 DFMA/thread (512 FMA/warp): do the pipes overlap?
__global__ void __launch_bounds__(1024) k_mixed(double* out, double a, double b)
{
    double c[16], d[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { c[i] = threadIdx.x + i; d[i] = threadIdx.x - i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            dmma884(c[2 * i], c[2 * i + 1], a, b);
            d[2 * i] = fma(d[2 * i], a, b);
            d[2 * i + 1] = fma(d[2 * i + 1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i] + d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
double timeit(K kern, int blocks, int threads, double* out)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; w++) kern<<<blocks, threads>>>(out, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0));
        kern<<<blocks, threads>>>(out, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best * 1e-3;
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 64 * 1024));
    for (int wpsm : {4, 8, 16, 32}) {   // warps per SM (1 block of wpsm*32 threads per SM x 2 blocks)
        int threads = wpsm * 32 > 1024 ? 1024 : wpsm * 32;
        int blocks = sms * (wpsm * 32 / threads) * 4;
        double nthreads = (double)blocks * threads, nwarps = nthreads / 32;
        double t;
        t = timeit(k_dfma, blocks, threads, out);
        printf("warps/SM-block %2d  DFMA      : %8.2f TFLOP/s\n", wpsm, 2.0 * nthreads * 16 * ITERS / t * 1e-12);
        t = timeit(k_dmma884, blocks, threads, out);
        printf("warps/SM-block %2d  DMMA884   : %8.2f TFLOP/s\n", wpsm, 2.0 * nwarps * 8 * 256 * ITERS / t * 1e-12);
        t = timeit(k_dmma1684, blocks, threads, out);
        printf("warps/SM-block %2d  DMMA1684  : %8.2f TFLOP/s\n", wpsm, 2.0 * nwarps * 4 * 512 * ITERS / t * 1e-12);
        t = timeit(k_dmma1688, blocks, threads, out);
        printf("warps/SM-block %2d  DMMA1688  : %8.2f TFLOP/s\n", wpsm, 2.0 * nwarps * 4 * 1024 * ITERS / t * 1e-12);
        t = timeit(k_dmma16816, blocks, threads, out);
        printf("warps/SM-block %2d  DMMA16816 : %8.2f TFLOP/s\n", wpsm, 2.0 * nwarps * 4 * 2048 * ITERS / t * 1e-12);
        t = timeit(k_mixed, blocks, threads, out);
        printf("warps/SM-block %2d  MIXED     : %8.2f TFLOP/s (2048 mma + 512 fma per warp-iter)\n", wpsm,
               2.0 * nwarps * (8 * 256 + 512) * ITERS / t * 1e-12);
    }
    return 0;
}
