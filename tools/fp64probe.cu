// tools/fp64probe.cu -- calibrates the compute floor of the sweep kernel's op phase on the GPU it runs on:
//   * DFMA throughput (vector FP64 pipe) for several ILP / occupancy points,
//   * DMMA throughput (mma.sync.m8n8k4.f64, the FP64 tensor-core path),
//   * both issued together (same warp, and from different warps), to see whether they are separate pipes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64probe fp64probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// MODE 0: DFMA only; 1: DMMA only; 2: both in every warp; 3: even warps DFMA, odd warps DMMA
template <int MODE, int ILP>
__global__ void k(double* out, double a, double b, int iters)
{
    double x[ILP], c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; i++)
    {
        x[i] = threadIdx.x * 1e-3 + i;
        c[i][0] = x[i];
        c[i][1] = -x[i];
    }
    const bool odd = (threadIdx.x >> 5) & 1;
    for (int it = 0; it < iters; it++)
    {
        if (MODE == 0 || MODE == 2 || (MODE == 3 && !odd))
        {
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
        }
        if (MODE == 1 || MODE == 2 || (MODE == 3 && odd))
        {
#pragma unroll
            for (int i = 0; i < ILP; i++) dmma(c[i][0], c[i][1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i] + c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int ILP>
void run(int threads, int blocks_per_sm)
{
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int grid = sms * blocks_per_sm, iters = 20000;
    double* out;
    cudaMalloc(&out, sizeof(double) * grid * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE, ILP><<<grid, threads>>>(out, 1.0000001, 1e-9, 100);
    cudaEventRecord(e0);
    k<MODE, ILP><<<grid, threads>>>(out, 1.0000001, 1e-9, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)grid * threads / 32;
    double fma_macs = 0, mma_macs = 0;
    if (MODE == 0 || MODE == 2) fma_macs = warps * 32.0 * iters * ILP;
    if (MODE == 1 || MODE == 2) mma_macs = warps * 256.0 * iters * ILP;
    if (MODE == 3) { fma_macs = warps / 2 * 32.0 * iters * ILP; mma_macs = warps / 2 * 256.0 * iters * ILP; }
    const double clk = khz * 1e3;
    printf("mode=%d ILP=%d threads/SM=%4d: %7.2f ms  DFMA %6.1f MAC/clk/SM  DMMA %6.1f MAC/clk/SM  total %6.2f TFLOP/s (at %d MHz nominal)\n",
           MODE, ILP, threads * blocks_per_sm, ms, fma_macs / (ms * 1e-3) / sms / clk, mma_macs / (ms * 1e-3) / sms / clk,
           2 * (fma_macs + mma_macs) / ms * 1e-9, khz / 1000);
    cudaFree(out);
}

int main()
{
    run<0, 1>(256, 2); run<0, 4>(256, 2); run<0, 8>(256, 2); run<0, 8>(256, 3); run<0, 8>(512, 4);
    run<1, 1>(256, 2); run<1, 4>(256, 2); run<1, 8>(256, 2); run<1, 8>(512, 4);
    run<2, 4>(256, 2); run<2, 8>(256, 2); run<2, 8>(512, 4);
    run<3, 4>(256, 2); run<3, 8>(256, 2); run<3, 8>(512, 4);
    return 0;
}
