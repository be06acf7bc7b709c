// tools/fp64peak.cu -- measures the FP64 FMA throughput of the GPU (dependent chains of varying ILP), to calibrate
// the compute floor of the sweep kernel's op phase.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64peak fp64peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, double a, double b, int iters)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int threads, int blocks_per_sm)
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * blocks_per_sm, iters = 20000;
    double* out;
    cudaMalloc(&out, sizeof(double) * grid * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<grid, threads>>>(out, 1.0000001, 1e-9, 100);
    cudaEventRecord(e0);
    k<ILP><<<grid, threads>>>(out, 1.0000001, 1e-9, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)grid * threads * iters * ILP;
    printf("ILP=%d threads/SM=%d: %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", ILP, threads * blocks_per_sm,
           2 * fmas / ms * 1e-9, fmas / (ms * 1e-3) / sms / 1.965e9);
    cudaFree(out);
}
int main()
{
    run<1>(256, 2); run<2>(256, 2); run<4>(256, 2); run<8>(256, 2); run<8>(256, 3); run<8>(512, 4); run<16>(256, 2);
    return 0;
}
