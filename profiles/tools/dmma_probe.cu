// dmma_probe.cu -- FP64 tensor-core (mma.sync.m8n8k4.f64, SASS DMMA) throughput on B200 against the DFMA pipe, and the same for the
// shape K3 would need: Ke += B^T (C B) per integration point is a chain of 24x6 * 6x6 * 6x24 products, i.e. m8n8k4 tiles with K = 6
// padded to 8 (25 % of the tensor work is padding) and operands that have to be re-laid from the thread-per-point layout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_probe profiles/tools/dmma_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a0, double b0)
{
    double c[CHAINS][2];
#pragma unroll
    for (int q = 0; q < CHAINS; q++) { c[q][0] = threadIdx.x + q; c[q][1] = q; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < CHAINS; q++) dmma(c[q], a, b);
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < CHAINS; q++) s += c[q][0] + c[q][1];
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;
}
template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a0, double b0)
{
    double c[CHAINS];
#pragma unroll
    for (int q = 0; q < CHAINS; q++) c[q] = threadIdx.x + q;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < CHAINS; q++) c[q] = fma(c[q], a0, b0);
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < CHAINS; q++) s += c[q];
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;
}

template <class K>
static float time_kernel(K k, double* out, int blocks, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<<<blocks, 256>>>(out, 16, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<<<blocks, 256>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    const int blocks = 148 * 8, iters = 8192;
    double* out;
    cudaMalloc(&out, (size_t)blocks * 256 * 8);
    const double warps = (double)blocks * 8;
    {
        const float ms = time_kernel(k_dmma<8>, out, blocks, iters);
        const double n = warps * iters * 8; // warp-level DMMA instructions
        printf("DMMA m8n8k4.f64, 8 chains/warp : %8.3f ms  %7.2f G warp-instr/s = %6.2f TFLOP/s (512 flop each)\n", ms, n / ms * 1e-6, n * 512 / ms * 1e-9);
    }
    {
        const float ms = time_kernel(k_dmma<4>, out, blocks, iters);
        const double n = warps * iters * 4;
        printf("DMMA m8n8k4.f64, 4 chains/warp : %8.3f ms  %7.2f G warp-instr/s = %6.2f TFLOP/s\n", ms, n / ms * 1e-6, n * 512 / ms * 1e-9);
    }
    {
        const float ms = time_kernel(k_dfma<8>, out, blocks, iters);
        const double n = warps * iters * 8;
        printf("DFMA, 8 chains/thread          : %8.3f ms  %7.2f G warp-instr/s = %6.2f TFLOP/s (64 flop each)\n", ms, n / ms * 1e-6, n * 64 / ms * 1e-9);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
