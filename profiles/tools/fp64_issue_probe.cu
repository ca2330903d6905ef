// fp64_issue_probe.cu -- does a DFMA hold the warp scheduler's issue port for both of its pipe cycles?  8 independent DFMA chains
// per thread, with K independent FFMA / IMAD / LDS instructions interleaved per DFMA.  If the time does not move with K the other
// pipes issue in the shadow of the FP64 pipe; if it grows by one cycle per extra instruction the port is shared.
#include <cuda_runtime.h>
#include <cstdio>
template <int K, int KIND, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) probe(double* out, float* outf, int iters, double a, float b)
{
    __shared__ float sm[1024];
    sm[threadIdx.x] = b;
    __syncthreads();
    double x[8];
    float y[8];
    int z[8];
#pragma unroll
    for (int q = 0; q < 8; q++) { x[q] = threadIdx.x * 1e-3 + q; y[q] = threadIdx.x * 1e-3f + q; z[q] = threadIdx.x + q; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            x[q] = fma(x[q], a, 1e-9);
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (KIND == 0) y[(q + k) & 7] = fmaf(y[(q + k) & 7], b, 1e-9f);
                if (KIND == 1) z[(q + k) & 7] = z[(q + k) & 7] * 3 + it;
                if (KIND == 2) y[(q + k) & 7] += sm[(threadIdx.x + z[(q + k) & 7] + k) & 1023];
            }
        }
    }
    double s = 0;
    float t = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) { s += x[q]; t += y[q] + z[q]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    outf[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int K, int KIND, int WARPS>
void run(const char* name, double* out, float* outf)
{
    const int iters = 4096, blocks = 148 * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<K, KIND, WARPS><<<blocks, WARPS * 32>>>(out, outf, 16, 1.0000001, 1.0000001f);
    cudaEventRecord(e0);
    probe<K, KIND, WARPS><<<blocks, WARPS * 32>>>(out, outf, iters, 1.0000001, 1.0000001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = (double)blocks * WARPS * 32 * iters * 8;
    printf("%-28s warps/CTA %2d (x4 CTAs/SM): %8.3f ms  %7.2f T DFMA/s = %6.2f TFLOP/s  (%.1f DFMA lanes/clk/SM at 1.965 GHz)\n", name, WARPS, ms, dfma / ms * 1e-9,
           2 * dfma / ms * 1e-9, dfma / (ms * 1e-3) / 148 / 1.965e9);
}
int main()
{
    double* out;
    float* outf;
    cudaMalloc(&out, 148 * 4 * 1024 * 8);
    cudaMalloc(&outf, 148 * 4 * 1024 * 4);
    run<0, 0, 4>("DFMA only", out, outf);
    run<0, 0, 2>("DFMA only", out, outf);
    run<0, 0, 1>("DFMA only", out, outf);
    run<1, 0, 4>("DFMA + 1 FFMA", out, outf);
    run<2, 0, 4>("DFMA + 2 FFMA", out, outf);
    run<1, 1, 4>("DFMA + 1 IMAD", out, outf);
    run<2, 1, 4>("DFMA + 2 IMAD", out, outf);
    run<1, 2, 4>("DFMA + 1 LDS", out, outf);
    run<1, 0, 1>("DFMA + 1 FFMA", out, outf);
    run<2, 0, 1>("DFMA + 2 FFMA", out, outf);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
