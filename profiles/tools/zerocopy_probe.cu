// Probe: SM-driven PCIe transfers through mapped pinned host memory vs copy engines (B200 box).  Build: nvcc -O3 -arch=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void k_copy(const double2* __restrict__ src, double2* __restrict__ dst, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
int main()
{
    const size_t bytes = 74181672 / 16 * 16, n2 = bytes / 16;
    double *h_in, *h_out, *d_a, *d_b;
    CK(cudaHostAlloc(&h_in, bytes, cudaHostAllocMapped));
    CK(cudaHostAlloc(&h_out, bytes, cudaHostAllocMapped));
    CK(cudaMalloc(&d_a, bytes));
    CK(cudaMalloc(&d_b, bytes));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks : {148, 296, 592, 1184}) {
        for (int mode = 0; mode < 3; mode++) { // 0 h2d, 1 d2h, 2 both
            float best = 1e9;
            for (int rep = 0; rep < 6; rep++) {
                CK(cudaDeviceSynchronize());
                cudaEventRecord(e0, s1);
                cudaStreamWaitEvent(s2, e0, 0);
                if (mode != 1) k_copy<<<blocks, 256, 0, s1>>>((const double2*)h_in, (double2*)d_a, n2);
                if (mode != 0) k_copy<<<blocks, 256, 0, s2>>>((const double2*)d_b, (double2*)h_out, n2);
                cudaEventRecord(e1, s2);
                cudaStreamWaitEvent(s1, e1, 0);
                cudaEventRecord(e1, s1);
                CK(cudaDeviceSynchronize());
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep && ms < best) best = ms;
            }
            printf("blocks %4d mode %s: %.3f ms  %.1f GB/s per direction\n", blocks, mode == 0 ? "h2d " : mode == 1 ? "d2h " : "both", best, bytes / best * 1e-6);
        }
    }
    // mixed: one direction on a copy engine, the other by a kernel; and chunked copy-engine duplex at two chunk sizes
    for (int mode = 3; mode < 8; mode++) {
        float best = 1e9;
        for (int rep = 0; rep < 6; rep++) {
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0, s1);
            cudaStreamWaitEvent(s2, e0, 0);
            if (mode == 3) { cudaMemcpyAsync(d_a, h_in, bytes, cudaMemcpyHostToDevice, s1); k_copy<<<296, 256, 0, s2>>>((const double2*)d_b, (double2*)h_out, n2); }
            if (mode == 4) { k_copy<<<296, 256, 0, s1>>>((const double2*)h_in, (double2*)d_a, n2); cudaMemcpyAsync(h_out, d_b, bytes, cudaMemcpyDeviceToHost, s2); }
            if (mode == 5) { cudaMemcpyAsync(d_a, h_in, bytes, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(h_out, d_b, bytes, cudaMemcpyDeviceToHost, s2); }
            if (mode == 6 || mode == 7) {
                const int k = mode == 6 ? 16 : 48;
                const size_t c = bytes / k / 16 * 16;
                for (int i = 0; i < k; i++) {
                    cudaMemcpyAsync((char*)d_a + i * c, (char*)h_in + i * c, c, cudaMemcpyHostToDevice, s1);
                    cudaMemcpyAsync((char*)h_out + i * c, (char*)d_b + i * c, c, cudaMemcpyDeviceToHost, s2);
                }
            }
            cudaEventRecord(e1, s2);
            cudaStreamWaitEvent(s1, e1, 0);
            cudaEventRecord(e1, s1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        const char* nm[] = {"", "", "", "h2d copy-engine + d2h kernel", "h2d kernel + d2h copy-engine", "both copy-engine, 1 piece", "both copy-engine, 16 pieces", "both copy-engine, 48 pieces"};
        printf("%s: %.3f ms  %.1f GB/s per direction\n", nm[mode], best, bytes / best * 1e-6);
    }
    return 0;
}
