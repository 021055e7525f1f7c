// FP64 FMA throughput of the device (the denominator of "FP64-pipe utilisation against peak"):
// every thread runs 8 independent DFMA chains; grid = 148 SMs x 8 CTAs x 256 threads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64peak fp64peak.cu && ./fp64peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b), x1 = fma(x1, a, b), x2 = fma(x2, a, b), x3 = fma(x3, a, b);
        x4 = fma(x4, a, b), x5 = fma(x5, a, b), x6 = fma(x6, a, b), x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 20000;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double fma = (double)blocks * threads * iters * 8;
    printf("SMs %d: %.3f ms, %.2f TFLOP/s FP64 (FMA = 2 flops), %.1f DFMA/clk/SM at 1.965 GHz\n", sms, best,
           2 * fma / (best * 1e-3) / 1e12, fma / (best * 1e-3) / sms / 1.965e9);
    return cudaGetLastError() != cudaSuccess;
}
