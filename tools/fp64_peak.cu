// fp64 vector-pipe peak of the B200 (BASELINE.md asks the first build to measure it).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak tools/fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_kernel(double *out, double a, double b, int iters)
{
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 12345.678) out[0] = s;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 1 << 16, threads = 256, blocks = prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<8><<<blocks, threads>>>(out, 0.999999, 1e-9, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fma = (double)blocks * threads * 8 * iters;
        printf("%s: %d SMs, DFMA %.2f T FMA/s = %.2f TFLOP/s (2 flop per FMA), %.3f ms\n", prop.name,
               prop.multiProcessorCount, fma / ms / 1e9, 2 * fma / ms / 1e9, ms);
    }
    return 0;
}
