// Accuracy of the reciprocal used on the column path (msed_column.cuh fast_rcp): MUFU.RCP64H seed,
// then either two Newton steps or one cubic step.  Prints the worst relative error of each against
// IEEE 1.0/x over a dense sweep of positive normal inputs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rcp_accuracy tools/rcp_accuracy.cu && ./rcp_accuracy
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__device__ double seed(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }

__global__ void sweep(double *worst, long long n)
{
    double w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        // mantissa sweep in [1,2) with irregular low bits, spread over 40 binades
        double m = 1.0 + (double)i / (double)n + 1.3e-13 * (double)(i % 7919);
        double x = ldexp(m, (int)(i % 40) - 20);
        double exact = 1.0 / x;
        double y = seed(x);
        w0 = fmax(w0, fabs(y - exact) / exact);
        double e = fma(-x, y, 1.0);
        double n1 = fma(y, e, y);
        double e2 = fma(-x, n1, 1.0);
        double n2 = fma(n1, e2, n1);                 // two Newton steps
        double c1 = fma(y, fma(e, e, e), y);         // one cubic step
        w1 = fmax(w1, fabs(n1 - exact) / exact);
        w2 = fmax(w2, fabs(n2 - exact) / exact);
        w3 = fmax(w3, fabs(c1 - exact) / exact);
    }
    // per-thread maxima -> global (values are non-negative doubles: integer max on the bits is exact)
    atomicMax((unsigned long long *)&worst[0], (unsigned long long)__double_as_longlong(w0));
    atomicMax((unsigned long long *)&worst[1], (unsigned long long)__double_as_longlong(w1));
    atomicMax((unsigned long long *)&worst[2], (unsigned long long)__double_as_longlong(w2));
    atomicMax((unsigned long long *)&worst[3], (unsigned long long)__double_as_longlong(w3));
}

int main()
{
    double *d, h[4];
    cudaMalloc(&d, 32);
    cudaMemset(d, 0, 32);
    sweep<<<148 * 8, 256>>>(d, 1LL << 32);
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("seed        : max rel err %.3e (2^%.1f)\n", h[0], log2(h[0]));
    printf("1 Newton    : max rel err %.3e (2^%.1f)\n", h[1], log2(h[1]));
    printf("2 Newton    : max rel err %.3e (2^%.1f)  = %.2f ulp\n", h[2], log2(h[2]), h[2] / 1.11e-16);
    printf("1 cubic step: max rel err %.3e (2^%.1f)  = %.2f ulp\n", h[3], log2(h[3]), h[3] / 1.11e-16);
    return 0;
}
