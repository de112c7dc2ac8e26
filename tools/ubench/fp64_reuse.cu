// Does the operand reuse cache lift the 3-distinct-register DFMA penalty?
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(256) kern(int iters, const double *in, double *sink)
{
    double a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = in[threadIdx.x + i];
        b[i] = in[threadIdx.x + 8 + i];
        c[i] = in[threadIdx.x + 16 + i];
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (KIND == 0) a[i] = __fma_rn(b[0], c[i], a[i]);       // operand A shared by all 8 consecutive DFMAs
                if (KIND == 1) a[i] = __fma_rn(b[i >> 1], c[i], a[i]);  // operand A shared by pairs
                if (KIND == 2) a[i] = __fma_rn(b[i], c[i], a[i]);       // nothing shared
                if (KIND == 3) a[i] = __fma_rn(b[i >> 1], c[i >> 1], a[i]); // A and B shared by pairs
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
}

template <int KIND>
void run(const char *name, int sms, const double *in, double *sink)
{
    const int iters = 2000, blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<KIND><<<blocks, 256>>>(iters / 4, in, sink);
    cudaEventRecord(e0);
    kern<KIND><<<blocks, 256>>>(iters, in, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * 256 * iters * 64;
    printf("%-52s %8.3f ms  %6.2f Tinst-lanes/s (peak 18.6)\n", name, ms, ops / ms / 1e9);
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *sink, *in;
    cudaMalloc(&sink, 8); cudaMalloc(&in, 4096 * 8);
    cudaMemset(in, 0, 4096 * 8);
    run<2>("a[i] = b[i]*c[i] + a[i]   (nothing shared)", sms, in, sink);
    run<1>("a[i] = b[i/2]*c[i] + a[i] (A shared by pairs)", sms, in, sink);
    run<0>("a[i] = b[0]*c[i] + a[i]   (A shared by all)", sms, in, sink);
    run<3>("a[i] = b[i/2]*c[i/2]+a[i] (A,B shared by pairs)", sms, in, sink);
    return 0;
}
