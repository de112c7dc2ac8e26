// FP64 pipe throughput versus the number of DISTINCT 64-bit register operands per instruction.
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
                if (KIND == 0) a[i] = __fma_rn(a[i], 0.999999, 1e-9);        // 1 register operand
                if (KIND == 1) a[i] = __fma_rn(a[i], b[i], a[i]);            // 2 distinct
                if (KIND == 2) a[i] = __fma_rn(a[i], b[i], c[i]);            // 3 distinct
                if (KIND == 3) a[i] = __fma_rn(b[i], c[(i + 1) & 7], a[i]);  // 3 distinct, no self-multiply
                if (KIND == 4) a[i] = __dmul_rn(a[i], b[i]);                 // DMUL 2 distinct
                if (KIND == 5) a[i] = __dadd_rn(a[i], b[i]);                 // DADD 2 distinct
                if (KIND == 6) a[i] = __fma_rn(a[i], b[i], 0.5);             // 2 distinct + imm
                if (KIND == 7) a[i] = __fma_rn(b[i], c[i], a[i]);            // 3 distinct, b/c same index
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
    printf("%-44s %8.3f ms  %6.2f Tinst-lanes/s (peak 18.6)\n", name, ms, ops / ms / 1e9);
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *sink, *in;
    cudaMalloc(&sink, 8); cudaMalloc(&in, 4096 * 8);
    cudaMemset(in, 0, 4096 * 8);
    run<0>("DFMA a = a*imm + imm (1 reg)", sms, in, sink);
    run<1>("DFMA a = a*b + a (2 distinct)", sms, in, sink);
    run<6>("DFMA a = a*b + 0.5 (2 distinct + imm)", sms, in, sink);
    run<2>("DFMA a = a*b + c (3 distinct)", sms, in, sink);
    run<3>("DFMA a = b*c' + a (3 distinct)", sms, in, sink);
    run<7>("DFMA a = b*c + a (3 distinct)", sms, in, sink);
    run<4>("DMUL a = a*b (2 distinct)", sms, in, sink);
    run<5>("DADD a = a+b (2 distinct)", sms, in, sink);
    return 0;
}
