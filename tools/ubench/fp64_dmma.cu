// Is the FP64 tensor path (DMMA m8n8k4) a separate pipe from the FP64 vector pipe on sm_100a?
// Three kernels: DFMA only, DMMA only, and both interleaved (independent chains).  If DMMA ran
// beside DFMA, "mixed" would take max(dfma, dmma); if they share the FP64 units it takes the sum.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int NF, int NM> // NF DFMA and NM DMMA per inner step
__global__ void __launch_bounds__(256) kern(int iters, double seed, double *sink)
{
    double a[8], c0[4], c1[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) { c0[i] = seed * i; c1[i] = seed + i; }
    const double m = 0.999999, c = 1e-9;
    const double fa = 1e-3 * (threadIdx.x & 7), fb = 1e-3 * (threadIdx.x >> 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < NF; ++i) a[i & 7] = __fma_rn(a[i & 7], m, c);
#pragma unroll
            for (int i = 0; i < NM; ++i) dmma(c0[i & 3], c1[i & 3], fa, fb);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c0[i] + c1[i];
    if (s == 123.456) sink[0] = s;
}

template <int NF, int NM>
void run(const char *name, int sms, double *sink)
{
    const int iters = 2000, blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<NF, NM><<<blocks, 256>>>(iters / 4, 1.0, sink);
    cudaEventRecord(e0);
    kern<NF, NM><<<blocks, 256>>>(iters, 1.0, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double thr = (double)blocks * 256 * iters * 8;
    const double f_fma = thr * NF, f_mma = thr / 32 * NM * 256; // m8n8k4 = 256 FMA per warp instruction
    printf("%-24s %8.3f ms  dfma %6.2f TFLOP/s  dmma %6.2f TFLOP/s\n", name, ms, 2 * f_fma / ms / 1e9,
           2 * f_mma / ms / 1e9);
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *sink; cudaMalloc(&sink, 8);
    run<8, 0>("8 dfma", sms, sink);
    run<0, 1>("1 dmma", sms, sink);
    run<0, 4>("4 dmma", sms, sink);
    run<8, 1>("8 dfma + 1 dmma", sms, sink);
    run<8, 2>("8 dfma + 2 dmma", sms, sink);
    run<8, 4>("8 dfma + 4 dmma", sms, sink);
    return 0;
}
