// Which non-FP64 instructions issue "for free" between DFMAs?  Per 6 DFMAs (the ratio in the pair
// kernel: ~6 other instructions per 35 FP64) one instruction of a given kind is added.
// Baseline: 2 cycles per DFMA.  Fully exposed extra instruction: +1 cycle per 12 (+8.3 %).
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(256) kern(int iters, const double *in, double *sink, unsigned *isink)
{
    __shared__ double tab[256];
    tab[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double a[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) a[i] = in[threadIdx.x + i];
    unsigned x = threadIdx.x * 2654435761u, y = threadIdx.x + 7;
    double m = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int i = 0; i < 6; ++i) a[i] = __fma_rn(a[i], 0.999999, 1e-9);
            if (KIND == 1) x = x * 3u + y;                                  // IMAD
            if (KIND == 2) x = (x ^ y) & 0x3ff0u | y;                       // LOP3
            if (KIND == 3) x = (x << 4) + 0x40u;                            // shift/IADD family
            if (KIND == 4) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a[0])); m += r; } // MUFU (+1 DADD)
            if (KIND == 5) m += tab[(x + u) & 255];                         // LDS (+1 DADD, index ALU)
            if (KIND == 6) m += tab[u];                                     // LDS uniform address (+1 DADD)
            if (KIND == 7) m += 1.0;                                        // the extra DADD alone (reference for 4..6)
            if (KIND == 8) { x = x * 3u + y; y = (y ^ x) | 5u; }            // IMAD + LOP3 (2 ops)
        }
    }
    double s = m;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
    if (x + y == 0x12345678u) isink[0] = x;
}

template <int KIND>
void run(const char *name, int sms, const double *in, double *sink, unsigned *isink)
{
    const int iters = 1500, blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<KIND><<<blocks, 256>>>(iters / 4, in, sink, isink);
    cudaEventRecord(e0);
    kern<KIND><<<blocks, 256>>>(iters, in, sink, isink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %8.3f ms\n", name, ms);
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *sink, *in; unsigned *isink;
    cudaMalloc(&sink, 8); cudaMalloc(&in, 4096 * 8); cudaMalloc(&isink, 4);
    cudaMemset(in, 0, 4096 * 8);
    run<0>("6 DFMA", sms, in, sink, isink);
    run<1>("6 DFMA + IMAD", sms, in, sink, isink);
    run<2>("6 DFMA + LOP3", sms, in, sink, isink);
    run<3>("6 DFMA + shift/add", sms, in, sink, isink);
    run<8>("6 DFMA + IMAD + LOP3", sms, in, sink, isink);
    run<7>("6 DFMA + DADD", sms, in, sink, isink);
    run<4>("6 DFMA + DADD + MUFU.RSQ64H", sms, in, sink, isink);
    run<5>("6 DFMA + DADD + LDS (indexed)", sms, in, sink, isink);
    run<6>("6 DFMA + DADD + LDS (uniform)", sms, in, sink, isink);
    return 0;
}
