// Does a non-FP64 instruction issue "for free" next to DFMAs on sm_100a?
// Each variant runs 8 independent DFMA chains per thread plus K independent integer/FP32 ops
// per DFMA.  If the FP64 pipe (16 lanes/SMSP -> 2 cycles per warp instruction) leaves the
// issue port free every other cycle, time stays flat up to K = 1.
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int KIND>
__global__ void __launch_bounds__(256) kern(int iters, double seed, double *sink, unsigned *isink)
{
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x + i;
    unsigned x[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 7 + i; f[i] = threadIdx.x + i; }
    const double m = 0.999999, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = __fma_rn(a[i], m, c);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (KIND == 0) x[i] = (x[i] ^ (x[(i + 1) & 7] >> 3)) + 0x9e3779b9u; // ALU: LOP3/SHF/IADD
                    if (KIND == 1) f[i] = __fmaf_rn(f[i], 0.999f, 1e-3f);               // FMA pipe
                }
            }
        }
    }
    double s = 0; unsigned xs = 0; float fs = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += a[i]; xs ^= x[i]; fs += f[i]; }
    if (s == 123.456) sink[0] = s;
    if (xs == 0x12345678u || fs == 1.2345f) isink[0] = xs;
}

template <int K, int KIND>
void run(const char *name, int sms, double *sink, unsigned *isink)
{
    const int iters = 2000, blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<K, KIND><<<blocks, 256>>>(iters / 4, 1.0, sink, isink);
    cudaEventRecord(e0);
    kern<K, KIND><<<blocks, 256>>>(iters, 1.0, sink, isink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)blocks * 256 * iters * 64;
    printf("%-28s %8.3f ms  %6.2f TFLOP/s fp64\n", name, ms, 2 * fmas / ms / 1e9);
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *sink; unsigned *isink;
    cudaMalloc(&sink, 8); cudaMalloc(&isink, 4);
    run<0, 0>("dfma only", sms, sink, isink);
    run<1, 0>("dfma + 3 alu ops each", sms, sink, isink);
    run<2, 0>("dfma + 6 alu ops each", sms, sink, isink);
    run<1, 1>("dfma + 1 ffma each", sms, sink, isink);
    run<2, 1>("dfma + 2 ffma each", sms, sink, isink);
    run<4, 1>("dfma + 4 ffma each", sms, sink, isink);
    return 0;
}
