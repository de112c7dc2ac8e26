// Does the FP64 pipe lose cycles when the two 64-bit vector-register sources of a DFMA sit in the same
// register bank?  a[i] = fma(x[(i + OFF) % 32], c, a[i]) over 32 accumulators: the sources are one x and one a
// register pair; OFF shifts which x meets which a, i.e. the relation of their register numbers.  The SASS
// (cuobjdump -sass) shows the physical registers ptxas chose; compare the timings with the parity of
// (register number / 2) of the two sources.
#include <cstdio>
#include <cuda_runtime.h>

template <int OFF>
__global__ void __launch_bounds__(256) kern(int iters, double seed, double *sink)
{
    double a[32], x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        a[i] = seed + threadIdx.x + i;
        x[i] = 1.0 - 1e-9 * (threadIdx.x + 3 * i);
    }
    const double c = 0.999999;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = __fma_rn(x[(i + OFF) % 32], c, a[i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += a[i] + x[i];
    if (s == 123.456) sink[0] = s;
}

template <int OFF>
void run(int sms, double *sink)
{
    const int iters = 4000, blocks = sms * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<OFF><<<blocks, 256>>>(iters / 4, 1.0, sink);
    cudaEventRecord(e0);
    kern<OFF><<<blocks, 256>>>(iters, 1.0, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)blocks * 256 * iters * 128;
    printf("OFF %d: %8.3f ms  %6.2f TFLOP/s\n", OFF, ms, 2 * fmas / ms / 1e9);
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *sink; cudaMalloc(&sink, 8);
    run<0>(sms, sink); run<1>(sms, sink); run<2>(sms, sink); run<3>(sms, sink); run<4>(sms, sink); run<5>(sms, sink);
    return 0;
}
