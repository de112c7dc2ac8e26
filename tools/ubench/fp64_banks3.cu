// Three distinct 64-bit vector-register sources: a[i] = fma(x[(i + O1) % 32], y[(i + O2) % 32], a[i]).
// fp64_operands.cu measured 3 cycles instead of 2 for one register assignment; here the offsets shift the
// relation between the three register numbers, to see whether some combinations are conflict-free.
// Read the physical registers from cuobjdump -sass and compare with the timings.
#include <cstdio>
#include <cuda_runtime.h>

template <int O1, int O2>
__global__ void __launch_bounds__(256) kern(int iters, double seed, double *sink)
{
    double a[32], x[32], y[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        a[i] = seed + threadIdx.x + i;
        x[i] = 1.0 - 1e-9 * (threadIdx.x + 3 * i);
        y[i] = 1.0 - 1e-10 * (threadIdx.x + 5 * i);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = __fma_rn(x[(i + O1) % 32], y[(i + O2) % 32], a[i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += a[i] + x[i] + y[i];
    if (s == 123.456) sink[0] = s;
}

template <int O1, int O2>
void run(int sms, double *sink)
{
    const int iters = 4000, blocks = sms * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<O1, O2><<<blocks, 256>>>(iters / 4, 1.0, sink);
    cudaEventRecord(e0);
    kern<O1, O2><<<blocks, 256>>>(iters, 1.0, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)blocks * 256 * iters * 128;
    printf("O1 %d O2 %d: %8.3f ms  %6.2f TFLOP/s\n", O1, O2, ms, 2 * fmas / ms / 1e9);
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *sink; cudaMalloc(&sink, 8);
    run<0, 0>(sms, sink); run<0, 1>(sms, sink); run<1, 0>(sms, sink); run<1, 1>(sms, sink);
    run<0, 2>(sms, sink); run<2, 1>(sms, sink); run<1, 3>(sms, sink); run<2, 2>(sms, sink);
    return 0;
}
