// Measurement-only probes: the FP64 and HBM denominators the rooflines are quoted against,
// and a self-test of the in-kernel square root used by the Fresnel pair kernel.
#include "akb_common.cuh"

namespace {

using namespace akb;

// 8 independent DFMA chains per thread, fully unrolled: the FP64 pipe is the only limiter.
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double seed, double *sink)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fma_(a0, m, c); a1 = fma_(a1, m, c); a2 = fma_(a2, m, c); a3 = fma_(a3, m, c);
            a4 = fma_(a4, m, c); a5 = fma_(a5, m, c); a6 = fma_(a6, m, c); a7 = fma_(a7, m, c);
        }
    }
    double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) sink[0] = s; // never true: keeps the chains alive
}

__global__ void __launch_bounds__(256) copy_kernel(const int4 *__restrict__ src, int4 *__restrict__ dst, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[i];
}

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void sqrt_selftest_kernel(long long n, double lo, double hi, unsigned long long *mismatch,
                                     unsigned long long *max_err_bits)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double u = (double)(mix64((uint64_t)i) >> 11) * (1.0 / 9007199254740992.0);
    const double s = lo + (hi - lo) * u;
    double root, hinv;
    sqrt_and_half_rinv(s, root, hinv);
    const double ref = __dsqrt_rn(s);
    if (root != ref) atomicAdd(mismatch, 1ull);
    const double rel = fabs(2.0 * hinv * ref - 1.0);
    atomicMax(max_err_bits, (unsigned long long)__double_as_longlong(rel)); // rel >= 0: bit order == value order
}

} // namespace

extern "C" int akb_fp64_peak_probe(int iters, double *tflops, void *stream)
{
    AKB_REQUIRE(iters > 0 && tflops, "iters > 0 and tflops != NULL required");
    cudaStream_t st = (cudaStream_t)stream;
    int device = 0;
    AKB_CUDA(cudaGetDevice(&device));
    const int sms = sm_count(device);
    double *sink = nullptr;
    AKB_CUDA(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    AKB_CUDA(cudaEventCreate(&e0));
    AKB_CUDA(cudaEventCreate(&e1));
    const int blocks = sms * 8;
    dfma_peak_kernel<<<blocks, 256, 0, st>>>(iters / 8 + 1, 1.0, sink); // warm-up
    AKB_LAUNCH_CHECK();
    AKB_CUDA(cudaEventRecord(e0, st));
    dfma_peak_kernel<<<blocks, 256, 0, st>>>(iters, 1.0, sink);
    AKB_LAUNCH_CHECK();
    AKB_CUDA(cudaEventRecord(e1, st));
    AKB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    AKB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double fmas = (double)blocks * 256.0 * (double)iters * 16.0 * 8.0;
    *tflops = 2.0 * fmas / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return AKB_OK;
}

extern "C" int akb_hbm_copy_probe(int64_t nbytes, int reps, double *gbs, void *stream)
{
    AKB_REQUIRE(nbytes >= 16 && reps > 0 && gbs, "nbytes >= 16, reps > 0, gbs != NULL required");
    cudaStream_t st = (cudaStream_t)stream;
    int device = 0;
    AKB_CUDA(cudaGetDevice(&device));
    const long long n = nbytes / 16;
    int4 *a = nullptr, *b = nullptr;
    AKB_CUDA(cudaMalloc(&a, n * 16));
    AKB_CUDA(cudaMalloc(&b, n * 16));
    AKB_CUDA(cudaMemsetAsync(a, 1, n * 16, st));
    cudaEvent_t e0, e1;
    AKB_CUDA(cudaEventCreate(&e0));
    AKB_CUDA(cudaEventCreate(&e1));
    const int blocks = sm_count(device) * 16;
    double best = 0.0;
    for (int r = 0; r < reps + 1; ++r) {
        AKB_CUDA(cudaEventRecord(e0, st));
        copy_kernel<<<blocks, 256, 0, st>>>(a, b, n);
        AKB_LAUNCH_CHECK();
        AKB_CUDA(cudaEventRecord(e1, st));
        AKB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        AKB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double v = 2.0 * (double)n * 16.0 / (ms * 1e-3) / 1e9;
        if (r > 0 && v > best) best = v;
    }
    *gbs = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    return AKB_OK;
}

extern "C" int akb_selftest_sqrt(int64_t n, double lo, double hi, int64_t *mismatch, double *max_rinv_rel, void *stream)
{
    AKB_REQUIRE(n > 0 && lo > 0.0 && hi > lo && mismatch && max_rinv_rel, "bad selftest arguments");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *d = nullptr, h[2] = {0, 0};
    AKB_CUDA(cudaMalloc(&d, 2 * sizeof(unsigned long long)));
    AKB_CUDA(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), st));
    sqrt_selftest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, lo, hi, d, d + 1);
    AKB_LAUNCH_CHECK();
    AKB_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, st));
    AKB_CUDA(cudaStreamSynchronize(st));
    cudaFree(d);
    *mismatch = (int64_t)h[0];
    long long bits = (long long)h[1];
    double rel;
    memcpy(&rel, &bits, sizeof(rel));
    *max_rinv_rel = rel;
    return AKB_OK;
}
