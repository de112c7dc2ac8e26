// Library-wide state: thread-local last error, launch counter, per-device attributes.
#include <stdarg.h>

#include <mutex>

#include "akb_common.cuh"

namespace akb {

static thread_local char g_error[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches += n; }

static std::mutex g_mu;
static int g_sms[64];
static bool g_pool_tuned[64];

int sm_count(int device)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (device < 0 || device >= 64) return 148;
    if (g_sms[device] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
        g_sms[device] = v;
    }
    return g_sms[device];
}

void tune_pool(int device)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (device < 0 || device >= 64 || g_pool_tuned[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        // Scratch (packed tiles, partial sums <= 2 GiB, host-API slabs) stays cached in the device's default
        // stream-ordered pool between calls, up to 4 GiB; anything above goes back to the driver at the next
        // synchronisation, and akb_trim() hands back all of it (the pool is shared with nobody inside this
        // library, but it is the process-wide default pool: a co-resident framework may want the memory).
        uint64_t keep = 4ull << 30;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    g_pool_tuned[device] = true;
}

// One non-blocking stream per (host thread, device) for the synchronous "_host" entry points:
// creating and destroying a stream per call costs more than a C1-sized problem takes to compute.
cudaStream_t host_stream(int device)
{
    static thread_local cudaStream_t streams[64] = {};
    if (device < 0 || device >= 64) return nullptr;
    if (!streams[device]) {
        if (cudaStreamCreateWithFlags(&streams[device], cudaStreamNonBlocking) != cudaSuccess) streams[device] = nullptr;
    }
    return streams[device];
}

} // namespace akb

extern "C" const char *akb_last_error(void) { return akb::g_error; }

extern "C" int akb_version(void) { return 200; } // 0.2.0

extern "C" int akb_trim(int device)
{
    akb::DeviceScope scope;
    device = scope.enter(device);
    if (device < 0) return AKB_ERR_CUDA;
    cudaMemPool_t pool;
    AKB_CUDA(cudaDeviceSynchronize());
    AKB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    AKB_CUDA(cudaMemPoolTrimTo(pool, 0));
    return AKB_OK;
}

extern "C" int akb_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        akb::set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
        return AKB_ERR_CUDA;
    }
    return n;
}

extern "C" int64_t akb_launch_count(int reset)
{
    int64_t v = akb::g_launches;
    if (reset) akb::g_launches = 0;
    return v;
}
