// Multi-GPU form of path A: one process (or host thread) per GPU, detector points sharded, NCCL for the plumbing.
//
// Replaces forward_propagation_cupy_batch_multi_gpu (Wavecalc_raytrace_fromData_GPU0402.py:64-136) and its
// threaded twin (Wavecalc_raytrace_fromData_GPU0402_multi.py:123-229, process_on_gpu :64-121):
//   * cp.array_split(x, num_gpus) (GPU0402:77-79)          -> akb_shard_range: contiguous blocks, the first
//                                                             M % nranks blocks hold one extra point;
//   * back-surface arrays read from device 0 by peer access -> optional ncclBroadcast of the source set from rank 0;
//   * per-device batches of K1-K5 + ZGEMV (GPU0402:105-125) -> ONE akb_fresnel_sum on the rank's block, written
//                                                             straight into its slot of the full output;
//   * cp.concatenate(results) on device 0 (GPU0402:135)     -> in-place all-gather of the blocks: ncclAllGather when
//                                                             the blocks are equal, else one grouped ncclBroadcast per
//                                                             rank (uneven tail).  Every rank ends up
//                                                             with the full field.
// There is no data-path exchange inside the pair kernel (detector points are independent), so the collective is
// a plain NVLink/NVSwitch all-gather after seconds of FP64 work (C4: 67 MB) -- nothing to overlap.
//
// NCCL is bound at run time: dlopen of the already loaded libnccl.so.2 first (a caller that passes an ncclComm_t
// made by PyTorch must reach the same library instance), then the default search path, or $AKB_NCCL_LIB.  The
// library therefore has no link-time dependency on NCCL and loads on a box without it.
#include <dlfcn.h>
#include <stdlib.h>

#include <mutex>
#include <type_traits>

#include "akb_common.cuh"

namespace {

using namespace akb;

// the handful of NCCL declarations used here (nccl.h: ncclResult_t is an enum, ncclSuccess = 0; ncclUniqueId is a
// 128-byte struct passed BY VALUE; ncclFloat64 = 8; ncclUint8 = 1)
struct NcclUniqueId {
    char internal[128];
};
constexpr int kNcclSuccess = 0;
constexpr int kNcclFloat64 = 8;

struct Nccl {
    void *handle = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*CommCount)(void *, int *) = nullptr;
    int (*CommUserRank)(void *, int *) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    char why[256] = "";
};

const Nccl *nccl()
{
    static Nccl api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = getenv("AKB_NCCL_LIB");
        void *h = nullptr;
        if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); // the instance already in the process
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            snprintf(api.why, sizeof(api.why), "libnccl.so.2 not found (%s); set AKB_NCCL_LIB", dlerror());
            return;
        }
        bool ok = true;
        auto bind = [&](auto &fn, const char *name) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(h, name));
            if (!fn) {
                snprintf(api.why, sizeof(api.why), "%s missing from the NCCL library", name);
                ok = false;
            }
        };
        bind(api.GetErrorString, "ncclGetErrorString");
        bind(api.GetUniqueId, "ncclGetUniqueId");
        bind(api.CommInitRank, "ncclCommInitRank");
        bind(api.CommDestroy, "ncclCommDestroy");
        bind(api.CommCount, "ncclCommCount");
        bind(api.CommUserRank, "ncclCommUserRank");
        bind(api.Broadcast, "ncclBroadcast");
        bind(api.AllGather, "ncclAllGather");
        bind(api.GroupStart, "ncclGroupStart");
        bind(api.GroupEnd, "ncclGroupEnd");
        if (ok) api.handle = h;
    });
    return &api;
}

#define AKB_NCCL(expr)                                                                              \
    do {                                                                                            \
        int _r = (expr);                                                                            \
        if (_r != kNcclSuccess) {                                                                   \
            akb::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, N->GetErrorString(_r)); \
            return AKB_ERR_NCCL;                                                                    \
        }                                                                                           \
    } while (0)

#define AKB_NEED_NCCL()                                       \
    const Nccl *N = nccl();                                   \
    if (!N->handle) {                                         \
        akb::set_error("NCCL unavailable: %s", N->why);       \
        return AKB_ERR_NCCL;                                  \
    }

} // namespace

extern "C" int akb_nccl_unique_id(void *id128)
{
    AKB_REQUIRE(id128, "id128 must point to 128 bytes");
    AKB_NEED_NCCL();
    AKB_NCCL(N->GetUniqueId(static_cast<NcclUniqueId *>(id128)));
    return AKB_OK;
}

extern "C" int akb_nccl_comm_init(void **comm, int nranks, int rank, const void *id128)
{
    AKB_REQUIRE(comm && id128 && nranks > 0 && rank >= 0 && rank < nranks, "bad communicator arguments");
    AKB_NEED_NCCL();
    NcclUniqueId id = *static_cast<const NcclUniqueId *>(id128);
    AKB_NCCL(N->CommInitRank(comm, nranks, id, rank)); // on the calling thread's current device
    return AKB_OK;
}

extern "C" int akb_nccl_comm_destroy(void *comm)
{
    if (!comm) return AKB_OK;
    AKB_NEED_NCCL();
    AKB_NCCL(N->CommDestroy(comm));
    return AKB_OK;
}

extern "C" int akb_allgather_blocks(void *nccl_comm, int rank, int nranks, double *buf, int64_t total, int width,
                                    void *stream)
{
    AKB_REQUIRE(nccl_comm && nranks > 0 && rank >= 0 && rank < nranks && total >= 0 && width > 0, "bad all-gather arguments");
    if (total == 0 || nranks == 1) return AKB_OK;
    AKB_REQUIRE(buf, "NULL buffer");
    AKB_NEED_NCCL();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int n = 0, r = -1;
    AKB_NCCL(N->CommCount(nccl_comm, &n));
    AKB_NCCL(N->CommUserRank(nccl_comm, &r));
    AKB_REQUIRE(n == nranks && r == rank, "rank / nranks do not match the communicator");
    if (total % nranks == 0) { // equal blocks: in-place all-gather (send buffer = this rank's slot of the receive buffer)
        const size_t cnt = (size_t)(total / nranks) * width;
        AKB_NCCL(N->AllGather(buf + (size_t)rank * cnt, buf, cnt, kNcclFloat64, nccl_comm, st));
        return AKB_OK;
    }
    AKB_NCCL(N->GroupStart()); // uneven tail: one in-place broadcast per block, fused into one group
    for (int q = 0; q < nranks; ++q) {
        int64_t b = 0, c = 0;
        akb_shard_range(total, nranks, q, &b, &c);
        if (c == 0) continue;
        double *p = buf + (size_t)b * width;
        int rc = N->Broadcast(p, p, (size_t)c * width, kNcclFloat64, q, nccl_comm, st);
        if (rc != kNcclSuccess) {
            N->GroupEnd();
            set_error("ncclBroadcast failed: %s", N->GetErrorString(rc));
            return AKB_ERR_NCCL;
        }
    }
    AKB_NCCL(N->GroupEnd());
    return AKB_OK;
}

extern "C" int akb_fresnel_sum_sharded(void *nccl_comm, int rank, int nranks, const double *det_x, const double *det_y,
                                       const double *det_z, int64_t M, double *src_x, double *src_y, double *src_z,
                                       double *src_u, double *src_ds, int64_t N_src, double k, double *out, int mode,
                                       int broadcast_sources, void *stream)
{
    AKB_REQUIRE(nranks > 0 && rank >= 0 && rank < nranks, "rank must be in [0, nranks)");
    AKB_REQUIRE(M >= 0 && N_src >= 0, "M and N must be non-negative");
    AKB_REQUIRE(nranks == 1 || nccl_comm, "an ncclComm_t is required for nranks > 1");
    if (M == 0) return AKB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (nranks > 1 && broadcast_sources && N_src > 0) {
        // rank 0 holds the back surface (the reference keeps it on device 0, GPU0402:36-38): replicate it once
        AKB_NEED_NCCL();
        AKB_REQUIRE(src_x && src_y && src_z && src_u, "source pointers must not be NULL");
        AKB_NCCL(N->GroupStart());
        int rc = kNcclSuccess;
        double *rows[4] = {src_x, src_y, src_z, src_ds};
        for (int q = 0; q < 4 && rc == kNcclSuccess; ++q)
            if (rows[q]) rc = N->Broadcast(rows[q], rows[q], (size_t)N_src, kNcclFloat64, 0, nccl_comm, st);
        if (rc == kNcclSuccess) rc = N->Broadcast(src_u, src_u, 2 * (size_t)N_src, kNcclFloat64, 0, nccl_comm, st);
        if (rc != kNcclSuccess) {
            N->GroupEnd();
            set_error("ncclBroadcast of the source set failed: %s", N->GetErrorString(rc));
            return AKB_ERR_NCCL;
        }
        AKB_NCCL(N->GroupEnd());
    }
    int64_t begin = 0, count = 0;
    int rc = akb_shard_range(M, nranks, rank, &begin, &count);
    if (rc) return rc;
    if (count > 0) {
        AKB_REQUIRE(det_x && det_y && det_z && out, "detector/out pointers must not be NULL");
        rc = akb_fresnel_sum(det_x + begin, det_y + begin, det_z + begin, count, src_x, src_y, src_z, src_u, src_ds, N_src,
                             k, out + 2 * begin, mode, stream);
        if (rc) return rc;
    }
    if (nranks == 1) return AKB_OK;
    return akb_allgather_blocks(nccl_comm, rank, nranks, out, M, 2, stream);
}
