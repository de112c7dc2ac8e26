// Path A: fused Huygens-Fresnel pair sum for sm_100a.
//
//   out[i] = sum_j (u_j ds_j) exp(-i k r_ij) / r_ij
//
// Replaces the ten CuPy elementwise launches + ZGEMV of GPU0402:112-121 (and the numba loop of
// CPU0402:71-85) with ONE kernel that never materialises the (batch x N_back) matrix:
//   * a small pack kernel fuses u*ds (CPU0402:102 / K6), pads the source set to whole tiles and
//     lays it out tile-contiguous [tile][sx|sy|sz|w_re|w_im][TILE] so that one TMA bulk copy
//     (cp.async.bulk -> SASS UBLKCP) brings a tile into shared memory;
//   * the pair kernel keeps DPT detector points + their complex accumulators in registers,
//     streams source tiles through a 3-stage mbarrier ring and evaluates every pair with
//     ~41 FP64-pipe instructions (no libdevice sincos: exact Cody-Waite reduction of k*r);
//   * bound: FP64 ALU (DFMA pipe).  HBM traffic is 40 B per source per detector-block, i.e.
//     ~0 B per pair; there is no dense contraction, so no tensor cores.
#include "akb_common.cuh"

namespace {

using namespace akb;

constexpr int TILE = 512;                    // source points per shared-memory stage
constexpr int ROWS = 5;                      // sx, sy, sz, w_re, w_im
constexpr int TILE_DOUBLES = ROWS * TILE;
constexpr int TILE_BYTES = TILE_DOUBLES * 8; // 20 KiB
constexpr int STAGES = 3;
constexpr int THREADS = 256;
constexpr int SMEM_BYTES = STAGES * TILE_BYTES + STAGES * 8;

struct PhaseConst {
    double k;    // FAITHFUL: phase = fl(k * r)
    double q_hi; // EXACT: quarter turns per metre, k*(2/pi) = q_hi + q_lo
    double q_lo;
};

// ---------------------------------------------------------------- pack
__global__ void pack_sources_kernel(const double *__restrict__ sx, const double *__restrict__ sy,
                                    const double *__restrict__ sz, const double *__restrict__ u,
                                    const double *__restrict__ ds, long long N, long long padded,
                                    double *__restrict__ packed)
{
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= padded) return;
    long long jj = j < N ? j : N - 1; // padding repeats the last point (finite r) with zero weight
    double wr = 0.0, wi = 0.0;
    if (j < N) {
        double d = ds ? ds[j] : 1.0;
        // u*ds exactly as NumPy's complex*real (CPU0402:102); the factor 2 (exact) pairs with
        // the 1/(2r) that the in-kernel square root produces for free.
        wr = mul(2.0, mul(u[2 * j], d));
        wi = mul(2.0, mul(u[2 * j + 1], d));
    }
    long long tile = j / TILE;
    int o = (int)(j % TILE);
    double *t = packed + tile * TILE_DOUBLES;
    t[0 * TILE + o] = sx[jj];
    t[1 * TILE + o] = sy[jj];
    t[2 * TILE + o] = sz[jj];
    t[3 * TILE + o] = wr;
    t[4 * TILE + o] = wi;
}

// ---------------------------------------------------------------- mbarrier / TMA bulk helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// ---------------------------------------------------------------- one (detector, source) pair
template <int MODE>
__device__ __forceinline__ void accumulate_pair(double X, double Y, double Z, double sx, double sy, double sz,
                                                double wr, double wi, const PhaseConst &pc, double &acc_re,
                                                double &acc_im)
{
    const double ddx = sub(X, sx), ddy = sub(Y, sy), ddz = sub(Z, sz);
    double s, root, hinv, f;
    int q;
    if (MODE == AKB_PHASE_FAITHFUL) {
        // CPU0402:76-80: (dx*dx + dy*dy) + dz*dz, one rounding per operation
        s = add(add(mul(ddx, ddx), mul(ddy, ddy)), mul(ddz, ddz));
        sqrt_and_half_rinv(s, root, hinv);
        const double p = mul(pc.k, root); // CPU0402:82: |phase| = fl(k*dist)
        reduce_pio2(p, q, f);
    } else {
        s = fma_(ddz, ddz, fma_(ddy, ddy, mul(ddx, ddx)));
        sqrt_and_half_rinv(s, root, hinv);
        // quarter turns: n = rint(r*q), f = r*q - n without ever rounding k*r
        const double t = fma_(root, pc.q_hi, AKB_RND_MAGIC);
        q = __double2loint(t);
        const double n = sub(t, AKB_RND_MAGIC);
        f = fma_(root, pc.q_hi, -n);
        f = fma_(root, pc.q_lo, f);
        f = mul(f, AKB_PIO2_HI);
    }
    double cf, sf, c, sn;
    scaled_sincos_kernel(f, hinv, cf, sf); // (cos f, sin f) / (2r)
    apply_quadrant(q, cf, sf, c, sn);
    // (wr + i wi) * (c - i sn)            [exp(-i k r)/r, CPU0402:81-84]
    acc_re = fma_(wr, c, acc_re);
    acc_re = fma_(wi, sn, acc_re);
    acc_im = fma_(wi, c, acc_im);
    acc_im = fma_(-wr, sn, acc_im);
}

// ---------------------------------------------------------------- pair kernel
// grid.x: blocks of THREADS*DPT detector points; grid.y: splits of the source tiles.
// out: [gridDim.y][M] complex partial sums (gridDim.y == 1 -> the result itself).
template <int DPT, int MODE>
__global__ void __launch_bounds__(THREADS) fresnel_pairs_kernel(
    const double *__restrict__ det_x, const double *__restrict__ det_y, const double *__restrict__ det_z,
    long long M, const double *__restrict__ packed, int tiles_total, int tiles_per_split, long long n_padded,
    PhaseConst pc, double *__restrict__ out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *tiles = reinterpret_cast<double *>(smem_raw);
    const uint32_t tiles_s = smem_u32(tiles);
    const uint32_t bars_s = smem_u32(smem_raw + STAGES * TILE_BYTES);

    const int t0 = blockIdx.y * tiles_per_split;
    const int t1 = min(t0 + tiles_per_split, tiles_total);
    const long long base = (long long)blockIdx.x * (THREADS * DPT) + threadIdx.x;

    double X[DPT], Y[DPT], Z[DPT], ar[DPT], ai[DPT];
#pragma unroll
    for (int d = 0; d < DPT; ++d) {
        long long i = base + (long long)d * THREADS;
        long long ic = i < M ? i : M - 1;
        X[d] = det_x[ic];
        Y[d] = det_y[ic];
        Z[d] = det_z[ic];
        ar[d] = 0.0;
        ai[d] = 0.0;
    }

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(bars_s + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES && t0 + s < t1; ++s) {
            mbar_expect_tx(bars_s + 8 * s, TILE_BYTES);
            tma_bulk_load(tiles_s + s * TILE_BYTES, packed + (long long)(t0 + s) * TILE_DOUBLES, TILE_BYTES,
                          bars_s + 8 * s);
        }
    }

    int stage = 0;
    uint32_t parity = 0;
    for (int t = t0; t < t1; ++t) {
        mbar_wait(bars_s + 8 * stage, parity);
        const double *T = tiles + stage * TILE_DOUBLES;
        const long long left = n_padded - (long long)t * TILE;
        const int cnt = left < TILE ? (int)left : TILE; // multiple of 2
#pragma unroll 1
        for (int j = 0; j < cnt; j += 2) {
            const double2 vx = *reinterpret_cast<const double2 *>(T + 0 * TILE + j);
            const double2 vy = *reinterpret_cast<const double2 *>(T + 1 * TILE + j);
            const double2 vz = *reinterpret_cast<const double2 *>(T + 2 * TILE + j);
            const double2 vr = *reinterpret_cast<const double2 *>(T + 3 * TILE + j);
            const double2 vi = *reinterpret_cast<const double2 *>(T + 4 * TILE + j);
#pragma unroll
            for (int d = 0; d < DPT; ++d) accumulate_pair<MODE>(X[d], Y[d], Z[d], vx.x, vy.x, vz.x, vr.x, vi.x, pc, ar[d], ai[d]);
#pragma unroll
            for (int d = 0; d < DPT; ++d) accumulate_pair<MODE>(X[d], Y[d], Z[d], vx.y, vy.y, vz.y, vr.y, vi.y, pc, ar[d], ai[d]);
        }
        __syncthreads(); // every thread is done with this stage
        if (threadIdx.x == 0 && t + STAGES < t1) {
            mbar_expect_tx(bars_s + 8 * stage, TILE_BYTES);
            tma_bulk_load(tiles_s + stage * TILE_BYTES, packed + (long long)(t + STAGES) * TILE_DOUBLES, TILE_BYTES,
                          bars_s + 8 * stage);
        }
        if (++stage == STAGES) {
            stage = 0;
            parity ^= 1;
        }
    }

    double2 *o = reinterpret_cast<double2 *>(out) + (long long)blockIdx.y * M;
#pragma unroll
    for (int d = 0; d < DPT; ++d) {
        long long i = base + (long long)d * THREADS;
        if (i < M) o[i] = make_double2(ar[d], ai[d]);
    }
}

// deterministic reduction of the per-split partial sums (fixed order)
__global__ void reduce_partials_kernel(const double2 *__restrict__ part, int splits, long long M,
                                       double2 *__restrict__ out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    double re = 0.0, im = 0.0;
    for (int s = 0; s < splits; ++s) {
        double2 v = part[(long long)s * M + i];
        re += v.x;
        im += v.y;
    }
    out[i] = make_double2(re, im);
}

__global__ void fill_zero_kernel(double *p, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0.0;
}

// k*(2/pi) as an unevaluated sum hi+lo (double-double product, ~2^-100 relative)
void quarter_turns_per_metre(double k, double &hi, double &lo)
{
    const double t_hi = 6.36619772367581382433e-01;  // 2/pi
    const double t_lo = -3.935735335036497e-17;      // 2/pi - t_hi
    double ph = k * t_hi;
    double pl = __builtin_fma(k, t_hi, -ph) + k * t_lo;
    hi = ph + pl;
    lo = pl - (hi - ph);
}

template <int DPT, int MODE>
int launch_pairs(const double *dx, const double *dy, const double *dz, long long M, const double *packed,
                 int tiles_total, long long n_padded, PhaseConst pc, double *out, int splits, int tiles_per_split,
                 cudaStream_t st)
{
    auto kern = fresnel_pairs_kernel<DPT, MODE>;
    AKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    dim3 grid((unsigned)((M + THREADS * DPT - 1) / (THREADS * DPT)), (unsigned)splits);
    kern<<<grid, THREADS, SMEM_BYTES, st>>>(dx, dy, dz, M, packed, tiles_total, tiles_per_split, n_padded, pc, out);
    AKB_LAUNCH_CHECK();
    return AKB_OK;
}

template <int DPT, int MODE>
int resident_blocks_per_sm(int *out)
{
    auto kern = fresnel_pairs_kernel<DPT, MODE>;
    AKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    AKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, kern, THREADS, SMEM_BYTES));
    if (*out < 1) *out = 1;
    return AKB_OK;
}

constexpr int DPT_MAIN = 2;

// optional in-library timing of the last akb_fresnel_sum call of this thread (bench.py roofline)
struct Timing {
    bool enabled = false;
    bool valid = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr}; // call start, pairs start, pairs end, call end
    int splits = 0, per_sm = 0;
    long long blocks_x = 0;
};
thread_local Timing g_timing;

int timing_mark(int idx, cudaStream_t st)
{
    if (!g_timing.enabled) return AKB_OK;
    if (!g_timing.ev[idx]) AKB_CUDA(cudaEventCreate(&g_timing.ev[idx]));
    AKB_CUDA(cudaEventRecord(g_timing.ev[idx], st));
    return AKB_OK;
}

} // namespace

extern "C" int akb_fresnel_sum(const double *det_x, const double *det_y, const double *det_z, int64_t M,
                               const double *src_x, const double *src_y, const double *src_z,
                               const double *src_u, const double *src_ds, int64_t N, double k, double *out,
                               int mode, void *stream)
{
    AKB_REQUIRE(M >= 0 && N >= 0, "M and N must be non-negative");
    AKB_REQUIRE(mode == AKB_PHASE_FAITHFUL || mode == AKB_PHASE_EXACT, "mode must be AKB_PHASE_FAITHFUL or AKB_PHASE_EXACT");
    if (M == 0) return AKB_OK;
    AKB_REQUIRE(det_x && det_y && det_z && out, "detector/out pointers must not be NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N == 0) { // empty sum (np.sum of an empty array) = 0
        fill_zero_kernel<<<(unsigned)((2 * M + 255) / 256), 256, 0, st>>>(out, 2 * M);
        AKB_LAUNCH_CHECK();
        return AKB_OK;
    }
    AKB_REQUIRE(src_x && src_y && src_z && src_u, "source pointers must not be NULL");
    AKB_REQUIRE(k >= 0.0 && k < 3.0e13, "wave number k must be in [0, 3e13)");

    int device = 0;
    AKB_CUDA(cudaGetDevice(&device));
    tune_pool(device);
    const int sms = sm_count(device);

    const int tiles_total = (int)((N + TILE - 1) / TILE);
    const long long padded = (long long)tiles_total * TILE;
    const long long n_padded = (N + 1) & ~1LL;

    // ---- plan: split the source tiles so the grid fills whole waves
    int per_sm = 1;
    int rc = (mode == AKB_PHASE_FAITHFUL) ? resident_blocks_per_sm<DPT_MAIN, AKB_PHASE_FAITHFUL>(&per_sm)
                                          : resident_blocks_per_sm<DPT_MAIN, AKB_PHASE_EXACT>(&per_sm);
    if (rc) return rc;
    const long long slots = (long long)sms * per_sm;
    const long long blocks_x = (M + THREADS * DPT_MAIN - 1) / (THREADS * DPT_MAIN);
    int splits = 1;
    {
        double best = -1.0;
        const int max_splits = tiles_total < 64 ? tiles_total : 64;
        for (int s = 1; s <= max_splits; ++s) {
            const int tps = (tiles_total + s - 1) / s;
            const int s_eff = (tiles_total + tps - 1) / tps;
            if (s_eff != s) continue;
            const double waves = (double)(blocks_x * s) / (double)slots;
            const double full = waves <= 1.0 ? 1.0 : (double)(long long)(waves + 0.999999);
            double eff = waves / full;
            // the last split may hold fewer tiles: account for the imbalance
            eff *= (double)tiles_total / ((double)tps * s);
            if (eff > best + 0.02) {
                best = eff;
                splits = s;
            }
            if (eff >= 0.97) break;
        }
    }
    const int tiles_per_split = (tiles_total + splits - 1) / splits;

    g_timing.valid = false;
    g_timing.splits = splits;
    g_timing.per_sm = per_sm;
    g_timing.blocks_x = blocks_x;
    if ((rc = timing_mark(0, st))) return rc;

    double *packed = nullptr, *partial = nullptr;
    AKB_CUDA(cudaMallocAsync(&packed, (size_t)padded * ROWS * sizeof(double), st));
    if (splits > 1) AKB_CUDA(cudaMallocAsync(&partial, (size_t)splits * M * 2 * sizeof(double), st));

    pack_sources_kernel<<<(unsigned)((padded + 255) / 256), 256, 0, st>>>(src_x, src_y, src_z, src_u, src_ds, N,
                                                                          padded, packed);
    AKB_LAUNCH_CHECK();

    PhaseConst pc;
    pc.k = k;
    quarter_turns_per_metre(k, pc.q_hi, pc.q_lo);
    double *dst = splits > 1 ? partial : out;
    if ((rc = timing_mark(1, st))) return rc;
    rc = (mode == AKB_PHASE_FAITHFUL)
             ? launch_pairs<DPT_MAIN, AKB_PHASE_FAITHFUL>(det_x, det_y, det_z, M, packed, tiles_total, n_padded, pc,
                                                           dst, splits, tiles_per_split, st)
             : launch_pairs<DPT_MAIN, AKB_PHASE_EXACT>(det_x, det_y, det_z, M, packed, tiles_total, n_padded, pc, dst,
                                                        splits, tiles_per_split, st);
    if (rc) return rc;
    if ((rc = timing_mark(2, st))) return rc;
    if (splits > 1) {
        reduce_partials_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const double2 *>(partial), splits, M, reinterpret_cast<double2 *>(out));
        AKB_LAUNCH_CHECK();
        AKB_CUDA(cudaFreeAsync(partial, st));
    }
    AKB_CUDA(cudaFreeAsync(packed, st));
    if ((rc = timing_mark(3, st))) return rc;
    g_timing.valid = g_timing.enabled;
    return AKB_OK;
}

extern "C" int akb_fresnel_timing(int enable)
{
    g_timing.enabled = enable != 0;
    g_timing.valid = false;
    return AKB_OK;
}

extern "C" int akb_fresnel_last_timing(double *pairs_ms, double *total_ms, int *splits, int64_t *blocks_x,
                                       int *blocks_per_sm)
{
    AKB_REQUIRE(g_timing.valid, "no timed akb_fresnel_sum call on this thread (akb_fresnel_timing(1) first)");
    AKB_CUDA(cudaEventSynchronize(g_timing.ev[3]));
    float a = 0.f, b = 0.f;
    AKB_CUDA(cudaEventElapsedTime(&a, g_timing.ev[1], g_timing.ev[2]));
    AKB_CUDA(cudaEventElapsedTime(&b, g_timing.ev[0], g_timing.ev[3]));
    if (pairs_ms) *pairs_ms = a;
    if (total_ms) *total_ms = b;
    if (splits) *splits = g_timing.splits;
    if (blocks_x) *blocks_x = g_timing.blocks_x;
    if (blocks_per_sm) *blocks_per_sm = g_timing.per_sm;
    return AKB_OK;
}

extern "C" int akb_fresnel_sum_host(const double *det_x, const double *det_y, const double *det_z, int64_t M,
                                    const double *src_x, const double *src_y, const double *src_z,
                                    const double *src_u, const double *src_ds, int64_t N, double k, double *out,
                                    int mode, int device)
{
    AKB_REQUIRE(M >= 0 && N >= 0, "M and N must be non-negative");
    if (M == 0) return AKB_OK;
    AKB_REQUIRE(det_x && det_y && det_z && out, "detector/out pointers must not be NULL");
    AKB_REQUIRE(N == 0 || (src_x && src_y && src_z && src_u), "source pointers must not be NULL");
    AKB_CUDA(cudaSetDevice(device));
    tune_pool(device);
    cudaStream_t st;
    AKB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const size_t mb = (size_t)M * sizeof(double), nb = (size_t)(N > 0 ? N : 1) * sizeof(double);
    double *d = nullptr;
    // one slab: det xyz (3M) | out (2M) | src xyz (3N) | u (2N) | ds (N)
    const size_t total = 5 * mb + 6 * nb;
    int rc = AKB_OK;
    cudaError_t e = cudaMallocAsync(&d, total, st);
    if (e != cudaSuccess) {
        set_error("cudaMallocAsync(%zu) failed: %s", total, cudaGetErrorString(e));
        cudaStreamDestroy(st);
        return AKB_ERR_CUDA;
    }
    double *ddx = d, *ddy = d + M, *ddz = d + 2 * M, *dout = d + 3 * M;
    double *dsx = d + 5 * M, *dsy = dsx + (N > 0 ? N : 1), *dsz = dsy + (N > 0 ? N : 1), *du = dsz + (N > 0 ? N : 1);
    double *dds = du + 2 * (N > 0 ? N : 1);
#define H2D(dst, src, bytes)                                                                    \
    if (rc == AKB_OK && (bytes) > 0 && cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) { \
        set_error("H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));              \
        rc = AKB_ERR_CUDA;                                                                      \
    }
    H2D(ddx, det_x, mb) H2D(ddy, det_y, mb) H2D(ddz, det_z, mb)
    if (N > 0) {
        H2D(dsx, src_x, nb) H2D(dsy, src_y, nb) H2D(dsz, src_z, nb) H2D(du, src_u, 2 * nb)
        if (src_ds) { H2D(dds, src_ds, nb) }
    }
#undef H2D
    if (rc == AKB_OK)
        rc = akb_fresnel_sum(ddx, ddy, ddz, M, dsx, dsy, dsz, du, src_ds ? dds : nullptr, N, k, dout, mode, st);
    if (rc == AKB_OK && cudaMemcpyAsync(out, dout, 2 * mb, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
        set_error("D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = AKB_ERR_CUDA;
    }
    cudaFreeAsync(d, st);
    e = cudaStreamSynchronize(st);
    if (rc == AKB_OK && e != cudaSuccess) {
        set_error("stream synchronize failed: %s", cudaGetErrorString(e));
        rc = AKB_ERR_CUDA;
    }
    cudaStreamDestroy(st);
    return rc;
}

extern "C" int akb_shard_range(int64_t total, int nranks, int rank, int64_t *begin, int64_t *count)
{
    AKB_REQUIRE(total >= 0 && nranks > 0 && rank >= 0 && rank < nranks && begin && count, "bad shard arguments");
    const int64_t base = total / nranks, extra = total % nranks;
    *count = base + (rank < extra ? 1 : 0);
    *begin = rank * base + (rank < extra ? rank : extra);
    return AKB_OK;
}
