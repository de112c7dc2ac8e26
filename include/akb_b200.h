/*
 * akb_b200.h -- C-ABI of the B200-native (sm_100a) replacement for the two data-parallel
 * hot paths of Kakekakechan/AKBRaytracing.
 *
 * The reference has no FFI: its boundary is a set of Python functions.
 * Every entry point below names the reference callable it stands in for (file:line into
 * the reference tree).  The Python mirror of those callables lives in
 * akbraytracing_b200/{wavecalc,raytrace,handoff}.py and binds this header with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; akb_last_error() (thread local)
 *     describes the last failure of the calling thread;
 *   - "_host" entry points take HOST pointers, do H2D + kernels + D2H on `device` (< 0: the
 *     calling thread's current device), are synchronous, and apply the reference's all-or-nothing NaN / normalisation semantics;
 *   - all other entry points take DEVICE pointers, are asynchronous on `stream`
 *     (a cudaStream_t passed as void*, NULL = legacy default stream) and never synchronise;
 *   - arrays are the reference's: float64 C-contiguous (3,N) (row 0 = x, row 1 = y,
 *     row 2 = z), float64[N], complex128[N] as interleaved (re,im) doubles;
 *   - buffers are caller-owned; scratch comes from the stream-ordered CUDA pool and is
 *     released on the same stream before the call returns;
 *   - there is no CPU fallback: without a CUDA device every call fails with AKB_ERR_CUDA.
 */
#ifndef AKB_B200_H
#define AKB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AKB_OK 0
#define AKB_ERR_ARG (-1)  /* bad argument (NULL pointer, negative size, ...) */
#define AKB_ERR_CUDA (-2) /* CUDA runtime error, see akb_last_error() */
#define AKB_ERR_NCCL (-3) /* NCCL missing or an NCCL call failed, see akb_last_error() */

/* phase arithmetic of akb_fresnel_sum */
#define AKB_PHASE_FAITHFUL 0 /* r and k*r rounded exactly like NumPy/numba (default) */
#define AKB_PHASE_EXACT 1    /* fused r^2; the product k*r is never rounded (reduced in quarter \
                                turns); r itself is still one rounded double */
#define AKB_PHASE_REFERENCED 2 /* optical path relative to a per-tile reference point: r is never \
                                  rounded as a whole, phases keep ~1e-8 rad even at k*r ~ 1e12 */

/* entries of the int[AKB_NFLAGS] status block written by the ray kernels */
#define AKB_FLAG_MISS 0      /* number of rays with not(D > 0)                  ER3D:31 */
#define AKB_FLAG_ZERO_NORM 1 /* bit 2k: zero normal at mirror k, bit 2k+1: zero reflected vector */
#define AKB_FLAG_MISS_MASK 2 /* bit k: mirror k of a chain saw a ray with not(D > 0) */
#define AKB_NFLAGS 4

/* skip_normalize bit masks (same bit layout as AKB_FLAG_ZERO_NORM) */
#define AKB_SKIP_NORMAL(k) (1u << (2 * (k)))
#define AKB_SKIP_REFLECT(k) (1u << (2 * (k) + 1))

#define AKB_MAX_MIRRORS 8

const char *akb_last_error(void);
int akb_version(void);
/* number of CUDA devices visible, <0 on error (replaces cp.cuda.runtime.getDeviceCount(), GPU0402:13) */
int akb_device_count(void);
/* Give the scratch memory cached for `device` (< 0: current device) back to the driver.  Synchronises the device.
 * Scratch lives in the device's default stream-ordered pool with a release threshold of 4 GiB. */
int akb_trim(int device);

/* ------------------------------------------------------------------ path A
 * Huygens-Fresnel pair sum
 *     out[i] = sum_j (u[j]*ds[j]) * exp(-1j*k*r_ij) / r_ij ,  r_ij = |det_i - src_j|
 * Replaces compute_u_parallel (Wavecalc_raytrace_fromData_CPU0402.py:71-85) together with the
 * ds pre-multiplication of forward_propagation_numpy_batch (CPU0402:102), i.e. the body of
 * forward_propagation_numpy_batch (CPU0402:87-124), forward_propagation_cupy_batch
 * (Wavecalc_raytrace_fromData_GPU0402.py:139-201) and one device's share of
 * forward_propagation_cupy_batch_multi_gpu (GPU0402:64-136; _multi.py:64-229).
 *   det_x/y/z  float64[M]     detector ("front") points
 *   src_x/y/z  float64[N]     source ("back") points
 *   src_u      complex128[N]  field on the back surface, (re,im) interleaved
 *   src_ds     float64[N] or NULL (= 1)
 *   out        complex128[M]
 *   mode       AKB_PHASE_FAITHFUL | AKB_PHASE_EXACT | AKB_PHASE_REFERENCED
 * Domain: |k| < 1e12 (k < 0 is evaluated as the conjugate problem, like the reference's exp(1j*(-k*dist)))
 * and |k|*r < 3.4e12 rad for every pair (the phase is reduced exactly as an integer
 * multiple of 2*pi/4096 below 2^51; the reference's largest stage, 146 m at 1.35 nm, is 6.8e11 rad);
 * beyond that the result is undefined.  r = 0 yields NaN/inf like the
 * reference.  The summation order over j differs from the reference's (tiles, fixed-order partial sums):
 * results are bit-reproducible run to run and agree with the reference to ~1e-13 relative L2.
 */
int akb_fresnel_sum(const double *det_x, const double *det_y, const double *det_z, int64_t M,
                    const double *src_x, const double *src_y, const double *src_z,
                    const double *src_u, const double *src_ds, int64_t N, double k,
                    double *out, int mode, void *stream);

int akb_fresnel_sum_host(const double *det_x, const double *det_y, const double *det_z, int64_t M,
                         const double *src_x, const double *src_y, const double *src_z,
                         const double *src_u, const double *src_ds, int64_t N, double k,
                         double *out, int mode, int device);

/* Through-focus stack (BASELINE config 5; the reference computes one plane per run of the Wavecalc script,
 * CPU0402:330-370): the field of one source surface on `planes` detector planes x = x_planes[p] that share the pixels
 * (det_y[i], det_z[i]), i < M.  out = complex128[planes][M].  One launch over the plane-major flat detector set: with
 * meshgrid-ordered pixels every block lies in one row of one plane and takes the planar-row loop (REFERENCED: with the row
 * expansion), bit-identical to akb_fresnel_sum plane by plane up to the order of the partial sums.  x_planes is a DEVICE
 * pointer like the other arrays. */
int akb_fresnel_sum_planes(const double *det_y, const double *det_z, int64_t M, const double *x_planes, int planes,
                           const double *src_x, const double *src_y, const double *src_z, const double *src_u,
                           const double *src_ds, int64_t N, double k, double *out, int mode, void *stream);

/* Contiguous block [begin, begin+count) of rank `rank` when `total` detector points are split
 * over `nranks` devices exactly like cp.array_split (GPU0402:77-79): the first total%nranks
 * blocks hold one extra point. */
int akb_shard_range(int64_t total, int nranks, int rank, int64_t *begin, int64_t *count);

/* ------------------------------------------------------------------ path A on several GPUs
 * forward_propagation_cupy_batch_multi_gpu (GPU0402:64-136; threaded twin GPU0402_multi.py:123-229 with
 * process_on_gpu :64-121), for ONE process or host thread per GPU: rank `rank` of `nranks` computes its
 * akb_shard_range block of the M detector points against the full source set, straight into its slot of
 * `out`, and the blocks are all-gathered in place over NCCL (the replacement of cp.concatenate, GPU0402:135;
 * ncclAllGather for equal blocks, one grouped ncclBroadcast per block otherwise).  On return (stream order)
 * every rank holds the full complex128[M] field.
 *   nccl_comm          an ncclComm_t (as void*) spanning the nranks devices: the caller's own, PyTorch's
 *                      (ProcessGroupNCCL._comm_ptr()), or one made with akb_nccl_comm_init; may be NULL when nranks == 1
 *   det_x/y/z          float64[M], the FULL detector arrays on every rank (the reference splits views of them)
 *   src_*              as akb_fresnel_sum; with broadcast_sources != 0 rank 0's source arrays are first replicated
 *                      into the other ranks' (caller-allocated) buffers with ncclBroadcast -- the reference keeps
 *                      the back surface on device 0 and reads it by peer access (GPU0402:36-38); src_ds must then be
 *                      NULL on every rank or on none (the broadcasts are collective)
 * Like every collective, the call must be made by all nranks ranks with the same M, N, mode and flags.
 * NCCL is bound at run time (the libnccl.so.2 already loaded in the process, else the default search path, else
 * $AKB_NCCL_LIB); without it the call fails with AKB_ERR_NCCL when nranks > 1. */
int akb_fresnel_sum_sharded(void *nccl_comm, int rank, int nranks, const double *det_x, const double *det_y,
                            const double *det_z, int64_t M, double *src_x, double *src_y, double *src_z,
                            double *src_u, double *src_ds, int64_t N, double k, double *out, int mode,
                            int broadcast_sources, void *stream);

/* In-place all-gather of akb_shard_range blocks: buf holds `total` items of `width` doubles each; on entry the
 * block of `rank` is valid, on return (stream order) all of them are. */
int akb_allgather_blocks(void *nccl_comm, int rank, int nranks, double *buf, int64_t total, int width, void *stream);

/* Communicator helpers for callers that have none: rank 0 calls akb_nccl_unique_id and hands the 128 bytes to
 * the other ranks by any means; every rank then calls akb_nccl_comm_init on its device (ncclCommInitRank). */
int akb_nccl_unique_id(void *id128);
int akb_nccl_comm_init(void **comm, int nranks, int rank, const void *id128);
int akb_nccl_comm_destroy(void *comm);

/* Measurement aid: with akb_fresnel_timing(1) every akb_fresnel_sum call of this thread records
 * CUDA events on its stream; akb_fresnel_last_timing() waits for the last call and returns the
 * duration of the pair kernel alone, of the whole call (pack + pairs + reduce), and the plan. */
int akb_fresnel_timing(int enable);
int akb_fresnel_last_timing(double *pairs_ms, double *total_ms, int *splits, int64_t *blocks_x,
                            int *blocks_per_sm);

/* Test / measurement aid: number of pair-kernel blocks on the current device that found their detector
 * points on a plane x = const with rows aligned to the threads ("planar-row" loop)
 * since the last reset.  Synchronises with the device. */
int akb_fresnel_row_blocks(int64_t *row_blocks, int reset);

/* name of the pair-kernel variant in use (default, or chosen with AKB_FRESNEL_VARIANT=<n>) */
const char *akb_fresnel_variant_name(void);

/* kernel launches issued by the calling thread since the last reset (bench.py "gpu_launches") */
int64_t akb_launch_count(int reset);

/* ------------------------------------------------------------------ path B
 * Ray / quadric-mirror kernels.  coeffs = the reference's 10 doubles [a..j] of
 * a x^2+b y^2+c z^2+d xy+e xz+f yz+g x+h y+i z+j = 0, built on the HOST by the caller
 * (never rebuilt on device: SURVEY.md H2).  coeffs pointers are HOST pointers everywhere.
 * flags = int[AKB_NFLAGS] in device memory, zeroed by the entry point on `stream`.
 */

/* mirr_ray_intersection(coeffs, ray, source, negative) -- EllipseRaytrace3D.py:18-45,
 * AKB_raytrace_20250312.py:445-471.  Rays with not(D>0) are counted in
 * flags[AKB_FLAG_MISS] (D<0 comes out NaN per ray, D==0 finite); the caller applies the
 * reference's whole-array NaN fill. */
int akb_mirr_ray_intersection(const double *coeffs, const double *ray, const double *source, int64_t N,
                              int negative, double *point, int *flags, void *stream);

/* norm_vector(coeffs, point) -- ER3D:61-71, BIG:626-636. */
int akb_norm_vector(const double *coeffs, const double *point, int64_t N, double *normal,
                    unsigned skip_normalize, int *flags, void *stream);

/* reflect_ray(ray, N) -- ER3D:47-55, BIG:502-509. */
int akb_reflect_ray(const double *ray, const double *normal, int64_t N, double *reflect,
                    unsigned skip_normalize, int *flags, void *stream);

/* normalize_vector(vector) -- ER3D:57-59, BIG:530-532 (out may alias vec). */
int akb_normalize_vector(const double *vec, int64_t N, double *out, unsigned skip_normalize, int *flags,
                         void *stream);

/* plane_ray_intersection(coeffs, ray, source) -- ER3D:145-157, BIG:873-885 (coeffs[6..9] used). */
int akb_plane_ray_intersection(const double *coeffs, const double *ray, const double *source, int64_t N,
                               double *point, void *stream);

/* ell.calc_reflect(inc_vector, inc_points) -- ER3D:241-245: intersect -> normal -> reflect in ONE
 * pass over HBM.  normal may be NULL (then 96 B/ray instead of 120 B/ray are moved). */
int akb_intersect_reflect(const double *coeffs, const double *ray, const double *source, int64_t N,
                          int negative, double *point, double *normal, double *reflect,
                          unsigned skip_normalize, int *flags, void *stream);

/* K-mirror chain + detector plane in one kernel: the call sequence of BIG:2881-2905 (AKB,
 * Wolter III+I: hyp_v, ell_v, ell_h, hyp_h with negative=1 on the 4th), BIG:11039-11054 (KB)
 * and AKB_raytrace_III_I_20250710.py:1586-1599.
 *   coeffs   double[K][10] (host), negative int[K] (host), plane double[10] (host) or NULL
 *   points   [K][3][N]            hit points P_k
 *   normals  [K][3][N] or NULL, reflects [K][3][N] or NULL (all reflected directions)
 *   last_reflect [3][N] or NULL   (direction after the last mirror)
 *   det      [3][N] or NULL       (plane_ray_intersection of the last ray)
 *   dist     [K][N] or NULL       segment lengths |P_k - P_{k-1}| (BIG:2884-2897), P_{-1} = source
 *   opl      [N] or NULL          optical path dist_0 + dist_1 + ... (+ |det - P_K| when a plane is given), summed
 *                                 left to right like totalDist (BIG:3621-3623)
 * HBM traffic per ray: 48 B in + 24 B per mirror + 24 (last_reflect) + 24 (det) + 8 K (dist) + 8 (opl) for the
 * outputs requested; two rays per thread, streaming loads/stores.
 */
int akb_trace_chain(const double *coeffs, const int *negative, int K, const double *plane,
                    const double *ray, const double *source, int64_t N, double *points, double *normals,
                    double *reflects, double *last_reflect, double *det, double *dist, double *opl,
                    unsigned skip_normalize, int *flags, void *stream);

/* The tail of plot_result_debug(p, 'ray_wave') in one pass (BIG:3516-3558, 3611-3631; III_I:1867-1971): the last
 * mirror's hit points and outgoing directions are rotated into the detector frame
 *     v' = R_y (R_z v),   P' = R_y (R_z (P - pivot)) + pivot       (rotate_vectors / rotate_points, BIG:917-944)
 * (rot_z = rot_y = NULL: no rotation), intersected with the plane x = plane_x and, when plane2_x != NULL, with
 * the defocused plane x = *plane2_x (plane_ray_intersection with coeffs [0,..,1,0,0,-x]), and the optical path
 *     opl = dist_0 + ... + dist_{K-1} + |det - P'|     (totalDist, BIG:3621-3623; opl2 with det2: totalDist2)
 * is formed per ray.  rot_z, rot_y: row-major 3x3 (host); pivot: double[3] (host); dist: [K][N] segment lengths of
 * akb_trace_chain (device; K = 0: only the last segment); every output may be NULL. */
int akb_wavefront_opl(const double *last_point, const double *last_dir, const double *dist, int K, int64_t N,
                      const double *rot_z, const double *rot_y, const double *pivot, double plane_x,
                      const double *plane2_x, double *point_rot, double *dir_rot, double *det, double *det2,
                      double *opl, double *opl2, void *stream);

/* B geometries (K quadrics + one plane each) trace the SAME ray bundle in one launch: the pattern of the
 * focus / alignment scans auto_focus_NA (BIG:12746-12895: ~1600 calls of the tracer with 53x53 rays,
 * each reduced to np.std of the detector y and z, BIG:12786-12787).
 *   coeffs [B][K][10] (host), negative int[K] (host, shared), planes [B][10] (host)
 *   ray, source [3][n] (device, shared by all geometries)
 *   det   [B][3][n] (device)   detector points per geometry
 *   stats [B][4] (device) or NULL: mean_y, std_y, mean_z, std_z of det (population std, like np.std)
 *   miss  int[B] (device): rays with not(D > 0) per geometry */
int akb_trace_chain_batched(const double *coeffs, const int *negative, int K, const double *planes, int B,
                            const double *ray, const double *source, int64_t n, double *det, double *stats,
                            int *miss, void *stream);

/* Host-buffer forms (H2D, kernel, D2H, reference NaN-fill / all-or-nothing normalisation applied).
 * host_flags (int[AKB_NFLAGS], may be NULL) receives the raw flags of the last pass. */
int akb_intersect_reflect_host(const double *coeffs, const double *ray, const double *source, int64_t N,
                               int negative, double *point, double *normal, double *reflect,
                               int *host_flags, int device);
int akb_trace_chain_host(const double *coeffs, const int *negative, int K, const double *plane,
                         const double *ray, const double *source, int64_t N, double *points,
                         double *normals, double *reflects, double *last_reflect, double *det,
                         double *dist, double *opl, int *host_flags, int device);

/* ------------------------------------------------------------------ hand-off helpers
 * calc_dS(points, ray_num_V, ray_num_H) -- AKB_raytrace_20250312.py:13418-13473. points (3,nV*nH). */
int akb_calc_ds(const double *points, int64_t nV, int64_t nH, double *dS, void *stream);

/* u[j] = amp[j] * exp(-1j*k*opl[j]) (amp may be NULL = 1): the field a traced wavefront carries onto
 * the last mirror, with the same exact phase reduction as akb_fresnel_sum. */
int akb_opl_to_field(const double *opl, const double *amp, int64_t N, double k, double *u, void *stream);

/* ------------------------------------------------------------------ probes (measurement only) */
/* Dependent-chain-free DFMA loop on every SM: achieved FP64 TFLOP/s (2 flop per DFMA). */
int akb_fp64_peak_probe(int iters, double *tflops, void *stream);
/* Device-to-device copy of nbytes, best of reps: achieved GB/s (read+write bytes). */
int akb_hbm_copy_probe(int64_t nbytes, int reps, double *gbs, void *stream);
/* Checks the in-kernel sqrt / reciprocal used by akb_fresnel_sum against __dsqrt_rn on n
 * pseudo-random inputs in [lo,hi): *mismatch = #inputs whose root differs from the correctly
 * rounded one, *max_rinv_rel = worst relative error of the reciprocal root. */
int akb_selftest_sqrt(int64_t n, double lo, double hi, int64_t *mismatch, double *max_rinv_rel, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AKB_B200_H */
