// Hand-off helpers between the ray path and the field path:
//   * akb_calc_ds     -- calc_dS (BIG:13418-13473): the O(N) Python double loop as one launch;
//   * akb_opl_to_field -- exp(-i k OPL) of a traced wavefront with the exact phase reduction.
#include "akb_common.cuh"

namespace {

using namespace akb;

struct P3 {
    double x, y, z;
};

__device__ __forceinline__ P3 at(const double *pts, long long N, long long idx)
{
    return {pts[idx], pts[N + idx], pts[2 * N + idx]};
}

// |(v1-v0) x (v2-v0)| / 2   (BIG:13446-13450)
__device__ __forceinline__ double tri_area(const P3 &v0, const P3 &v1, const P3 &v2)
{
    const double ax = sub(v1.x, v0.x), ay = sub(v1.y, v0.y), az = sub(v1.z, v0.z);
    const double bx = sub(v2.x, v0.x), by = sub(v2.y, v0.y), bz = sub(v2.z, v0.z);
    const double cx = sub(mul(ay, bz), mul(az, by));
    const double cy = sub(mul(az, bx), mul(ax, bz));
    const double cz = sub(mul(ax, by), mul(ay, bx));
    return mul(__dsqrt_rn(add(add(mul(cx, cx), mul(cy, cy)), mul(cz, cz))), 0.5);
}

// The reference fills edges by copying the nearest interior value and corners by copying the
// diagonal interior neighbour (BIG:13455-13471): in closed form every grid point takes the
// interior value at its (row, column) clamped into [1, n-2].
__global__ void calc_ds_kernel(const double *__restrict__ pts, long long nV, long long nH, double *__restrict__ dS)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long N = nV * nH;
    if (idx >= N) return;
    if (nV < 3 || nH < 3) { // no interior point: the reference leaves zeros
        dS[idx] = 0.0;
        return;
    }
    long long i = idx / nH, j = idx % nH;
    i = i < 1 ? 1 : (i > nV - 2 ? nV - 2 : i);
    j = j < 1 ? 1 : (j > nH - 2 ? nH - 2 : j);
    const long long c = i * nH + j;
    const P3 p = at(pts, N, c), pr = at(pts, N, c + 1), pl = at(pts, N, c - 1);
    const P3 pu = at(pts, N, c - nH), pd = at(pts, N, c + nH);
    double s = 0.0;
    s = add(s, tri_area(p, pr, pu));
    s = add(s, tri_area(p, pu, pl));
    s = add(s, tri_area(p, pl, pd));
    s = add(s, tri_area(p, pd, pr));
    dS[idx] = s;
}

__global__ void opl_to_field_kernel(const double *__restrict__ opl, const double *__restrict__ amp, long long N,
                                    double k, double2 *__restrict__ u)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double p = mul(k, opl[i]); // the angle NumPy would hand to exp(-1j*k*opl)
    const double a = amp ? amp[i] : 1.0;
    const double pa = fabs(p);
    int q;
    double f, cf, sf, c, s;
    reduce_pio2(pa, q, f);
    scaled_sincos_kernel(f, 1.0, cf, sf);
    apply_quadrant(q, cf, sf, c, s);
    if (p < 0.0) s = -s;
    u[i] = make_double2(mul(a, c), mul(a, -s));
}

} // namespace

extern "C" int akb_calc_ds(const double *points, int64_t nV, int64_t nH, double *dS, void *stream)
{
    AKB_REQUIRE(nV >= 0 && nH >= 0, "grid sizes must be non-negative");
    const long long N = (long long)nV * nH;
    if (N == 0) return AKB_OK;
    AKB_REQUIRE(points && dS, "NULL pointer");
    calc_ds_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(points, nV, nH, dS);
    AKB_LAUNCH_CHECK();
    return AKB_OK;
}

extern "C" int akb_opl_to_field(const double *opl, const double *amp, int64_t N, double k, double *u, void *stream)
{
    AKB_REQUIRE(N >= 0, "N must be non-negative");
    if (N == 0) return AKB_OK;
    AKB_REQUIRE(opl && u, "NULL pointer");
    opl_to_field_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        opl, amp, N, k, reinterpret_cast<double2 *>(u));
    AKB_LAUNCH_CHECK();
    return AKB_OK;
}
