// Shared helpers for the libakb_b200 translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/akb_b200.h"

namespace akb {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int sm_count(int device);
// make the stream-ordered pool of `device` keep its memory between calls (once per device)
void tune_pool(int device);
// cached non-blocking stream of the calling host thread on `device` (nullptr on failure)
cudaStream_t host_stream(int device);

// The "_host" entry points run on `device` but leave the caller's current device as they found it
// (a co-resident PyTorch / CuPy keeps its own notion of the current device).
struct DeviceScope {
    int prev = -1;
    bool changed = false;
    // device < 0: stay on the current device.  Returns the device in use, or < 0 on error (akb_last_error set).
    int enter(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            set_error("cudaGetDevice failed: %s", cudaGetErrorString(cudaGetLastError()));
            return -1;
        }
        if (device < 0 || device == prev) return prev;
        if (cudaSetDevice(device) != cudaSuccess) {
            set_error("cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(cudaGetLastError()));
            return -1;
        }
        changed = true;
        return device;
    }
    ~DeviceScope()
    {
        if (changed) cudaSetDevice(prev);
    }
};

#define AKB_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            akb::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return AKB_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define AKB_REQUIRE(cond, msg)                                   \
    do {                                                         \
        if (!(cond)) {                                           \
            akb::set_error("invalid argument: %s", msg);         \
            return AKB_ERR_ARG;                                  \
        }                                                        \
    } while (0)

#define AKB_LAUNCH_CHECK()                                                                    \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            akb::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return AKB_ERR_CUDA;                                                              \
        }                                                                                     \
        akb::count_launch();                                                                  \
    } while (0)

// ---- IEEE double arithmetic that the compiler may never contract into FMA.
// NumPy / numba evaluate the reference formulas one rounded operation at a time; these
// wrappers keep that order regardless of -fmad.
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }

// MUFU.RSQ64H: ~2^-21 relative approximation of 1/sqrt(x) (low word of the result is zero)
__device__ __forceinline__ double rsqrt_approx(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// sqrt(s) correctly rounded (up to a ~1e-9 chance of the neighbouring double) and
// half_rinv = 1/(2 sqrt(s)) to 1.3e-12 relative (measured), from ONE MUFU + 6 FP64 pipe ops:
// coupled Goldschmidt step on (g,h) = (s*y, y/2), then the Markstein residual correction.
__device__ __forceinline__ void sqrt_and_half_rinv(double s, double &root, double &half_rinv)
{
    double y = rsqrt_approx(s);
    double g = mul(s, y);
    // y/2 by decrementing the exponent field on the integer pipe (exact; the low word of the MUFU
    // result is zero): one FP64-pipe slot less per pair.  y = inf / 0 / NaN only arise together
    // with g = NaN (s = 0, inf, NaN), which poisons every later step exactly as before.
    double h = __hiloint2double(__double2hiint(y) - 0x00100000, 0);
    double e = fma_(-g, h, 0.5);
    g = fma_(g, e, g);
    h = fma_(h, e, h);
    double d = fma_(-g, g, s);
    root = fma_(d, h, g);
    half_rinv = h;
}

// The same without the residual correction: a second Goldschmidt step on g.  sqrt(s) to <= 1.3 ulp (0.33 ulp on average
// instead of 0.25; 77 % of the results are the correctly rounded ones) with the last instruction reading two registers
// instead of three -- for the modes that do not promise the reference's roundings.
__device__ __forceinline__ void sqrt_and_half_rinv_g2(double s, double &root, double &half_rinv)
{
    double y = rsqrt_approx(s);
    double g = mul(s, y);
    double h = __hiloint2double(__double2hiint(y) - 0x00100000, 0);
    double e = fma_(-g, h, 0.5);
    g = fma_(g, e, g);
    h = fma_(h, e, h);
    e = fma_(-g, h, 0.5);
    root = fma_(g, e, g);
    half_rinv = h;
}

// ---- sin/cos kernels on |f| <= pi/4 (fdlibm k_sin/k_cos minimax coefficients, < 2^-58)
#define AKB_S1 -1.66666666666666324348e-01
#define AKB_S2 8.33333333332248946124e-03
#define AKB_S3 -1.98412698298579493134e-04
#define AKB_S4 2.75573137070700676789e-06
#define AKB_S5 -2.50507602534068634195e-08
#define AKB_S6 1.58969099521155010221e-10
#define AKB_C1 4.16666666666666019037e-02
#define AKB_C2 -1.38888888888741095749e-03
#define AKB_C3 2.48015872894767294178e-05
#define AKB_C4 -2.75573143513906633035e-07
#define AKB_C5 2.08757232129817482790e-09
#define AKB_C6 -1.13596475577881948265e-11

#define AKB_TWO_OVER_PI 6.36619772367581382433e-01
#define AKB_PIO2_HI 1.57079632679489655800e+00 /* 0x3FF921FB54442D18 */
#define AKB_PIO2_LO 6.12323399573676603587e-17 /* 0x3C91A62633145C07 */
#define AKB_RND_MAGIC 6755399441055744.0       /* 1.5 * 2^52 */

// (scale*cos(f), scale*sin(f)) for |f| <= ~pi/4; 17 FP64 pipe ops including the scaling.
__device__ __forceinline__ void scaled_sincos_kernel(double f, double scale, double &sc, double &ss)
{
    double z = mul(f, f);
    double p = fma_(AKB_S6, z, AKB_S5);
    double q = fma_(AKB_C6, z, AKB_C5);
    p = fma_(p, z, AKB_S4);
    q = fma_(q, z, AKB_C4);
    p = fma_(p, z, AKB_S3);
    q = fma_(q, z, AKB_C3);
    p = fma_(p, z, AKB_S2);
    q = fma_(q, z, AKB_C2);
    p = fma_(p, z, AKB_S1);
    q = fma_(q, z, AKB_C1);
    q = fma_(q, z, -0.5);
    double hf = mul(scale, f);
    double hz = mul(scale, z);
    double hfz = mul(hf, z);
    ss = fma_(hfz, p, hf);
    sc = fma_(hz, q, scale);
}

// Rotate (cos f, sin f) by quadrant q (angle = q*pi/2 + f): returns cos/sin of the full angle.
__device__ __forceinline__ void apply_quadrant(int q, double cf, double sf, double &c, double &s)
{
    const bool swap = q & 1;
    double cc = swap ? sf : cf;
    double ssn = swap ? cf : sf;
    const int csign = ((q + 1) & 2) << 30; // quadrants 1,2: cos negative
    const int ssign = (q & 2) << 30;       // quadrants 2,3: sin negative
    c = __hiloint2double(__double2hiint(cc) ^ csign, __double2loint(cc));
    s = __hiloint2double(__double2hiint(ssn) ^ ssign, __double2loint(ssn));
}

// exact reduction of a non-negative double angle p (< 2^45) to quadrant + remainder
__device__ __forceinline__ void reduce_pio2(double p, int &q, double &f)
{
    double t = fma_(p, AKB_TWO_OVER_PI, AKB_RND_MAGIC);
    q = __double2loint(t);
    double n = sub(t, AKB_RND_MAGIC);
    f = fma_(n, -AKB_PIO2_HI, p); // exact: p - n*PIO2_HI fits in 53 bits
    f = fma_(n, -AKB_PIO2_LO, f);
}

} // namespace akb
