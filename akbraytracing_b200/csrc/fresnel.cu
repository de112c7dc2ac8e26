// Path A: fused Huygens-Fresnel pair sum for sm_100a.
//
//   out[i] = sum_j (u_j ds_j) exp(-i k r_ij) / r_ij
//
// Replaces the ten CuPy elementwise launches + ZGEMV of GPU0402:112-121 (and the numba loop of
// CPU0402:71-85) with ONE kernel that never materialises the (batch x N_back) matrix:
//   * a small pack kernel fuses u*ds (CPU0402:102 / K6), turns the weight into polar form
//     (|w|, arg w split into whole table steps + remainder), pads the source set to whole tiles
//     and lays it out tile-contiguous [tile][sx|sy|sz|a|g|magic][TILE] so that one TMA bulk copy
//     (cp.async.bulk -> SASS UBLKCP) brings a tile into shared memory;
//   * the pair kernel keeps DPT detector points + their complex accumulators in registers and
//     streams source tiles through an mbarrier ring;
//   * per pair: r = sqrt(dx^2+dy^2+dz^2) from one MUFU.RSQ64H + 6 FP64 ops (correctly rounded,
//     also yields 1/(2r)), the phase k*r reduced EXACTLY (Cody-Waite with FMA) to a multiple of
//     2*pi/4096 plus a remainder, the weight phase folded into the table index and the remainder,
//     cos/tan of the remainder from two-term polynomials, the multiple looked up in a 4096-entry
//     (cos, sin) table in shared memory, a real multiply-accumulate: 31 FP64-pipe instructions
//     and ~8 others per pair, no libdevice sincos (its Payne-Hanek slow path is unusable at
//     k*r ~ 1e9);
//   * three phase modes: FAITHFUL (NumPy/numba rounding, the parity default), EXACT (k*r never
//     rounded), REFERENCED (optical path relative to a per-tile reference point);
//   * bound: FP64 ALU (DFMA pipe, 64 lanes/SM; the FP64 tensor path shares that pipe,
//     tools/ubench/fp64_dmma.cu).  HBM traffic is 48 B per source per detector block, i.e. ~0 B
//     per pair; there is no dense contraction, so no tensor cores.
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "akb_common.cuh"

namespace {

using namespace akb;

constexpr int HEAD = 4; // per-tile header after the rows: reference point c_T (x, y, z) + pad

// Formulation flags of a kernel variant (template parameter FORM).
//   FORM_TAN      rotate with tan f:  h e^{-i(theta+f)} = cf (C - S tan f) - i cf (S + C tan f),  cf = h cos f
//   FORM_SHORTCOS cos f = 1 - f^2/2  (only where |f| <= pi/2048: the f^4/24 term is <= 2.3e-13)
//   FORM_POLAR    sources carry |w|, and arg(w) is folded into the phase: the whole multiples of the
//                 table step go into the rounding constant (MAGIC - m_j, i.e. into the table index, free),
//                 the remainder is added to f (one DADD); the complex multiply-accumulate with w_j
//                 (4 DFMA) becomes a real one (2 DFMA).
//   FORM_WFOLD    (with FORM_TAN | FORM_POLAR | FORM_SHORTCOS) the weight rides in the cosine polynomial and the
//                 amplitude in the accumulation:  W = h (|w| - |w|/2 f^2)  (one DFMA with the per-source
//                 constants |w| and -|w|/2, one DMUL),  a = C - S tan f,  b = S + C tan f  (two DFMA on the
//                 raw table entry),  acc += W (a, -b)  (two DFMA): 9 instead of 11 FP64 instructions for
//                 polynomials + rotation + accumulation -- the table entry is never scaled component-wise.
//   FORM_E2       (REFERENCED, with FORM_WFOLD) an 8th row carries |e_j|^2 of the source's offset e_j from its tile's
//                 reference point, so that  r^2 - |D|^2 = |e|^2 + e . (-2D)  costs three DFMAs per pair (one on
//                 planar-row blocks) instead of three DADDs, a DMUL and two DFMAs.
//   FORM_SWZ      the table is stored XOR-swizzled, entry m at m ^ ((m >> 3) & 7): a warp whose lanes step through
//                 the table with an even stride (2, 4, 8 times an odd number) then hits eight different bank
//                 groups per quarter warp instead of 4, 2 or 1 (A/B variant).
//   FORM_ROWT     (REFERENCED, with FORM_E2, 4 points per thread) on planar-row blocks only the thread's FIRST point
//                 takes the square root; its three neighbours in the row, a few nanometres away, get their path
//                 difference from the expansion of r along the row,  r(y0 + D) - r(y0) = a D + b D^2  with
//                 a = (y0 - Y_j)/r,  b = (1 - a^2)/(2 r)  (third order: |a| D^3/(2 r^2), guarded below 1e-11 rad
//                 per block from the bounding box of the source set), and 1/(2r) to first order: 3 DFMAs instead
//                 of 10 instructions per pair for three pairs out of four.
//   FORM_PLANES   (FAITHFUL / EXACT, 4 points per thread) through-focus stacks: a thread owns ONE pixel (y, z) on FOUR
//                 planes x = x_p.  (y - Y_j)^2 and (z - Z_j)^2 are then one value per (thread, source), and (x_p - X_j)^2
//                 one value per (plane, source) that the block computes once per tile (row 0 in place of the sx row,
//                 three more rows in a scratch area): r^2 costs 2 additions + 1 shared instruction per pair instead
//                 of 4.5, with the reference's operations in the reference's order.  Launched by akb_fresnel_sum_planes
//                 only: det_x = the plane positions, M = pixels per plane, pc.planes = number of planes, grid.z = groups
//                 of four planes, out = [plane][pixel].
enum { FORM_TAN = 1, FORM_SHORTCOS = 2, FORM_POLAR = 4, FORM_WFOLD = 8, FORM_E2 = 16, FORM_SWZ = 32, FORM_ROWT = 64,
       FORM_PLANES = 128 };
// rows of a packed source tile: sx, sy, sz, then (w_re, w_im) or (|w|, -frac(arg w), MAGIC - m [, -|w|/2 [, |e|^2]])
__host__ __device__ constexpr int rows_of(int form)
{
    return (form & FORM_E2) ? 8 : (form & FORM_WFOLD) ? 7 : (form & FORM_POLAR) ? 6 : 5;
}
constexpr int MAX_ROWS = 8;

struct PhaseConst {
    double k;        // FAITHFUL: phase = fl(k * r)
    double inv_u;    // 1/u, u = 2*pi/TBL = reduction unit = angular step of the table
    double neg_u_hi; // -u as hi + lo
    double neg_u_lo;
    double u;        // EXACT: remainder in units -> radians
    double q_hi;     // EXACT: units per metre, k/u = q_hi + q_lo
    double q_lo;
    double im_sign;  // -1 for k < 0: exp(+i|k|r) = conj(exp(-i|k|r)), evaluated with conjugated weights
    // Copies of q_hi, q_lo, u for the once-per-(detector, tile) code of REFERENCED mode.  Without them ptxas keeps the
    // values in vector registers for that code and then feeds the pair loop's DFMAs from those registers: three
    // register reads per DFMA instead of two plus a uniform operand (+1.75 FP64-pipe cycles each, three per pair).
    double tq_hi, tq_lo, tu;
    int planes; // FORM_PLANES: number of detector planes (0 otherwise)
};

// order-preserving map double -> unsigned 64-bit (for atomicMin / atomicMax on coordinates) and back
__device__ __forceinline__ unsigned long long dkey(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k)
{
    return __longlong_as_double((long long)(k ^ ((k >> 63) ? 0x8000000000000000ull : ~0ull)));
}

// ---------------------------------------------------------------- pack
__global__ void pack_sources_kernel(const double *__restrict__ sx, const double *__restrict__ sy,
                                    const double *__restrict__ sz, const double *__restrict__ u,
                                    const double *__restrict__ ds, long long N, long long padded, int tile,
                                    int relative, int polar, int ROWS, double im_sign, double inv_u, double neg_u_hi, double neg_u_lo,
                                    double *__restrict__ packed, unsigned long long *__restrict__ bbox)
{
    if (bbox) { // bounding box of the source set as order-preserving integer keys: min x,y,z | max x,y,z (FORM_ROWT)
        const long long jb = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const long long jc = jb < N ? jb : N - 1;
        unsigned long long kx = dkey(sx[jc]), ky = dkey(sy[jc]), kz = dkey(sz[jc]);
        unsigned long long mn[3] = {kx, ky, kz}, mx[3] = {kx, ky, kz};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn[c], o), b = __shfl_xor_sync(0xffffffffu, mx[c], o);
                mn[c] = a < mn[c] ? a : mn[c];
                mx[c] = b > mx[c] ? b : mx[c];
            }
        if ((threadIdx.x & 31) == 0)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                atomicMin(&bbox[c], mn[c]);
                atomicMax(&bbox[3 + c], mx[c]);
            }
    }
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= padded) return;
    long long jj = j < N ? j : N - 1; // padding repeats the last point (finite r) with zero weight
    double wr = 0.0, wi = 0.0;
    if (j < N) {
        double d = ds ? ds[j] : 1.0;
        // u*ds exactly as NumPy's complex*real (CPU0402:102); the factor 2 (exact) pairs with
        // the 1/(2r) that the in-kernel square root produces for free.
        wr = mul(2.0, mul(u[2 * j], d));
        wi = mul(im_sign, mul(2.0, mul(u[2 * j + 1], d)));
    }
    long long t_idx = j / tile;
    int o = (int)(j % tile);
    double *t = packed + t_idx * (long long)(ROWS * tile + HEAD);
    // reference point of the tile = its first source; REFERENCED mode stores the sources relative to it
    // (differences of nearby points: exact or within 1e-18 m)
    const long long jc = t_idx * tile < N ? t_idx * tile : N - 1;
    const double cx = sx[jc], cy = sy[jc], cz = sz[jc];
    const double ex = relative ? sub(sx[jj], cx) : sx[jj];
    const double ey = relative ? sub(sy[jj], cy) : sy[jj];
    const double ez = relative ? sub(sz[jj], cz) : sz[jj];
    t[0 * tile + o] = ex;
    t[1 * tile + o] = ey;
    t[2 * tile + o] = ez;
    if (ROWS > 7) t[7 * tile + o] = relative ? fma_(ex, ex, fma_(ey, ey, mul(ez, ez))) : 0.0; // FORM_E2: |e|^2
    if (polar) {
        // w = |w| e^{i phi},  phi = m u + g  (u = table step, m integer, |g| <= u/2):
        //   w e^{-i p} = |w| e^{-i((p - g) - m u)}  ->  table index n - m, remainder f - g
        const double mag = hypot(wr, wi);
        const double phi = mag > 0.0 ? atan2(wi, wr) : 0.0;
        const double m = rint(mul(phi, inv_u));
        const double g = fma_(m, neg_u_lo, fma_(m, neg_u_hi, phi));
        t[3 * tile + o] = mag;
        t[4 * tile + o] = -g;
        t[5 * tile + o] = sub(AKB_RND_MAGIC, m);
        if (ROWS > 6) t[6 * tile + o] = mul(mag, -0.5); // FORM_WFOLD: the f^2 coefficient of |w| cos f
    } else {
        t[3 * tile + o] = wr;
        t[4 * tile + o] = wi;
    }
    if (o == 0) {
        t[ROWS * tile + 0] = cx;
        t[ROWS * tile + 1] = cy;
        t[ROWS * tile + 2] = cz;
        t[ROWS * tile + 3] = 0.0;
    }
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

// blocks that took the planar-row loop (test / measurement aid, akb_fresnel_row_blocks)
__device__ unsigned long long g_row_blocks;

// ---------------------------------------------------------------- constants
// FP64 constants live in constant memory so that they reach the DFMA as a c[bank][offset] /
// uniform-register operand: as literals ptxas re-materialises each of them with two UMOVs per
// loop iteration (30 extra issue slots per 4 pairs in the first version of this kernel).
enum { KC_MAGIC, KC_NEG_MAGIC, KC_T1, KC_T2, KC_TC1, KC_TT1, KC_TT2, KC_COUNT };
__constant__ double KC[KC_COUNT] = {
    AKB_RND_MAGIC, -AKB_RND_MAGIC,
    -1.0 / 6.0, 1.0 / 120.0, // sin f = f (1 + z (T1 + z T2)),  |f| <= pi/512: next term 1e-17 relative
    1.0 / 24.0,              // cos f = 1 + z (-1/2 + z TC1),              next term 8e-17
    1.0 / 3.0, 2.0 / 15.0};  // tan f = f (1 + z (TT1 + z TT2)),           next term 17/315 z^3

// ---------------------------------------------------------------- one (detector, source) pair, in three phases
//
// Cost model (measured on B200, tools/ubench/fp64_*.cu): the FP64 pipe takes one warp instruction
// every 2 cycles per SM sub-partition, but 3 cycles when the instruction reads THREE distinct
// 64-bit vector registers that the operand-reuse cache does not supply.  Constants (c[], uniform
// registers, immediates) are free.  Hence: polynomials are closed with the constant 1.0, the
// amplitude is applied by plain multiplications, and accumulations are ordered so that consecutive
// instructions share an operand.  Per pair (FAITHFUL, default variant): 31 FP64 instructions, ~3 of them 3-read.
//
// The 2*DPT pairs of one loop iteration go through the phases together, which puts the table load
// of a pair ~40 FP64 instructions ahead of its use.
struct PairA {
    double p; // FAITHFUL: fl(k*r); EXACT: r; REFERENCED: r - r_ref
    double h; // 1/(2r)
    double t; // MAGIC + rint(phase / u); its low word is the table index
};

// ---- REFERENCED mode: optical path relative to a per-tile reference point -------------------------
// r_ij is never formed as one double (its rounding alone is k*ulp(r)/2 = 7e-5 rad at 146 m, 1.35 nm).
// For detector point d and the reference point c of a source tile,  D = d - c,  S = |D|^2 (double-double, high
// word S, low word folded into the reference phase) and  R = RN(sqrt(S))  are evaluated once per
// (detector, tile); for a source s_j = c + e_j of the tile
//     ds  = r_ij^2 - |D|^2 = sum_c e_c (e_c - 2 D_c)            (small numbers, full relative precision)
//     rho = sqrt(S + ds)  by one Goldschmidt step from MUFU.RSQ64H:  g1 ~ rho (2^-41),  h1 ~ 1/(2 rho)
//     r_ij - R = (g1 - R) + (S - g1^2 + ds) h1                   (the Newton residual of rho, evaluated AROUND R:
//                                                                 g1 - R is exact, S - g1^2 + ds loses nothing)
// so the path difference carries 1 ulp of ITSELF (<= 2e-18 m for a 15 mm tile) without any division, at the cost
// of the plain square root (11 FP64 instructions for ds -> (r - R, 1/(2r)); the quotient form ds/(rho + R) of
// round 1 took 15).  The phase of a pair is  k R (mod 2 pi, per (detector, tile))  +  k (r_ij - R)  reduced
// exactly.  The first part is NOT added per pair: the accumulator of a detector point lives in the phase frame
// of the current tile and is rotated by exp(-i k (R_prev - R_new)) when the tile changes (a complex multiply per
// (detector, tile)), and by exp(-i k R_last) before it is stored.
struct RefCtx {
    double gx, gy, gz; // -2 D (high words)
    double s_ref;      // S: |D|^2 (high word)
    double r_ref;      // R = RN(sqrt(S))
    double phi;        // frac(k (R + low word correction) / u) in [-1/2, 1/2]
    int n_ref;         // rint(k R / u) mod 2^32
};

__device__ __forceinline__ void two_sum(double a, double b, double &s, double &e)
{
    s = add(a, b);
    const double bb = sub(s, a);
    e = add(sub(a, sub(s, bb)), sub(b, bb));
}

__device__ __forceinline__ RefCtx make_ref_ctx(double X, double Y, double Z, double cx, double cy, double cz,
                                               const PhaseConst &pc, double magic)
{
    double dh[3], dl[3];
    two_sum(X, -cx, dh[0], dl[0]); // D = d - c exactly as hi + lo
    two_sum(Y, -cy, dh[1], dl[1]);
    two_sum(Z, -cz, dh[2], dl[2]);
    double sh = 0.0, sl = 0.0; // |D|^2 in double-double
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double p = mul(dh[c], dh[c]);
        const double pe = fma_(dh[c], dh[c], -p) + 2.0 * dh[c] * dl[c];
        double s1, e1;
        two_sum(sh, p, s1, e1);
        sh = s1;
        sl = sl + e1 + pe;
    }
    two_sum(sh, sl, sh, sl);
    RefCtx r;
    r.gx = -2.0 * dh[0];
    r.gy = -2.0 * dh[1];
    r.gz = -2.0 * dh[2];
    r.s_ref = sh;
    const double r0 = __dsqrt_rn(sh);
    // the pairs measure their distance from R = r0 itself (not from sqrt(sh)); what the reference phase still
    // owes is the low word of |D|^2:  sqrt(sh + sl + ds) = sqrt(sh + ds) + sl / (2 r)  (r ~ r0 here)
    double rl = 0.0;
    if (r0 > 0.0) rl = __ddiv_rn(sl, 2.0 * r0);
    r.r_ref = r0;
    // k R / u = n_ref + phi with (r0 + rl) * (q_hi + q_lo)
    const double t0 = fma_(r0, pc.tq_hi, magic);
    const double n0 = add(t0, KC[KC_NEG_MAGIC]);
    double phi = fma_(r0, pc.tq_hi, -n0);
    phi = fma_(r0, pc.tq_lo, phi);
    phi = fma_(rl, pc.tq_hi, phi);
    r.phi = phi;
    r.n_ref = __double2loint(t0);
    return r;
}

// Table entry (q mod TBL): one LOP3 for the index, then LDS.128 [R.X16 + UR] -- the scaling and the
// (uniform) table base ride in the load's address mode, so the table needs no alignment.  (An XOR
// swizzle of the bank-group bits that spreads power-of-two index strides across the lanes was
// measured twice: 442 vs 449 Gterms/s on C3 in the round-1 kernel, 522.8 vs 521.7 in the present one --
// the shared-memory pipe is not what limits this kernel.)
template <int TBL, bool SWZ = false>
__device__ __forceinline__ double2 table_entry(const double2 *table, int q)
{
    if (SWZ) return table[(q ^ ((q >> 3) & 7)) & (TBL - 1)];
    return table[q & (TBL - 1)];
}

// acc <- acc * exp(-i (m u + dphi u)):  m whole table steps (any int: the table is periodic in 2^32) and a
// fraction |dphi| <= 1.  `sg` = +1 when the imaginary accumulator holds -Im (FORM_WFOLD), -1 when it holds +Im.
template <int TBL, bool SWZ>
__device__ __forceinline__ void rotate_acc(double &ar, double &ai, const double2 *table, int m, double dphi,
                                           const PhaseConst &pc, double sg)
{
    const double2 cs = table_entry<TBL, SWZ>(table, m);
    const double x = mul(dphi, pc.tu); // |x| <= 2 pi / TBL
    const double z = mul(x, x);
    const double c = fma_(z, fma_(z, 1.0 / 24.0, -0.5), 1.0);                    // next term x^6/720 <= 2e-20
    const double sn = mul(x, fma_(z, fma_(z, 1.0 / 120.0, -1.0 / 6.0), 1.0));    // next term x^7/5040
    const double cr = fma_(-cs.y, sn, mul(cs.x, c));                             // cos(m u + x)
    const double sr = mul(sg, fma_(cs.x, sn, mul(cs.y, c)));                     // sin(m u + x), signed for the storage form
    const double nr = fma_(-ai, sr, mul(ar, cr));
    ai = fma_(ar, sr, mul(ai, cr));
    ar = nr;
}

// phase A of a REFERENCED pair from ds = r^2 - |D|^2
__device__ __forceinline__ PairA pair_phase_a_ref_from_ds(const RefCtx &rc, double ds, const PhaseConst &pc, double magic)
{
    const double s = add(rc.s_ref, ds);
    const double y = rsqrt_approx(s);
    const double g = mul(s, y);
    const double h0 = __hiloint2double(__double2hiint(y) - 0x00100000, 0); // y/2, exact (akb_common.cuh)
    const double e = fma_(-g, h0, 0.5);
    const double g1 = fma_(g, e, g);
    PairA a;
    a.h = fma_(h0, e, h0);
    const double d = add(fma_(-g1, g1, rc.s_ref), ds);  // S + ds - g1^2: the first part is -ds to within 2^-41, the sum exact
    a.p = fma_(d, a.h, sub(g1, rc.r_ref));              // r - R
    a.t = fma_(a.p, pc.q_hi, magic);                    // MAGIC - m_j + rint(k (r - R)/u)
    return a;
}

// (ex, ey, ez) = source - tile reference; r^2 - |D|^2 = sum_c e_c (e_c - 2 D_c) = |e|^2 + e . g,  g = -2 D.
// E2: |e|^2 comes from the tile (FORM_E2).
template <bool E2>
__device__ __forceinline__ PairA pair_phase_a_ref(const RefCtx &rc, double ex, double ey, double ez, double e2,
                                                  const PhaseConst &pc, double magic)
{
    if (E2) return pair_phase_a_ref_from_ds(rc, fma_(ex, rc.gx, fma_(ey, rc.gy, fma_(ez, rc.gz, e2))), pc, magic);
    const double ax = add(ex, rc.gx), ay = add(ey, rc.gy), az = add(ez, rc.gz);
    return pair_phase_a_ref_from_ds(rc, fma_(ex, ax, fma_(ey, ay, mul(ez, az))), pc, magic);
}

// planar-row block: the x and z terms (E2: and |e|^2) are shared by the thread's points, xz holds them
template <bool E2>
__device__ __forceinline__ PairA pair_phase_a_ref_row(const RefCtx &rc, double xz, double ey, const PhaseConst &pc,
                                                      double magic)
{
    return pair_phase_a_ref_from_ds(rc, E2 ? fma_(ey, rc.gy, xz) : fma_(ey, add(ey, rc.gy), xz), pc, magic);
}

// r^2 -> (phase, 1/(2r), MAGIC + rint(phase/u))
template <int MODE>
__device__ __forceinline__ PairA pair_phase_a_from_s(double s, const PhaseConst &pc, double magic)
{
    PairA a;
    double root;
    if (MODE == AKB_PHASE_FAITHFUL)
        sqrt_and_half_rinv(s, root, a.h); // correctly rounded, like np.sqrt (CPU0402:80)
    else
        sqrt_and_half_rinv_g2(s, root, a.h);
    if (MODE == AKB_PHASE_FAITHFUL) {
        a.p = mul(pc.k, root); // CPU0402:82: |phase| = fl(k*dist)
        a.t = fma_(a.p, pc.inv_u, magic);
    } else {
        a.p = root;
        a.t = fma_(root, pc.q_hi, magic); // k*r is never rounded
    }
    return a;
}

template <int MODE>
__device__ __forceinline__ PairA pair_phase_a(double X, double Y, double Z, double sx, double sy, double sz,
                                              const PhaseConst &pc, double magic)
{
    const double ddx = sub(X, sx), ddy = sub(Y, sy), ddz = sub(Z, sz);
    // FAITHFUL: CPU0402:76-80, (dx*dx + dy*dy) + dz*dz with one rounding per operation
    const double s = MODE == AKB_PHASE_FAITHFUL ? add(add(mul(ddx, ddx), mul(ddy, ddy)), mul(ddz, ddz))
                                                : fma_(ddz, ddz, fma_(ddy, ddy, mul(ddx, ddx)));
    return pair_phase_a_from_s<MODE>(s, pc, magic);
}

// The same pair in a "planar row" block (see the kernel): dxx = fl((X - sx)^2) comes from the tile (one
// value per source for the whole block), ddz = Z - sz is shared by the thread's points.  Same operations
// in the same order as pair_phase_a, hence the same bits.
template <int MODE>
__device__ __forceinline__ PairA pair_phase_a_row(double dxx, double Y, double sy, double ddz, double dzz,
                                                  const PhaseConst &pc, double magic)
{
    const double ddy = sub(Y, sy);
    const double s = MODE == AKB_PHASE_FAITHFUL ? add(add(dxx, mul(ddy, ddy)), dzz)
                                                : fma_(ddz, ddz, fma_(ddy, ddy, dxx));
    return pair_phase_a_from_s<MODE>(s, pc, magic);
}

// exact Cody-Waite reduction + h*(cos f, sin f): n = rint(p/u), f = p - n*u.  fma(n, -u_hi, p) is
// exact (the difference fits in 53 bits); the u_lo term restores the bits of u beyond double.
// TAN = false: returns (cf, sf) = h (cos f, sin f).
// TAN = true:  returns cf = h cos f and sf = tan f, for the rotation  h e^{-i(theta+f)} =
//              cf (C - S tan f) - i cf (S + C tan f): one FP64 instruction less per pair.
//
// FORM_POLAR: `magic_j` = MAGIC - m_j is the rounding constant a.t was formed with and `g_j` the
// (negated) remainder of the source's weight phase; |f| then reaches pi/TBL * 2.
//
// FORM_WFOLD: `w_j` = |w_j| and `nhw_j` = -|w_j|/2 enter the cosine polynomial, so cf comes back as
// W = h |w_j| cos f, the factor of the whole pair.
template <int MODE, int TBL, int FORM>
__device__ __forceinline__ void pair_phase_b(const PairA &a, const PhaseConst &pc, double t2,
                                             double magic_j, double g_j, double w_j, double nhw_j, double &cf,
                                             double &sf)
{
    constexpr bool POLAR = (FORM & FORM_POLAR) != 0;
    static_assert(!(FORM & FORM_WFOLD) || (FORM & (FORM_TAN | FORM_POLAR | FORM_SHORTCOS)) ==
                                              (FORM_TAN | FORM_POLAR | FORM_SHORTCOS),
                  "FORM_WFOLD builds on the tan / polar / short-cosine formulation");
    constexpr int STEPS = POLAR ? TBL / 2 : TBL; // |f| <= pi / STEPS
    const double n = POLAR ? sub(a.t, magic_j) : add(a.t, KC[KC_NEG_MAGIC]);
    double f;
    if (MODE == AKB_PHASE_FAITHFUL) {
        f = fma_(n, pc.neg_u_hi, a.p);
        f = fma_(n, pc.neg_u_lo, f);
        if (POLAR) f = add(f, g_j);
    } else {
        f = fma_(a.p, pc.q_hi, -n); // exact: the difference fits in 53 bits
        f = fma_(a.p, pc.q_lo, f);
        f = POLAR ? fma_(f, pc.u, g_j) : mul(f, pc.u);
    }
    const double z = mul(f, f);
    static_assert(!(FORM & FORM_SHORTCOS) || STEPS >= 2048, "short cosine needs |f| <= pi/2048");
    const double c1 = (FORM & FORM_WFOLD)      ? fma_(z, nhw_j, w_j)                       // |w| (1 - f^2/2)
                      : (FORM & FORM_SHORTCOS) ? fma_(z, -0.5, 1.0)                        // f^4/24 <= 2.3e-13
                                               : fma_(z, fma_(KC[KC_TC1], z, -0.5), 1.0);  // cos f
    if (FORM & FORM_TAN) {
        double e1;
        if (STEPS >= 1024) {
            e1 = fma_(z, KC[KC_TT1], 1.0); // |f| <= pi/1024: the 2 f^5/15 term is <= 3.6e-14 (1.1e-15 at pi/2048)
        } else {
            e1 = fma_(z, fma_(t2, z, KC[KC_TT1]), 1.0); // tan f / f  (t2 = 2/15)
        }
        sf = mul(f, e1);
        cf = mul(a.h, c1);
        return;
    }
    double e1;
    if (STEPS >= 1024) {
        e1 = fma_(z, KC[KC_T1], 1.0); // |f| <= pi/1024: the z^2/120 term is 7e-13 of |f| <= 2e-15 absolute
    } else {
        e1 = fma_(z, fma_(t2, z, KC[KC_T1]), 1.0); // sin f / f
    }
    const double hf = mul(a.h, f);
    sf = mul(hf, e1);
    cf = mul(a.h, c1);
}

// Phase C -- rotate by the tabulated multiple (C, S) = (cos, sin)(2*pi*m/TBL) and accumulate
// (wr + i wi) * (c - i sn)   [exp(-i k r)/r, CPU0402:81-84] -- is written out in the kernel loop.

// ---------------------------------------------------------------- pair kernel
// grid.x: blocks of THREADS*DPT detector points; grid.y: splits of the source tiles.
// out: [gridDim.y][M] complex partial sums (gridDim.y == 1 -> the result itself).
template <int TILE, int STAGES, int TBL, int FORM>
struct PairCfg {
    static constexpr int kRows = rows_of(FORM);
    static constexpr int kTileBytes = (kRows * TILE + HEAD) * 8; // rows + the tile's reference point
    static constexpr int kTableBytes = TBL * 16;
    // tiles | mbarriers (padded to 16 B) | 2 loop constants | table
    static constexpr int kBarBytes = (STAGES * 8 + 15) & ~15;
    static constexpr int kTableOffset = STAGES * kTileBytes + kBarBytes + 16;
    static constexpr int kScratchOffset = kTableOffset + kTableBytes; // FORM_PLANES: (x_p - X_j)^2 of planes 1..3
    static constexpr int kSmemBytes = kScratchOffset + ((FORM & FORM_PLANES) ? 3 * TILE * 8 : 0);
};

template <int DPT, int MODE, int TILE, int STAGES, int TBL, int MINB, int FORM, int SPI, int THREADS>
__global__ void __launch_bounds__(THREADS, MINB) fresnel_pairs_kernel(
    const double *__restrict__ det_x, const double *__restrict__ det_y, const double *__restrict__ det_z,
    long long M, const double *__restrict__ packed, int tiles_total, int tiles_per_split, long long n_padded,
    const __grid_constant__ PhaseConst pc, double *__restrict__ out)
{
    using Cfg = PairCfg<TILE, STAGES, TBL, FORM>;
    constexpr int ROWS = Cfg::kRows;
    constexpr bool POLAR = (FORM & FORM_POLAR) != 0;
    constexpr bool TAN = (FORM & FORM_TAN) != 0;
    constexpr int TILE_DOUBLES = ROWS * TILE + HEAD;
    constexpr bool REF = MODE == AKB_PHASE_REFERENCED;
    constexpr bool E2 = (FORM & FORM_E2) != 0;
    constexpr bool ROWT = REF && (FORM & FORM_ROWT) != 0;
    constexpr bool PL = (FORM & FORM_PLANES) != 0;
    static_assert(!PL || (DPT == 4 && SPI == 2 && MODE != AKB_PHASE_REFERENCED), "FORM_PLANES: 4 planes per thread, FAITHFUL / EXACT");
    static_assert(!(FORM & FORM_ROWT) || (E2 && DPT >= 2), "FORM_ROWT builds on FORM_E2 and several points per thread");
    static_assert(SPI == 1 || SPI == 2, "1 or 2 sources per loop iteration");
    constexpr int NP = SPI * DPT; // pairs per loop iteration: DPT detector points x SPI sources
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *tiles = reinterpret_cast<double *>(smem_raw);
    const uint32_t tiles_s = smem_u32(tiles);
    const uint32_t bars_s = tiles_s + STAGES * Cfg::kTileBytes;
    const uint32_t consts_s = bars_s + Cfg::kBarBytes;
    double2 *table = reinterpret_cast<double2 *>(smem_raw + Cfg::kTableOffset);

    const int t0 = blockIdx.y * tiles_per_split;
    const int t1 = min(t0 + tiles_per_split, tiles_total);
    // a thread owns DPT CONSECUTIVE detector points: in a meshgrid-ordered focal grid they lie in one row
    // (FORM_PLANES: one pixel on DPT consecutive planes)
    const long long base = PL ? (long long)blockIdx.x * THREADS + threadIdx.x : ((long long)blockIdx.x * THREADS + threadIdx.x) * DPT;
    double *scratch = reinterpret_cast<double *>(smem_raw + Cfg::kScratchOffset);

    double X[DPT], Y[DPT], Z[DPT], ar[DPT], ai[DPT];
#pragma unroll
    for (int d = 0; d < DPT; ++d) {
        if (PL) {
            const int pl = (int)blockIdx.z * DPT + d;
            const long long ic = base < M ? base : M - 1;
            X[d] = det_x[pl < pc.planes ? pl : pc.planes - 1];
            Y[d] = det_y[ic];
            Z[d] = det_z[ic];
        } else {
            long long i = base + d;
            long long ic = i < M ? i : M - 1;
            X[d] = det_x[ic];
            Y[d] = det_y[ic];
            Z[d] = det_z[ic];
        }
        ar[d] = 0.0;
        ai[d] = 0.0;
    }
    // "Planar row" block: every detector point of the block has the same x (a detector plane x = const) and
    // the points of each thread share z (a row of the grid).  Then (X - sx)^2 is one value per SOURCE -- the
    // block computes it once per tile, in place of the sx row -- and (Z - sz)^2 one value per (thread,
    // source): 4.5 instead of 8 FP64 instructions for r^2 per pair, bit-identical results.  Detected here
    // from the coordinates themselves (block-uniform vote), so irregular detector sets (mirror surfaces)
    // simply take the general loop.
    bool row = PL; // FORM_PLANES: by construction
    if (!PL) {
        const long long i0 = (long long)blockIdx.x * THREADS * DPT;
        const double x0 = det_x[i0 < M ? i0 : M - 1];
        bool mine = true;
#pragma unroll
        for (int d = 0; d < DPT; ++d) mine = mine && X[d] == x0 && Z[d] == Z[0];
        row = __syncthreads_and(mine);
        if (ROWT && row) {
            // FORM_ROWT expands r along the thread's row about its first point: allowed when the third-order term
            // |a| D^3 / (2 r^2) stays below 1e-11 rad for every source, r bounded below by the distance from the point
            // to the bounding box of the source set (left behind the packed tiles by the pack kernel)
            const unsigned long long *bb = reinterpret_cast<const unsigned long long *>(packed + (long long)tiles_total * TILE_DOUBLES);
            double r2 = 0.0, dmax = 0.0;
            const double P0[3] = {X[0], Y[0], Z[0]};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double lo = dunkey(bb[c]), hi = dunkey(bb[3 + c]);
                const double gap = fmax(fmax(lo - P0[c], P0[c] - hi), 0.0);
                r2 = fma_(gap, gap, r2);
            }
#pragma unroll
            for (int d = 1; d < DPT; ++d) dmax = fmax(dmax, fabs(sub(Y[d], Y[0])));
            const bool ok = r2 > 0.0 && pc.k * dmax * dmax * dmax <= 2.0e-11 * r2 && dmax * dmax <= 1.0e-8 * r2;
            row = __syncthreads_and(ok); // else: the general loop (every point takes its own square root)
        }
        if (row && threadIdx.x == 0) atomicAdd(&g_row_blocks, 1ULL);
    }

    for (int m = threadIdx.x; m < TBL; m += THREADS) {
        double sv, cv;
        sincospi((double)m * (2.0 / TBL), &sv, &cv); // exact argument: accurate to < 1 ulp
        table[(FORM & FORM_SWZ) ? (m ^ ((m >> 3) & 7)) : m] = make_double2(cv, sv);
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(bars_s + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // Two constants share an instruction with another constant operand (fma(p, 1/u, MAGIC),
        // fma(T2, z, T1)); a DFMA takes only one constant/uniform operand, and ptxas would re-create
        // the second one with two IMAD.U32 per use.  Bouncing them through shared memory yields
        // loop-invariant VECTOR registers that cannot be re-materialised.
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(consts_s), "d"(KC[TAN ? KC_TT2 : KC_T2]) : "memory");
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(consts_s + 8), "d"(KC[KC_MAGIC]) : "memory");
    }
    __syncthreads();
    double t2, magic;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t2) : "r"(consts_s) : "memory");
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(magic) : "r"(consts_s + 8) : "memory");
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES && t0 + s < t1; ++s) {
            mbar_expect_tx(bars_s + 8 * s, Cfg::kTileBytes);
            tma_bulk_load(tiles_s + s * Cfg::kTileBytes, packed + (long long)(t0 + s) * TILE_DOUBLES, Cfg::kTileBytes,
                          bars_s + 8 * s);
        }
    }

    // one pass over this block's source tiles; ROW selects the planar-row form of r^2
    auto run_tiles = [&](auto row_tag) {
        constexpr bool ROW = decltype(row_tag)::value;
        int stage = 0;
        uint32_t parity = 0;
        RefCtx rc[REF ? DPT : 1]; // REFERENCED: the phase frame (tile reference) the accumulators live in
        constexpr double SG = (FORM & FORM_WFOLD) ? 1.0 : -1.0;
        constexpr bool RT = ROWT && ROW; // only the thread's first point carries a reference context / takes the root
        double dlt[DPT];                 // RT: offsets of the thread's points from its first one along the row
#pragma unroll
        for (int d = 0; d < DPT; ++d) dlt[d] = RT ? sub(Y[d], Y[0]) : 0.0;
        for (int t = t0; t < t1; ++t) {
            mbar_wait(bars_s + 8 * stage, parity);
            double *T = tiles + stage * TILE_DOUBLES;
            const long long left = n_padded - (long long)t * TILE;
            const int cnt = left < TILE ? (int)left : TILE; // multiple of 2
            if (RT) { // one context per thread: all its accumulators live in the phase frame of its first point
                const RefCtx nc = make_ref_ctx(X[0], Y[0], Z[0], T[ROWS * TILE + 0], T[ROWS * TILE + 1], T[ROWS * TILE + 2], pc, magic);
                if (t > t0) {
#pragma unroll
                    for (int d = 0; d < DPT; ++d)
                        rotate_acc<TBL, (FORM & FORM_SWZ) != 0>(ar[d], ai[d], table, rc[0].n_ref - nc.n_ref, sub(rc[0].phi, nc.phi), pc, SG);
                }
                rc[0] = nc;
            } else if (REF) { // once per (detector point, tile): double-double distance to the tile's reference point,
                       // and the accumulated field moves from the previous tile's phase frame into this one
#pragma unroll
                for (int d = 0; d < DPT; ++d) {
                    const RefCtx nc = make_ref_ctx(X[d], Y[d], Z[d], T[ROWS * TILE + 0], T[ROWS * TILE + 1],
                                                   T[ROWS * TILE + 2], pc, magic);
                    if (t > t0) rotate_acc<TBL, (FORM & FORM_SWZ) != 0>(ar[d], ai[d], table, rc[d].n_ref - nc.n_ref, sub(rc[d].phi, nc.phi), pc, SG);
                    rc[d] = nc;
                }
            }
            if (ROW) { // the sx row becomes fl((x0 - sx)^2), CPU0402:76-77 (REFERENCED: e_x (e_x - 2 D_x);
                       // FORM_E2: the |e|^2 row becomes |e|^2 + e_x (-2 D_x))
                for (int q = threadIdx.x; q < TILE; q += THREADS) {
                    if (PL) { // (x_p - X_j)^2 for the block's four planes: CPU0402:76-77 per plane
                        const double sxq = T[q];
#pragma unroll
                        for (int d = 0; d < DPT; ++d) {
                            const double ddx = sub(X[d], sxq);
                            (d == 0 ? T[q] : scratch[(d - 1) * TILE + q]) = mul(ddx, ddx);
                        }
                    } else if (REF && E2) {
                        T[7 * TILE + q] = fma_(T[q], rc[0].gx, T[7 * TILE + q]);
                    } else {
                        const double ddx = REF ? add(T[q], rc[0].gx) : sub(X[0], T[q]);
                        T[q] = mul(REF ? T[q] : ddx, ddx);
                    }
                }
                // generic-proxy writes to a stage that a later bulk copy (async proxy) overwrites
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
            }
#pragma unroll 1
            for (int j = 0; j < cnt; j += SPI) {
                // rows of the SPI sources of this iteration: x (ROW: dx^2), y, z, then (w_re, w_im) -- or,
                // FORM_POLAR: (|w|, -frac(arg w), MAGIC - m)
                double S[MAX_ROWS][SPI];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    if (SPI == 2) {
                        const double2 v = *reinterpret_cast<const double2 *>(T + r * TILE + j);
                        S[r][0] = v.x;
                        S[r][SPI - 1] = v.y;
                    } else {
                        S[r][0] = T[r * TILE + j];
                    }
                }
                if (!POLAR) {
#pragma unroll
                    for (int q = 0; q < SPI; ++q) S[5][q] = magic;
                }
                double ddz[SPI], dzz[SPI], ddy[SPI], dyy[SPI], dxp[DPT][SPI];
                if (PL) { // one pixel per thread: the y term is shared by its four planes like the z term
#pragma unroll
                    for (int q = 0; q < SPI; ++q) {
                        ddy[q] = sub(Y[0], S[1][q]);
                        dyy[q] = mul(ddy[q], ddy[q]);
                        dxp[0][q] = S[0][q];
                    }
#pragma unroll
                    for (int d = 1; d < DPT; ++d) {
                        const double2 v = *reinterpret_cast<const double2 *>(scratch + (d - 1) * TILE + j);
                        dxp[d][0] = v.x;
                        dxp[d][SPI - 1] = v.y;
                    }
                }
                if (ROW) {
#pragma unroll
                    for (int q = 0; q < SPI; ++q) {
                        if (REF && E2) { // |e|^2 + e_x g_x (from the tile) + e_z g_z
                            ddz[q] = 0.0;
                            dzz[q] = fma_(S[2][q], rc[0].gz, S[7][q]);
                            continue;
                        }
                        ddz[q] = REF ? add(S[2][q], rc[0].gz) : sub(Z[0], S[2][q]);
                        dzz[q] = mul(REF ? S[2][q] : ddz[q], ddz[q]);
                        if (REF) dzz[q] = add(S[0][q], dzz[q]); // the x and z terms of r^2 - r_ref^2
                    }
                }
                PairA a[NP];
                double2 cs[NP];
                double cf[NP], sf[NP];
                double rta[SPI], rtb[SPI], rtm[SPI]; // RT: dr/dy, (d2r/dy2)/2 and d(1/2r)/dy at the thread's first point
#pragma unroll
                for (int d = 0; d < DPT; ++d) {
#pragma unroll
                    for (int q = 0; q < SPI; ++q) {
                        const int i = SPI * d + q;
                        if (RT && d > 0) {
                            if (d == 1) { // from the first point's pair: w = y0 - Y_j = -(g_y/2 + e_y), h = 1/(2r)
                                const double h = a[q].h;
                                rta[q] = mul(-fma_(2.0, S[1][q], rc[0].gy), h);          // a = 2 w h = w / r
                                rtb[q] = mul(fma_(-rta[q], rta[q], 1.0), h);            // b = (1 - a^2) / (2r)
                                rtm[q] = mul(mul(rta[q], h), mul(-2.0, h));             // d(1/2r)/dy = -a / (2 r^2)
                            }
                            a[i].p = fma_(dlt[d], fma_(rtb[q], dlt[d], rta[q]), a[q].p);  // r - R
                            a[i].h = fma_(rtm[q], dlt[d], a[q].h);
                            a[i].t = fma_(a[i].p, pc.q_hi, S[5][q]);
                        } else if (REF && ROW) {
                            a[i] = pair_phase_a_ref_row<E2>(rc[d], dzz[q], S[1][q], pc, S[5][q]);
                        } else if (REF) {
                            a[i] = pair_phase_a_ref<E2>(rc[d], S[0][q], S[1][q], S[2][q], S[7][q], pc, S[5][q]);
                        } else if (PL) {
                            // (dx*dx + dy*dy) + dz*dz, CPU0402:76-80 (EXACT: the fused form of pair_phase_a)
                            const double s2 = MODE == AKB_PHASE_FAITHFUL ? add(add(dxp[d][q], dyy[q]), dzz[q])
                                                                          : fma_(ddz[q], ddz[q], fma_(ddy[q], ddy[q], dxp[d][q]));
                            a[i] = pair_phase_a_from_s<MODE>(s2, pc, S[5][q]);
                        } else if (ROW) {
                            a[i] = pair_phase_a_row<MODE>(S[0][q], Y[d], S[1][q], ddz[q], dzz[q], pc, S[5][q]);
                        } else {
                            a[i] = pair_phase_a<MODE>(X[d], Y[d], Z[d], S[0][q], S[1][q], S[2][q], pc, S[5][q]);
                        }
                        cs[i] = table_entry<TBL, (FORM & FORM_SWZ) != 0>(table, __double2loint(a[i].t));
                    }
                }
#pragma unroll
                for (int i = 0; i < NP; ++i)
                    pair_phase_b<MODE, TBL, FORM>(a[i], pc, t2, S[5][i % SPI],
                                                  S[4][i % SPI], S[3][i % SPI], S[6][i % SPI], cf[i], sf[i]);
                // phase C: rotate by the table entry, then accumulate.  The accumulation is ordered by
                // source and by operation so that consecutive DFMAs share their first operand (the weight
                // of one source): served by the operand-reuse cache, they read 2 registers, not 3.
                if constexpr ((FORM & FORM_WFOLD) != 0) {
                    // cf = W = h |w_j| cos f, sf = tan f:  acc += W ((C - S tan f) - i (S + C tan f))
                    // Two passes, so that the two DFMAs that share an operand (tan f, then W) are neighbours in
                    // program order with their inputs long since ready: ptxas then keeps them adjacent and the
                    // operand-reuse cache serves the shared operand (a DFMA reading three fresh registers
                    // costs about two extra FP64-pipe cycles, tools/ubench/fp64_banks3.cu).
                    double ca[NP], sb[NP];
#pragma unroll
                    for (int i = 0; i < NP; ++i) {
                        ca[i] = fma_(-cs[i].y, sf[i], cs[i].x);
                        sb[i] = fma_(cs[i].x, sf[i], cs[i].y);
                    }
#pragma unroll
                    for (int i = 0; i < NP; ++i) {
                        ar[i / SPI] = fma_(cf[i], ca[i], ar[i / SPI]);
                        ai[i / SPI] = fma_(cf[i], sb[i], ai[i / SPI]); // -Im: the sign is applied at the store
                    }
                } else {
                double c[NP], sn[NP];
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    const double m1 = mul(cs[i].x, cf[i]);
                    const double m2 = mul(cs[i].y, cf[i]);
                    if (TAN) { // sf = tan f
                        c[i] = fma_(-m2, sf[i], m1);
                        sn[i] = fma_(m1, sf[i], m2);
                    } else {
                        c[i] = fma_(-cs[i].y, sf[i], m1);
                        sn[i] = fma_(cs[i].x, sf[i], m2);
                    }
                }
#pragma unroll
                for (int q = 0; q < SPI; ++q) {
                    const double wr = S[3][q], wi = S[4][q]; // POLAR: wr = |w_j| (real weight), wi unused
#pragma unroll
                    for (int d = 0; d < DPT; ++d) ar[d] = fma_(wr, c[SPI * d + q], ar[d]);
#pragma unroll
                    for (int d = 0; d < DPT; ++d) ai[d] = fma_(-wr, sn[SPI * d + q], ai[d]);
                    if (!POLAR) {
#pragma unroll
                        for (int d = 0; d < DPT; ++d) ai[d] = fma_(wi, c[SPI * d + q], ai[d]);
#pragma unroll
                        for (int d = 0; d < DPT; ++d) ar[d] = fma_(wi, sn[SPI * d + q], ar[d]);
                    }
                }
                }
            }
            __syncthreads(); // every thread is done with this stage
            if (threadIdx.x == 0 && t + STAGES < t1) {
                mbar_expect_tx(bars_s + 8 * stage, Cfg::kTileBytes);
                tma_bulk_load(tiles_s + stage * Cfg::kTileBytes, packed + (long long)(t + STAGES) * TILE_DOUBLES,
                              Cfg::kTileBytes, bars_s + 8 * stage);
            }
            if (++stage == STAGES) {
                stage = 0;
                parity ^= 1;
            }
        }
        if (REF && t1 > t0) { // out of the last tile's phase frame
#pragma unroll
            for (int d = 0; d < DPT; ++d)
                rotate_acc<TBL, (FORM & FORM_SWZ) != 0>(ar[d], ai[d], table, rc[RT ? 0 : d].n_ref, rc[RT ? 0 : d].phi, pc, SG);
        }
    };
    if (row)
        run_tiles(std::true_type{});
    else
        run_tiles(std::false_type{});

    double2 *o = reinterpret_cast<double2 *>(out) + (long long)blockIdx.y * (PL ? M * pc.planes : M);
#pragma unroll
    for (int d = 0; d < DPT; ++d) {
        const double2 v = make_double2(ar[d], mul((FORM & FORM_WFOLD) ? -pc.im_sign : pc.im_sign, ai[d]));
        if (PL) {
            const int pl = (int)blockIdx.z * DPT + d;
            if (pl < pc.planes && base < M) o[(long long)pl * M + base] = v;
        } else {
            long long i = base + d;
            if (i < M) o[i] = v;
        }
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

// ---------------------------------------------------------------- host side
// a * (b_hi + b_lo) as an unevaluated sum hi + lo (~2^-100 relative)
void dd_scale(double a, double b_hi, double b_lo, double &hi, double &lo)
{
    double ph = a * b_hi;
    double pl = __builtin_fma(a, b_hi, -ph) + a * b_lo;
    hi = ph + pl;
    lo = pl - (hi - ph);
}

PhaseConst make_phase_const(double k, int table)
{
    const double pi_hi = 3.141592653589793116e+00, pi_lo = 1.2246467991473532e-16;
    const double ipi_hi = 3.183098861837906912e-01, ipi_lo = -1.9678676675182486e-17; // 1/pi
    // u = 2*pi/table: a power-of-two multiple of pi, exact in both words
    const double su = 2.0 / table, si = table / 2.0;
    PhaseConst pc;
    pc.im_sign = k < 0.0 ? -1.0 : 1.0;
    k = k < 0.0 ? -k : k;
    pc.k = k;
    pc.u = pi_hi * su;
    pc.inv_u = ipi_hi * si;
    pc.neg_u_hi = -pi_hi * su;
    pc.neg_u_lo = -pi_lo * su;
    dd_scale(k, ipi_hi * si, ipi_lo * si, pc.q_hi, pc.q_lo);
    pc.tq_hi = pc.q_hi;
    pc.tq_lo = pc.q_lo;
    pc.tu = pc.u;
    pc.planes = 0;
    return pc;
}

struct KernelEntry {
    const char *name;
    int dpt, tile, stages, table, form, threads;
    const void *fn[3]; // per mode
    int smem;
};

// MODES: bit m set = the entry is built for phase mode m (an entry the default choice never takes in some mode is not
// compiled for it: every instantiation is ~100 KB of SASS and seconds of ptxas)
constexpr int BUILD_F = 1 << AKB_PHASE_FAITHFUL, BUILD_E = 1 << AKB_PHASE_EXACT, BUILD_R = 1 << AKB_PHASE_REFERENCED, BUILD_ALL = BUILD_F | BUILD_E | BUILD_R;

template <int DPT, int MODE, int TILE, int STAGES, int TBL, int MINB, int FORM, int SPI, int THREADS, bool BUILD>
const void *pair_kernel_ptr()
{
    if constexpr (BUILD)
        return reinterpret_cast<const void *>(&fresnel_pairs_kernel<DPT, MODE, TILE, STAGES, TBL, MINB, FORM, SPI, THREADS>);
    else
        return nullptr;
}

template <int DPT, int TILE, int STAGES, int TBL, int MINB, int FORM, int SPI = 2, int THREADS = 256, int MODES = BUILD_ALL>
KernelEntry make_entry(const char *name)
{
    KernelEntry e;
    e.name = name;
    e.dpt = DPT;
    e.tile = TILE;
    e.stages = STAGES;
    e.table = TBL;
    e.fn[0] = pair_kernel_ptr<DPT, AKB_PHASE_FAITHFUL, TILE, STAGES, TBL, MINB, FORM, SPI, THREADS, (MODES & BUILD_F) != 0>();
    e.fn[1] = pair_kernel_ptr<DPT, AKB_PHASE_EXACT, TILE, STAGES, TBL, MINB, FORM, SPI, THREADS, (MODES & BUILD_E) != 0>();
    e.fn[2] = pair_kernel_ptr<DPT, AKB_PHASE_REFERENCED, TILE, STAGES, TBL, MINB, FORM, SPI, THREADS, (MODES & BUILD_R) != 0>();
    e.smem = PairCfg<TILE, STAGES, TBL, FORM>::kSmemBytes;
    e.form = FORM;
    e.threads = THREADS;
    return e;
}

// Kernel variants: <points per thread, tile, stages, table entries, min resident blocks/SM, formulation
// [, sources per iteration, threads]>.  Entries 0..3 are the product: the size-aware default choice picks among
// them.  Further entries exist only in A/B builds (AKB_AB_VARIANTS=1 python -m akbraytracing_b200.build);
// AKB_FRESNEL_VARIANT=<n> forces one (tools/variant_bench.py) and an index that does not exist is an error.
constexpr int FORM_DEFAULT = FORM_TAN | FORM_POLAR | FORM_SHORTCOS | FORM_WFOLD;
enum { V_DEFAULT = 0, V_DPT2 = 1, V_DPT1 = 2, V_REFERENCED = 3 };
const KernelEntry *kernel_table(int *count_out)
{
    int dummy = 0;
    int *count = count_out ? count_out : &dummy;
    static const KernelEntry entries[] = {
        // default: 25.5 FP64 instructions per pair on planar-row blocks (29 in the general loop), 2 x 256
        // threads per SM (measured in round 1: 3 blocks/SM at 80 registers spill, 1 block/SM starves the FP64
        // pipe, block sizes that are not a multiple of 4 warps lose 10-20 %, one 640/768-thread block per SM
        // sharing one table loses 2-4 %)
        make_entry<4, 256, 3, 4096, 2, FORM_DEFAULT, 2, 256, BUILD_F | BUILD_E>("dpt4 tile256x3 table4096 tan polar shortcos wfold 2 blocks/SM"),
        // small problems: fewer points per thread, and a 2048-entry table (8 sincospi per thread and block)
        make_entry<2, 256, 3, 2048, 2, FORM_TAN | FORM_POLAR>("dpt2 tile256x3 table2048 tan polar"),
        make_entry<1, 128, 4, 2048, 2, FORM_TAN | FORM_POLAR>("dpt1 tile128x4 table2048 tan polar"),
        // REFERENCED: |e|^2 row + row expansion on planar-row blocks (20.75 FP64 instructions per pair there, 26 in the
        // general loop, which spills a little at 4 points per thread and is still faster than the 2-point form)
        make_entry<4, 256, 3, 4096, 2, FORM_DEFAULT | FORM_E2 | FORM_ROWT, 2, 256, BUILD_R>(
            "dpt4 tile256x3 table4096 wfold e2 row expansion 2 blocks/SM"),
#ifdef AKB_AB_VARIANTS
        // 4: the round-1 default (scaled table entry, 27.5 / 31 instructions per pair)
        make_entry<4, 256, 3, 4096, 2, FORM_TAN | FORM_POLAR | FORM_SHORTCOS>(
            "dpt4 tile256x3 table4096 tan polar shortcos 2 blocks/SM"),
        // 5: the first formulation (sin/cos polynomials, complex weights)
        make_entry<4, 512, 2, 1024, 3, 0>("dpt4 tile512x2 table1024 sincos 3 blocks/SM"),
        // 6, 7: one source per iteration; 2 points per thread with the default formulation
        make_entry<4, 256, 3, 4096, 2, FORM_DEFAULT, 1>("dpt4 spi1 tile256x3 table4096 wfold"),
        make_entry<8, 256, 3, 4096, 1, FORM_DEFAULT, 1>("dpt8 spi1 tile256x3 table4096 wfold 1 block/SM"),
        // 8: REFERENCED candidates: 4 points per thread with the |e|^2 row; 9: 2 points without it (round-2 first form)
        make_entry<4, 256, 3, 4096, 2, FORM_DEFAULT | FORM_E2>("dpt4 tile256x3 table4096 wfold e2 2 blocks/SM"),
        make_entry<2, 256, 3, 4096, 2, FORM_DEFAULT>("dpt2 tile256x3 table4096 wfold 2 blocks/SM"),
        // 10: the default with an XOR-swizzled table; 11: REFERENCED, 2 points per thread with the |e|^2 row, no row expansion
        make_entry<4, 256, 3, 4096, 2, FORM_DEFAULT | FORM_SWZ>("dpt4 tile256x3 table4096 wfold swizzled table"),
        make_entry<2, 256, 3, 4096, 2, FORM_DEFAULT | FORM_E2>("dpt2 tile256x3 table4096 wfold e2 2 blocks/SM"),
#endif
    };
    *count = (int)(sizeof(entries) / sizeof(entries[0]));
    return entries;
}

// Split the source tiles of one launch so that the grid (detector blocks x splits) fills whole waves of
// `slots` resident blocks.  Returns the split count; *eff_out = fraction of the slots x time rectangle
// that does work (wave quantisation x imbalance of the last split).
int plan_splits(long long blocks_x, int tiles_total, long long slots, long long M, double *eff_out)
{
    int splits = 1;
    double best = -1.0, best_raw = 0.0;
    int max_splits = tiles_total < 64 ? tiles_total : 64;
    const long long by_memory = (2LL << 30) / (16 * (M > 0 ? M : 1)); // partial sums [split][M] stay below 2 GiB
    if (max_splits > by_memory) max_splits = by_memory < 1 ? 1 : (int)by_memory;
    for (int s = 1; s <= max_splits; ++s) {
        const int tps = (tiles_total + s - 1) / s;
        const int s_eff = (tiles_total + tps - 1) / tps;
        if (s_eff != s) continue;
        const double waves = (double)(blocks_x * s) / (double)slots;
        const double full = waves <= 1.0 ? 1.0 : (double)(long long)(waves + 0.999999);
        double eff = waves / full;
        eff *= (double)tiles_total / ((double)tps * s); // the last split may hold fewer tiles
        // every split costs a table build per block and one more partial sum per detector point:
        // prefer fewer splits unless more of them fill the last wave measurably better
        const double score = eff - 2.0e-4 * s;
        if (score > best) {
            best = score;
            best_raw = eff;
            splits = s;
        }
    }
    if (eff_out) *eff_out = best_raw;
    return splits;
}

// resident blocks per SM of a variant (occupancy query, cached: every device of a box is the same part)
int resident_blocks(const KernelEntry &ke, int mode)
{
    static int cache[16][3]; // 0 = not asked yet
    int n = 0;
    const KernelEntry *e = kernel_table(&n);
    const int v = (int)(&ke - e);
    if (v >= 0 && v < 16 && cache[v][mode] > 0) return cache[v][mode];
    int per_sm = 1;
    if (!ke.fn[mode]) return 1; // not built for this mode (only reachable through AKB_FRESNEL_VARIANT)
    if (cudaFuncSetAttribute(ke.fn[mode], cudaFuncAttributeMaxDynamicSharedMemorySize, ke.smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ke.fn[mode], ke.threads, ke.smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    if (v >= 0 && v < 16) cache[v][mode] = per_sm;
    return per_sm;
}

// AKB_FRESNEL_VARIANT: -2 = not set, >= 0 = forced entry, -3 = set to something that does not exist
int forced_variant()
{
    static const int idx = [] {
        const char *v = getenv("AKB_FRESNEL_VARIANT");
        if (!v || !*v) return -2;
        int n = 0;
        kernel_table(&n);
        char *end = nullptr;
        const long want = strtol(v, &end, 10);
        return (end && *end == 0 && want >= 0 && want < n) ? (int)want : -3;
    }();
    return idx;
}

const KernelEntry &selected_kernel(int mode, long long M = -1, long long N = -1, int sms = 148)
{
    int n = 0;
    const KernelEntry *e = kernel_table(&n);
    const int idx = forced_variant();
    if (idx >= 0) return e[idx];
    int pick = mode == AKB_PHASE_REFERENCED ? V_REFERENCED : V_DEFAULT;
    // Small problems (C1: 64x64 detector points x 1e4 sources) cannot fill the SMs with 4 points per thread:
    // among the 4 / 2 / 1 points-per-thread variants take the one with the lowest estimated time,
    // FP64 instructions per pair / wave-fill efficiency of its best split plan.
    if (M > 0 && N > 0) {
        const int order[3] = {pick, V_DPT2, V_DPT1};
        const double instr[3] = {29.0, 32.0, 32.0};
        double best = 1e300;
        for (int o = 0; o < 3; ++o) {
            const KernelEntry &c = e[order[o]];
            const long long blocks = (M + c.threads * c.dpt - 1) / (c.threads * c.dpt);
            const long long tiles = (N + c.tile - 1) / c.tile;
            double eff = 0.0;
            plan_splits(blocks, (int)(tiles < (1LL << 30) ? tiles : (1LL << 30)), (long long)sms * resident_blocks(c, mode), M, &eff);
            const double cost = instr[o] / (eff > 1e-6 ? eff : 1e-6);
            if (cost < best * 0.97) { // a smaller variant must win by 3 %
                best = cost;
                pick = order[o];
            }
        }
    }
    return e[pick];
}

// optional in-library timing of the last akb_fresnel_sum call of this thread (bench.py roofline)
struct Timing {
    bool enabled = false;
    bool valid = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr}; // call start, pairs start, pairs end, call end
    int device = -1;                                          // events belong to one device
    int splits = 0, per_sm = 0;
    long long blocks_x = 0;
};
thread_local Timing g_timing;

int timing_mark(int idx, cudaStream_t st)
{
    if (!g_timing.enabled) return AKB_OK;
    int device = 0;
    AKB_CUDA(cudaGetDevice(&device));
    if (device != g_timing.device) {
        for (auto &e : g_timing.ev) {
            if (e) cudaEventDestroy(e);
            e = nullptr;
        }
        g_timing.device = device;
    }
    if (!g_timing.ev[idx]) AKB_CUDA(cudaEventCreate(&g_timing.ev[idx]));
    AKB_CUDA(cudaEventRecord(g_timing.ev[idx], st));
    return AKB_OK;
}

constexpr size_t kStageBytes = 4u << 20;

// pinned staging buffer of the calling host thread (allocated on first use, kept for the life of the thread)
char *host_stage()
{
    static thread_local char *buf = nullptr;
    if (!buf && cudaHostAlloc(reinterpret_cast<void **>(&buf), kStageBytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        buf = nullptr;
    }
    return buf;
}

} // namespace

namespace {

#ifdef AKB_AB_VARIANTS
// the through-focus kernel (FORM_PLANES): one pixel on four planes per thread; FAITHFUL and EXACT only.  24 instead of
// 25.5 FP64 instructions per pair, and still 8 % SLOWER than the plane-major flat set through the planar-row loop
// (568 vs 619 Gterms/s, profiles/r02_variants_ab.md section 7): ten row loads gate every loop iteration instead of seven
// and queue behind the other warps' table look-ups (ncu: short_scoreboard 1.69 instead of 0.42 warps per issue cycle, FP64
// pipe 74 % instead of 85 % active).  Kept as an A/B variant (AKB_PLANES_KERNEL=1 in an AKB_AB_VARIANTS build).
const KernelEntry &planes_kernel()
{
    static const KernelEntry e = make_entry<4, 256, 3, 4096, 2, FORM_DEFAULT | FORM_PLANES, 2, 256, BUILD_F | BUILD_E>(
        "dpt4 tile256x3 table4096 wfold, one pixel x four planes per thread");
    return e;
}
#else
const KernelEntry &planes_kernel() { return *kernel_table(nullptr); } // never launched in product builds (planes == 0 always)
#endif

// planes == 0: det_x/y/z are M detector points.  planes > 0 (akb_fresnel_sum_planes): det_x holds the `planes` plane
// positions, det_y/z the M pixels of one plane, out is [planes][M].
int fresnel_sum_impl(const double *det_x, const double *det_y, const double *det_z, int64_t M, int planes,
                     const double *src_x, const double *src_y, const double *src_z, const double *src_u,
                     const double *src_ds, int64_t N, double k, double *out, int mode, void *stream)
{
    AKB_REQUIRE(M >= 0 && N >= 0, "M and N must be non-negative");
    AKB_REQUIRE(mode == AKB_PHASE_FAITHFUL || mode == AKB_PHASE_EXACT || mode == AKB_PHASE_REFERENCED,
                "mode must be AKB_PHASE_FAITHFUL, AKB_PHASE_EXACT or AKB_PHASE_REFERENCED");
    if (M == 0) return AKB_OK;
    AKB_REQUIRE(det_x && det_y && det_z && out, "detector/out pointers must not be NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t M_out = planes > 0 ? M * planes : M; // field values written
    if (N == 0) { // empty sum (np.sum of an empty array) = 0
        fill_zero_kernel<<<(unsigned)((2 * M_out + 255) / 256), 256, 0, st>>>(out, 2 * M_out);
        AKB_LAUNCH_CHECK();
        return AKB_OK;
    }
    AKB_REQUIRE(src_x && src_y && src_z && src_u, "source pointers must not be NULL");
    AKB_REQUIRE(k > -1.0e12 && k < 1.0e12, "|k| must be below 1e12 (|k|*r must stay below 3.4e12 rad: 2^51 table steps)");
    AKB_REQUIRE(forced_variant() != -3, "AKB_FRESNEL_VARIANT names a kernel variant this build does not have "
                                        "(A/B variants need AKB_AB_VARIANTS=1 at build time)");

    int device = 0;
    AKB_CUDA(cudaGetDevice(&device));
    tune_pool(device);
    const int sms = sm_count(device);
    const KernelEntry &ke = planes > 0 ? planes_kernel() : selected_kernel(mode, M, N, sms);
    const void *kern = ke.fn[mode];
    AKB_REQUIRE(kern != nullptr, "the kernel variant forced with AKB_FRESNEL_VARIANT is not built for this phase mode");
    const int TILE = ke.tile;

    const int tiles_total = (int)((N + TILE - 1) / TILE);
    const long long padded = (long long)tiles_total * TILE;
    const long long n_padded = (N + 1) & ~1LL;

    // ---- plan: split the source tiles so the grid fills whole waves
    // (the attribute is per device: set it on every call, the occupancy figure itself is cached)
    AKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ke.smem));
    const int per_sm = resident_blocks(ke, mode);
    const long long slots = (long long)sms * per_sm;
    const long long blocks_x = planes > 0 ? (M + ke.threads - 1) / ke.threads : (M + ke.threads * ke.dpt - 1) / (ke.threads * ke.dpt);
    const int groups = planes > 0 ? (planes + ke.dpt - 1) / ke.dpt : 1; // grid.z: groups of four planes
    AKB_REQUIRE(groups <= 65535, "at most 262140 planes per call");
    const int splits = plan_splits(blocks_x * groups, tiles_total, slots, M_out, nullptr);
    int tiles_per_split = (tiles_total + splits - 1) / splits;

    int rc;
    g_timing.valid = false;
    g_timing.splits = splits;
    g_timing.per_sm = per_sm;
    g_timing.blocks_x = blocks_x;
    if ((rc = timing_mark(0, st))) return rc;

    // stream-ordered scratch, handed back to the pool on every exit path (also the error returns)
    struct Scratch {
        cudaStream_t st;
        double *p = nullptr;
        ~Scratch()
        {
            if (p) cudaFreeAsync(p, st);
        }
    } packed_s{st}, partial_s{st};
    const int rows = rows_of(ke.form);
    const size_t packed_doubles = (size_t)tiles_total * (rows * TILE + HEAD);
    AKB_CUDA(cudaMallocAsync(&packed_s.p, (packed_doubles + 8) * sizeof(double), st)); // + the source bounding box (FORM_ROWT)
    if (splits > 1) AKB_CUDA(cudaMallocAsync(&partial_s.p, (size_t)splits * M_out * 2 * sizeof(double), st));
    double *const packed = packed_s.p, *const partial = partial_s.p;

    PhaseConst pc = make_phase_const(k, ke.table);
    pc.planes = planes;
    unsigned long long *bbox = nullptr;
    if ((ke.form & FORM_ROWT) && mode == AKB_PHASE_REFERENCED) {
        bbox = reinterpret_cast<unsigned long long *>(packed + packed_doubles);
        AKB_CUDA(cudaMemsetAsync(bbox, 0xFF, 3 * sizeof(unsigned long long), st)); // minima start at the largest key
        AKB_CUDA(cudaMemsetAsync(bbox + 3, 0x00, 3 * sizeof(unsigned long long), st));
    }
    pack_sources_kernel<<<(unsigned)((padded + 255) / 256), 256, 0, st>>>(
        src_x, src_y, src_z, src_u, src_ds, N, padded, TILE, mode == AKB_PHASE_REFERENCED ? 1 : 0,
        (ke.form & FORM_POLAR) ? 1 : 0, rows, pc.im_sign, pc.inv_u, pc.neg_u_hi, pc.neg_u_lo, packed, bbox);
    AKB_LAUNCH_CHECK();

    double *dst = splits > 1 ? partial : out;
    if ((rc = timing_mark(1, st))) return rc;
    {
        long long M_ = M;
        int tt = tiles_total;
        const double *pk = packed;
        void *args[] = {(void *)&det_x, (void *)&det_y, (void *)&det_z, (void *)&M_, (void *)&pk, (void *)&tt,
                        (void *)&tiles_per_split, (void *)&n_padded, (void *)&pc, (void *)&dst};
        dim3 grid((unsigned)blocks_x, (unsigned)splits, (unsigned)groups);
        AKB_CUDA(cudaLaunchKernel(kern, grid, dim3(ke.threads), args, (size_t)ke.smem, st));
        count_launch();
    }
    if ((rc = timing_mark(2, st))) return rc;
    if (splits > 1) {
        reduce_partials_kernel<<<(unsigned)((M_out + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const double2 *>(partial), splits, M_out, reinterpret_cast<double2 *>(out));
        AKB_LAUNCH_CHECK();
    }
    if ((rc = timing_mark(3, st))) return rc;
    g_timing.valid = g_timing.enabled;
    return AKB_OK;
}

// (plane, pixel) -> flat detector arrays, for the phase mode the four-plane kernel is not built for
__global__ void expand_planes_kernel(const double *__restrict__ x_planes, const double *__restrict__ y, const double *__restrict__ z,
                                     long long M, int planes, double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * planes) return;
    const long long p = i / M, q = i - p * M;
    fx[i] = x_planes[p];
    fy[i] = y[q];
    fz[i] = z[q];
}

} // namespace

extern "C" int akb_fresnel_sum(const double *det_x, const double *det_y, const double *det_z, int64_t M,
                               const double *src_x, const double *src_y, const double *src_z,
                               const double *src_u, const double *src_ds, int64_t N, double k, double *out,
                               int mode, void *stream)
{
    return fresnel_sum_impl(det_x, det_y, det_z, M, 0, src_x, src_y, src_z, src_u, src_ds, N, k, out, mode, stream);
}

extern "C" int akb_fresnel_sum_planes(const double *det_y, const double *det_z, int64_t M, const double *x_planes, int planes,
                                      const double *src_x, const double *src_y, const double *src_z, const double *src_u,
                                      const double *src_ds, int64_t N, double k, double *out, int mode, void *stream)
{
    AKB_REQUIRE(planes >= 0 && M >= 0, "planes and M must be non-negative");
    if (planes == 0 || M == 0) return AKB_OK;
    AKB_REQUIRE(x_planes && det_y && det_z && out, "detector/out pointers must not be NULL");
#ifdef AKB_AB_VARIANTS
    if (mode != AKB_PHASE_REFERENCED && getenv("AKB_PLANES_KERNEL"))
        return fresnel_sum_impl(x_planes, det_y, det_z, M, planes, src_x, src_y, src_z, src_u, src_ds, N, k, out, mode, stream);
#endif
    // The plane-major flat detector set: every 1024-point block lies in one row of one plane and takes the planar-row
    // loop (REFERENCED: with the row expansion).  A kernel in which a thread keeps one pixel for four planes was built and
    // measured slower (see planes_kernel above).
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = (long long)M * planes;
    double *flat = nullptr;
    AKB_CUDA(cudaMallocAsync(&flat, 3 * (size_t)total * sizeof(double), st));
    expand_planes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x_planes, det_y, det_z, M, planes, flat, flat + total, flat + 2 * total);
    int rc = AKB_OK;
    if (cudaGetLastError() != cudaSuccess) {
        set_error("expand_planes_kernel launch failed");
        rc = AKB_ERR_CUDA;
    } else {
        count_launch();
        rc = fresnel_sum_impl(flat, flat + total, flat + 2 * total, total, 0, src_x, src_y, src_z, src_u, src_ds, N, k, out, mode, stream);
    }
    cudaFreeAsync(flat, st);
    return rc;
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

extern "C" int akb_fresnel_row_blocks(int64_t *row_blocks, int reset)
{
    unsigned long long v = 0;
    AKB_CUDA(cudaMemcpyFromSymbol(&v, g_row_blocks, sizeof(v))); // synchronises with the device
    if (row_blocks) *row_blocks = (int64_t)v;
    if (reset) {
        v = 0;
        AKB_CUDA(cudaMemcpyToSymbol(g_row_blocks, &v, sizeof(v)));
    }
    return AKB_OK;
}

extern "C" const char *akb_fresnel_variant_name(void) { return selected_kernel(AKB_PHASE_FAITHFUL).name; }

extern "C" int akb_fresnel_sum_host(const double *det_x, const double *det_y, const double *det_z, int64_t M,
                                    const double *src_x, const double *src_y, const double *src_z,
                                    const double *src_u, const double *src_ds, int64_t N, double k, double *out,
                                    int mode, int device)
{
    AKB_REQUIRE(M >= 0 && N >= 0, "M and N must be non-negative");
    if (M == 0) return AKB_OK;
    AKB_REQUIRE(det_x && det_y && det_z && out, "detector/out pointers must not be NULL");
    AKB_REQUIRE(N == 0 || (src_x && src_y && src_z && src_u), "source pointers must not be NULL");
    DeviceScope scope; // device < 0: the calling thread's current device; the caller's current device is restored on return
    device = scope.enter(device);
    if (device < 0) return AKB_ERR_CUDA;
    tune_pool(device);
    cudaStream_t st = host_stream(device);
    AKB_REQUIRE(st != nullptr, "could not create a stream on the device");
    const size_t mb = (size_t)M * sizeof(double), nb = (size_t)(N > 0 ? N : 1) * sizeof(double);
    double *d = nullptr;
    // one slab: out (2M) | u (2N) | det xyz (3M) | src xyz (3N) | ds (N); the two complex arrays
    // come first so that they stay 16-byte aligned for any M, N
    const size_t total = 5 * mb + 6 * nb;
    int rc = AKB_OK;
    cudaError_t e = cudaMallocAsync(&d, total, st);
    if (e != cudaSuccess) {
        set_error("cudaMallocAsync(%zu) failed: %s", total, cudaGetErrorString(e));
        return AKB_ERR_CUDA;
    }
    const int64_t Nn = N > 0 ? N : 1;
    double *dout = d, *du = d + 2 * M;
    double *ddx = du + 2 * Nn, *ddy = ddx + M, *ddz = ddy + M;
    double *dsx = ddz + M, *dsy = dsx + Nn, *dsz = dsy + Nn, *dds = dsz + Nn;
    // Small calls (C1: 4096 detector points x 1e4 sources, 0.5 MB) are dominated by per-copy latency: eight pageable
    // H2D copies cost more than the kernels.  Up to kStageBytes everything is gathered into ONE pinned staging
    // buffer of the calling thread (same layout as the device slab behind `du`) and moved with one copy each way.
    const size_t in_bytes = 3 * mb + (N > 0 ? 6 * nb : 0);
    char *stage = in_bytes + 2 * mb <= kStageBytes ? host_stage() : nullptr;
    if (stage) {
        char *w = stage;
        auto put = [&](const double *src, size_t bytes) {
            if (src) memcpy(w, src, bytes);
            w += bytes;
        };
        if (N > 0) put(src_u, 2 * nb); else w += 2 * nb;
        put(det_x, mb); put(det_y, mb); put(det_z, mb);
        if (N > 0) {
            put(src_x, nb); put(src_y, nb); put(src_z, nb);
            put(src_ds, nb);
        }
        if (cudaMemcpyAsync(du, stage, (size_t)(w - stage), cudaMemcpyHostToDevice, st) != cudaSuccess) {
            set_error("H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = AKB_ERR_CUDA;
        }
    } else {
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
    }
    if (rc == AKB_OK)
        rc = akb_fresnel_sum(ddx, ddy, ddz, M, dsx, dsy, dsz, du, src_ds ? dds : nullptr, N, k, dout, mode, st);
    // the result comes back through the staging buffer too (a pinned target keeps the copy asynchronous; the
    // inputs staged there have been consumed by the H2D copy, which precedes this one on the stream)
    double *back = stage ? reinterpret_cast<double *>(stage) : out;
    if (rc == AKB_OK && cudaMemcpyAsync(back, dout, 2 * mb, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
        set_error("D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = AKB_ERR_CUDA;
    }
    cudaFreeAsync(d, st);
    e = cudaStreamSynchronize(st);
    if (rc == AKB_OK && e != cudaSuccess) {
        set_error("stream synchronize failed: %s", cudaGetErrorString(e));
        rc = AKB_ERR_CUDA;
    }
    if (rc == AKB_OK && stage) memcpy(out, stage, 2 * mb);
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
