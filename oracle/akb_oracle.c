/*
 * akb_oracle.c -- CPU restatement of the two hot paths of Kakekakechan/AKBRaytracing.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under akbraytracing_b200/ may import, link or
 * call this file; it is used by tests/, by __graft_entry__.smoke() and by the
 * cpu_baseline / --impl reference legs of bench.py, as the checker and the CPU
 * baseline -- never as the product path.
 *
 * Parity pinning: the reference ships no golden vectors or tests.
 * This restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the
 * build container by tests/golden/make_golden.py (which imports the reference's
 * modules in place) and committed under tests/golden/ -- see tests/test_oracle_golden.py.
 *
 * All arithmetic is IEEE double, evaluated in the reference's operation order with
 * no fused multiply-add (build with -ffp-contract=off), because NumPy/numba never
 * fuse and the quadric coefficients are formed by cancellation.
 *
 * Citations: CPU0402 = Wavecalc_raytrace_fromData_CPU0402.py, ER3D = EllipseRaytrace3D.py,
 * BIG = AKB_raytrace_20250312.py (line numbers into /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ path A */

/*
 * Huygens-Fresnel pair sum, CPU0402:71-85 (compute_u_parallel) with the ds
 * pre-multiplication of CPU0402:102 (forward_propagation_numpy_batch):
 *
 *   w_j  = u_back_u[j] * ds[j]                       (complex * real)
 *   dist = sqrt((x_i-X_j)^2 + (y_i-Y_j)^2 + (z_i-Z_j)^2)
 *   u_i  = sum_j (1/dist) * exp(1j * (-k*dist)) * w_j
 *
 * x,y,z: M detector points; sx,sy,sz: N source points; u: N complex (re,im
 * interleaved); ds: N or NULL (=1).  out: M complex.  The reference sums j in
 * index order (numba lowers np.sum to a sequential loop); so do we.
 */
void orc_fresnel_sum(int64_t M, const double *x, const double *y, const double *z,
                     int64_t N, const double *sx, const double *sy, const double *sz,
                     const double *u, const double *ds, double k, double *out, int nthreads)
{
    double *w = (double *)malloc(sizeof(double) * 2 * (size_t)(N > 0 ? N : 1));
    for (int64_t j = 0; j < N; ++j) {
        double d = ds ? ds[j] : 1.0;
        w[2 * j] = u[2 * j] * d;
        w[2 * j + 1] = u[2 * j + 1] * d;
    }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < M; ++i) {
        const double xi = x[i], yi = y[i], zi = z[i];
        double acc_re = 0.0, acc_im = 0.0;
        for (int64_t j = 0; j < N; ++j) {
            double dx = xi - sx[j], dy = yi - sy[j], dz = zi - sz[j];
            double dist = sqrt((dx * dx + dy * dy) + dz * dz); /* CPU0402:76-80 */
            double amp = 1.0 / dist;                            /* CPU0402:81 */
            double phase = -k * dist;                           /* CPU0402:82 */
            double fr = amp * cos(phase), fi = amp * sin(phase); /* CPU0402:83 */
            acc_re += fr * w[2 * j] - fi * w[2 * j + 1];        /* CPU0402:84 */
            acc_im += fr * w[2 * j + 1] + fi * w[2 * j];
        }
        out[2 * i] = acc_re;
        out[2 * i + 1] = acc_im;
    }
    free(w);
}

/* ------------------------------------------------------------------ path B */
/* (3,N) row-major arrays: row 0 = x (or l), row 1 = y (m), row 2 = z (n). */

/*
 * ER3D:18-45 == BIG:445-471 mirr_ray_intersection, NumPy branch.
 * Returns the number of rays with not(D > 0).  Like the reference, when that count
 * is non-zero the WHOLE output is NaN (ER3D:31-33).
 */
int64_t orc_mirr_ray_intersection(const double *co, const double *ray, const double *src,
                                  int64_t N, int negative, double *point)
{
    const double a = co[0], b = co[1], c = co[2], d = co[3], e = co[4], f = co[5];
    const double g = co[6], h = co[7], i_ = co[8], j = co[9];
    const double *l = ray, *m = ray + N, *n = ray + 2 * N;
    const double *p = src, *q = src + N, *r = src + 2 * N;
    int64_t bad = 0;
    for (int64_t t_ = 0; t_ < N; ++t_) {
        double L = l[t_], Mm = m[t_], Nn = n[t_], P = p[t_], Q = q[t_], R = r[t_];
        /* ER3D:23 */
        double A = a * (L * L) + b * (Mm * Mm);
        A = A + c * (Nn * Nn);
        A = A + (d * Mm) * L;
        A = A + (e * Nn) * L;
        A = A + (f * Mm) * Nn;
        /* ER3D:24-26 */
        double B = ((2 * a) * P) * L + ((2 * b) * Q) * Mm;
        B = B + ((2 * c) * R) * Nn;
        B = B + d * (P * Mm + Q * L);
        B = B + e * (P * Nn + R * L);
        B = B + f * (R * Mm + Q * Nn);
        B = B + g * L;
        B = B + h * Mm;
        B = B + i_ * Nn;
        /* ER3D:27-28 */
        double C = a * (P * P) + b * (Q * Q);
        C = C + c * (R * R);
        C = C + (d * P) * Q;
        C = C + (e * P) * R;
        C = C + (f * Q) * R;
        C = C + g * P;
        C = C + h * Q;
        C = C + i_ * R;
        C = C + j;
        double D = B * B - (4 * A) * C; /* ER3D:30 */
        if (!(D > 0)) ++bad;
        double sq = sqrt(D);
        double t = negative ? (-B - sq) / (2 * A) : (-B + sq) / (2 * A); /* ER3D:35-38 */
        point[t_] = t * L + P;                                              /* ER3D:41-43 */
        point[N + t_] = t * Mm + Q;
        point[2 * N + t_] = t * Nn + R;
    }
    if (bad) {
        for (int64_t t_ = 0; t_ < 3 * N; ++t_) point[t_] = NAN;
    }
    return bad;
}

/*
 * ER3D:57-59 == BIG:530-532 normalize_vector: column-wise L2 normalise; if ANY
 * column norm is exactly 0 the input is returned unchanged.  In place.
 * norm = sqrt((v0^2 + v1^2) + v2^2)  (np.linalg.norm(axis=0): add.reduce over rows).
 * Returns the number of zero-norm columns.
 */
int64_t orc_normalize_vector(double *v, int64_t N)
{
    int64_t zeros = 0;
    for (int64_t t = 0; t < N; ++t) {
        double s = (v[t] * v[t] + v[N + t] * v[N + t]) + v[2 * N + t] * v[2 * N + t];
        if (sqrt(s) == 0.0) ++zeros; /* NaN != 0 is true in NumPy: NaN norms do not count */
    }
    if (zeros) return zeros;
    for (int64_t t = 0; t < N; ++t) {
        double s = (v[t] * v[t] + v[N + t] * v[N + t]) + v[2 * N + t] * v[2 * N + t];
        double nrm = sqrt(s);
        v[t] = v[t] / nrm;
        v[N + t] = v[N + t] / nrm;
        v[2 * N + t] = v[2 * N + t] / nrm;
    }
    return 0;
}

/* ER3D:61-71 == BIG:626-636 norm_vector: gradient of the quadric, normalised. */
int64_t orc_norm_vector(const double *co, const double *pt, int64_t N, double *nv)
{
    const double a = co[0], b = co[1], c = co[2], d = co[3], e = co[4], f = co[5];
    const double g = co[6], h = co[7], i_ = co[8];
    for (int64_t t = 0; t < N; ++t) {
        double x = pt[t], y = pt[N + t], z = pt[2 * N + t];
        nv[t] = (((2 * a) * x + d * y) + e * z) + g;          /* ER3D:66 */
        nv[N + t] = (((2 * b) * y + d * x) + f * z) + h;      /* ER3D:67 */
        nv[2 * N + t] = (((2 * c) * z + e * x) + f * y) + i_; /* ER3D:68 */
    }
    return orc_normalize_vector(nv, N); /* ER3D:70 */
}

/* ER3D:47-55 == BIG:502-509 reflect_ray: r - 2 (r.n) n, normalised. */
int64_t orc_reflect_ray(const double *ray, const double *nv, int64_t N, double *out)
{
    for (int64_t t = 0; t < N; ++t) {
        double l = ray[t], m = ray[N + t], n = ray[2 * N + t];
        double nx = nv[t], ny = nv[N + t], nz = nv[2 * N + t];
        double A = (l * nx + m * ny) + n * nz; /* ER3D:51 */
        double A2 = 2 * A;                     /* ER3D:52: ray - (2*A)*N */
        out[t] = l - A2 * nx;
        out[N + t] = m - A2 * ny;
        out[2 * N + t] = n - A2 * nz;
    }
    return orc_normalize_vector(out, N); /* ER3D:54 */
}

/* ER3D:145-157 == BIG:873-885 plane_ray_intersection (uses coeffs[6:10] only). */
void orc_plane_ray_intersection(const double *co, const double *ray, const double *src,
                                int64_t N, double *point)
{
    const double g = co[6], h = co[7], i_ = co[8], j = co[9];
    for (int64_t t_ = 0; t_ < N; ++t_) {
        double l = ray[t_], m = ray[N + t_], n = ray[2 * N + t_];
        double p = src[t_], q = src[N + t_], r = src[2 * N + t_];
        double num = ((g * p + h * q) + i_ * r) + j;
        double den = (g * l + h * m) + i_ * n;
        double t = -num / den; /* ER3D:150 */
        point[t_] = t * l + p;
        point[N + t_] = t * m + q;
        point[2 * N + t_] = t * n + r;
    }
}

/* Segment length ||b - a|| per ray, BIG:2884-2897 (np.linalg.norm(b - a, axis=0)). */
void orc_segment_length(const double *a, const double *b, int64_t N, double *out)
{
    for (int64_t t = 0; t < N; ++t) {
        double dx = b[t] - a[t], dy = b[N + t] - a[N + t], dz = b[2 * N + t] - a[2 * N + t];
        out[t] = sqrt((dx * dx + dy * dy) + dz * dz);
    }
}

/* ------------------------------------------------------- hand-off: calc_dS */

static double tri_area(const double *v0, const double *v1, const double *v2)
{
    double e1[3], e2[3];
    for (int c = 0; c < 3; ++c) {
        e1[c] = v1[c] - v0[c];
        e2[c] = v2[c] - v0[c];
    }
    double cx = e1[1] * e2[2] - e1[2] * e2[1];
    double cy = e1[2] * e2[0] - e1[0] * e2[2];
    double cz = e1[0] * e2[1] - e1[1] * e2[0];
    return sqrt((cx * cx + cy * cy) + cz * cz) / 2;
}

/*
 * BIG:13418-13473 calc_dS: area element of a (3, nV*nH) point cloud laid out as an
 * nV x nH grid: for interior points the sum of the four triangles (p,right,up),
 * (p,up,left), (p,left,down), (p,down,right); edges copy the nearest interior
 * value, corners the diagonal interior neighbour.  dS: nV*nH doubles.
 */
void orc_calc_dS(const double *pts, int64_t nV, int64_t nH, double *dS)
{
    const int64_t N = nV * nH;
    memset(dS, 0, sizeof(double) * (size_t)N);
#define PT(i, j, buf)                                      \
    do {                                                   \
        (buf)[0] = pts[(i) * nH + (j)];                    \
        (buf)[1] = pts[N + (i) * nH + (j)];                \
        (buf)[2] = pts[2 * N + (i) * nH + (j)];            \
    } while (0)
    for (int64_t i = 1; i < nV - 1; ++i)
        for (int64_t j = 1; j < nH - 1; ++j) {
            double p[3], pr[3], pl[3], pu[3], pd[3];
            PT(i, j, p); PT(i, j + 1, pr); PT(i, j - 1, pl); PT(i - 1, j, pu); PT(i + 1, j, pd);
            double s = 0.0;
            s += tri_area(p, pr, pu);
            s += tri_area(p, pu, pl);
            s += tri_area(p, pl, pd);
            s += tri_area(p, pd, pr);
            dS[i * nH + j] = s;
        }
#undef PT
    /* BIG:13456-13465: the elif chain visits rows first, so edge rows copy row 1 / nV-2 */
    for (int64_t i = 0; i < nV; ++i)
        for (int64_t j = 0; j < nH; ++j) {
            if (i == 0) dS[i * nH + j] = dS[1 * nH + j];
            else if (i == nV - 1) dS[i * nH + j] = dS[(nV - 2) * nH + j];
            else if (j == 0) dS[i * nH + j] = dS[i * nH + 1];
            else if (j == nH - 1) dS[i * nH + j] = dS[i * nH + (nH - 2)];
        }
    /* BIG:13468-13471 */
    dS[0] = dS[1 * nH + 1];
    dS[nH - 1] = dS[1 * nH + (nH - 2)];
    dS[(nV - 1) * nH] = dS[(nV - 2) * nH + 1];
    dS[(nV - 1) * nH + nH - 1] = dS[(nV - 2) * nH + (nH - 2)];
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
