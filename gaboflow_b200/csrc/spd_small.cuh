// In-register small-matrix routines for d x d SPD work (2 <= d <= 8): Mandel slots, Cholesky, triangular inverse,
// cyclic Jacobi.  Every loop is fully unrolled so that the arrays live in registers.
#pragma once
#include "gabo_common.cuh"

namespace gabo {

// Mandel slot of entry (r, c), r <= c, in the reference ordering (SPD_utils_tf.py:27,34-63): main diagonal first,
// then super-diagonal k = c - r = 1 .. d-1 in row order.
template <int D>
__host__ __device__ constexpr int mandel_slot(int r, int c) {
    return (c - r) * D - ((c - r) * (c - r - 1)) / 2 + r;
}

#define GABO_SYM(A, i, j) A[((i) < (j)) ? (i) : (j)][((i) < (j)) ? (j) : (i)]

// One cyclic Jacobi sweep bookkeeping: rotate rows/cols p,q of the symmetric matrix A (upper storage, A[i][j] i<=j)
// so that A[p][q] becomes zero.  Optionally accumulates the rotation into V (columns p,q).
template <int D, bool WITH_V>
__device__ __forceinline__ void jacobi_rotate(double (&A)[D][D], double (&V)[WITH_V ? D : 1][WITH_V ? D : 1],
                                              const int p, const int q) {
    const double apq = A[p][q];
    if (apq != 0.0) {
        const double app = A[p][p], aqq = A[q][q];
        const double dlt = aqq - app;
        const double h = sqrt(fma(dlt, dlt, 4.0 * apq * apq));
        const double t = (2.0 * apq) / (dlt + copysign(h, dlt));
        const double c = rsqrt(fma(t, t, 1.0));
        const double s = t * c;
        A[p][p] = fma(-t, apq, app);
        A[q][q] = fma(t, apq, aqq);
        A[p][q] = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (k != p && k != q) {
                const double akp = GABO_SYM(A, k, p), akq = GABO_SYM(A, k, q);
                GABO_SYM(A, k, p) = fma(c, akp, -s * akq);
                GABO_SYM(A, k, q) = fma(s, akp, c * akq);
            }
        }
        if (WITH_V) {
#pragma unroll
            for (int k = 0; k < (WITH_V ? D : 0); ++k) {
                const double vkp = V[k][p], vkq = V[k][q];
                V[k][p] = fma(c, vkp, -s * vkq);
                V[k][q] = fma(s, vkp, c * vkq);
            }
        }
    }
}

// Eigenvalues (on the diagonal of A on return) of a symmetric D x D matrix held in the upper triangle of A.
template <int D, bool WITH_V>
__device__ __forceinline__ void jacobi_eig(double (&A)[D][D], double (&V)[WITH_V ? D : 1][WITH_V ? D : 1]) {
    for (int sweep = 0; sweep < 16; ++sweep) {
        double off = 0.0, dg = 0.0;
#pragma unroll
        for (int p = 0; p < D; ++p) {
            dg = fma(A[p][p], A[p][p], dg);
#pragma unroll
            for (int q = p + 1; q < D; ++q) off = fma(A[p][q], A[p][q], off);
        }
        if (!(off > 1e-31 * dg)) break;   // also leaves on NaN
#pragma unroll
        for (int p = 0; p < D - 1; ++p)
#pragma unroll
            for (int q = p + 1; q < D; ++q) jacobi_rotate<D, WITH_V>(A, V, p, q);
    }
}

// 1/x to the last bit or two for finite, normal x: hardware seed + two Newton steps, no slow-path branch.
__device__ __forceinline__ double rcp_nr2(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    return fma(fma(-x, r, 1.0), r, r);
}

// ---- symmetric eigenvalues: Householder tridiagonalisation + implicit-shift QL, and the mixed-precision variant ---------
// Everything is unrolled over compile-time indices so d[], e[] and A stay in registers; the active QL block always starts
// at the compile-time row l, and its data-dependent bottom m only predicates the unrolled sweep steps.

// A (symmetric, upper storage, destroyed) -> tridiagonal (d, e); e[i] couples d[i] and d[i+1]; e[D-1] = 0 is a dummy slot.
template <int D>
__device__ __forceinline__ void householder_tridiag(double (&A)[D][D], double (&d)[D], double (&e)[D]) {
#pragma unroll
    for (int k = 0; k < D - 2; ++k) {
        const double x0 = A[k][k + 1];
        double tail = 0.0;
#pragma unroll
        for (int i = k + 2; i < D; ++i) tail = fma(A[k][i], A[k][i], tail);
        const double sigma = fma(x0, x0, tail);
        d[k] = A[k][k];
        if (tail > 0.0) {
            const double nrm = sigma * rsqrt(sigma);
            const double alpha = -copysign(nrm, x0);
            double v[D];
#pragma unroll
            for (int i = k + 1; i < D; ++i) v[i] = A[k][i];
            v[k + 1] = x0 - alpha;
            const double beta = rcp_nr2(fma(-alpha, x0, sigma));   // 2 / (v.v), v.v = 2 (sigma - alpha x0)
            double pv[D];
            double kk = 0.0;
#pragma unroll
            for (int i = k + 1; i < D; ++i) {
                double s = 0.0;
#pragma unroll
                for (int j = k + 1; j < D; ++j) s = fma(GABO_SYM(A, i, j), v[j], s);
                pv[i] = beta * s;
                kk = fma(pv[i], v[i], kk);
            }
            kk *= 0.5 * beta;
#pragma unroll
            for (int i = k + 1; i < D; ++i) pv[i] = fma(-kk, v[i], pv[i]);   // w
#pragma unroll
            for (int i = k + 1; i < D; ++i)
#pragma unroll
                for (int j = i; j < D; ++j) A[i][j] = fma(-pv[i], v[j], fma(-v[i], pv[j], A[i][j]));
            e[k] = alpha;
        } else {
            e[k] = x0;   // column already tridiagonal
        }
    }
    if (D >= 2) {
        d[D - 2] = A[D - 2][D - 2];
        e[D - 2] = A[D - 2][D - 1];
    }
    d[D - 1] = A[D - 1][D - 1];
    e[D - 1] = 0.0;
}

// Scalar-type glue of the QL iteration.  Convergence tests run on the exponents with INTEGER arithmetic (an FP64 compare
// occupies the FP64 pipe that bounds the SPD kernels): e is negligible against its neighbours when
// |e| < 2^-48 max(|d_i|, |d_i+1|) (double: 3.6e-15, far inside the 1e-12 budget) or 2^-22 max (float).
template <typename T> struct QlScalar;
template <> struct QlScalar<double> {
    static __device__ __forceinline__ double rsq(double x) { return rsqrt(x); }
    static __device__ __forceinline__ double fmad(double a, double b, double c) { return fma(a, b, c); }
    static __device__ __forceinline__ double cps(double a, double b) { return copysign(a, b); }
    static __device__ __forceinline__ double quot(double a, double b) { return a / b; }
    static __device__ __forceinline__ int mag(double x) { return __double2hiint(x) & 0x7fffffff; }
    static __device__ __forceinline__ bool is_zero(double x) { return ((__double2hiint(x) & 0x7fffffff) | __double2loint(x)) == 0; }
    static constexpr int kNegl = 49 << 20;
};
template <> struct QlScalar<float> {
    static __device__ __forceinline__ float rsq(float x) { return rsqrtf(x); }
    static __device__ __forceinline__ float fmad(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ float cps(float a, float b) { return copysignf(a, b); }
    static __device__ __forceinline__ float quot(float a, float b) { return __fdividef(a, b); }
    static __device__ __forceinline__ int mag(float x) { return __float_as_int(x) & 0x7fffffff; }
    static __device__ __forceinline__ bool is_zero(float x) { return (__float_as_int(x) & 0x7fffffff) == 0; }
    static constexpr int kNegl = 23 << 23;
};

// Eigenvalues of the symmetric tridiagonal (d, e), returned in d (unordered); e is destroyed.  Absolute accuracy
// ~ eps(T) * ||A||.
template <int D, typename T>
__device__ __forceinline__ void ql_tri_eigvals(T (&d)[D], T (&e)[D]) {
    using Q = QlScalar<T>;
    auto negligible = [](T ee, T da, T db) {
        const int he = Q::mag(ee), ha = Q::mag(da), hb = Q::mag(db);
        const int hm = ha > hb ? ha : hb;
        return he + Q::kNegl <= hm;
    };
#pragma unroll
    for (int l = 0; l < D - 1; ++l) {
        for (int iter = 0; iter < 40; ++iter) {
            // bottom of the unreduced block that starts at l
            int m = D - 1;
#pragma unroll
            for (int mm = D - 2; mm >= l; --mm)
                if (negligible(e[mm], d[mm], d[mm + 1])) m = mm;
            if (m == l) break;
            T dm = d[D - 1];
#pragma unroll
            for (int mm = l + 1; mm < D - 1; ++mm)
                if (m == mm) dm = d[mm];
            // Wilkinson shift: e^2 / (delta + sign(delta) sqrt(delta^2 + e^2)), delta = (d[l+1] - d[l]) / 2
            const T dl = T(0.5) * (d[l + 1] - d[l]);
            const T e2 = e[l] * e[l];
            const T hh = Q::fmad(dl, dl, e2);
            T g = dm - d[l] + Q::quot(e2, dl + Q::cps(hh * Q::rsq(hh), dl));
            T s = T(1), c = T(1), p = T(0);
            bool aborted = false;
#pragma unroll
            for (int i = D - 2; i >= l; --i) {
                if (i < m && !aborted) {
                    const T f = s * e[i], b = c * e[i];
                    const T h2 = Q::fmad(f, f, g * g);
                    if (Q::is_zero(h2)) {
                        d[i + 1] -= p;
                        aborted = true;
                    } else {
                        const T rr = Q::rsq(h2);
                        e[i + 1] = h2 * rr;
                        s = f * rr;
                        c = g * rr;
                        g = d[i + 1] - p;
                        const T r2 = Q::fmad(d[i] - g, s, T(2) * c * b);
                        p = s * r2;
                        d[i + 1] = g + p;
                        g = Q::fmad(c, r2, -b);
                    }
                }
            }
            if (!aborted) {
                d[l] -= p;
                e[l] = g;
            }
#pragma unroll
            for (int mm = l + 1; mm < D; ++mm)
                if (m == mm) e[mm] = T(0);
        }
    }
}

// FP32 QL for STARTING VALUES only (tridiag_eigvals_mixed validates what it is given, so this routine may fail): the
// sweep always runs over the whole block below l (no search for interior splits, no zero-rotation guard) and only e[l]
// is tested, against 2^-11 of its neighbours (the eigenvalue error is quadratic in the neglected coupling: ~1e-6); at
// most 6 iterations per eigenvalue.  About half the instructions of the
// guarded version above.
template <int D>
__device__ __forceinline__ void ql_tri_eigvals_start_f32(float (&d)[D], float (&e)[D]) {
#pragma unroll
    for (int l = 0; l < D - 1; ++l) {
        for (int iter = 0; iter < 6; ++iter) {
            const int he = __float_as_int(e[l]) & 0x7fffffff;
            const int ha = __float_as_int(d[l]) & 0x7fffffff, hb = __float_as_int(d[l + 1]) & 0x7fffffff;
            if (he + (16 << 23) <= (ha > hb ? ha : hb)) break;
            const float dl = 0.5f * (d[l + 1] - d[l]);
            const float e2 = e[l] * e[l];
            const float hh = fmaf(dl, dl, e2);
            float g = d[D - 1] - d[l] + __fdividef(e2, dl + copysignf(hh * rsqrtf(hh), dl));
            float s = 1.0f, c = 1.0f, p = 0.0f;
#pragma unroll
            for (int i = D - 2; i >= l; --i) {
                const float f = s * e[i], b = c * e[i];
                const float h2 = fmaf(f, f, g * g);
                const float rr = rsqrtf(h2);
                e[i + 1] = h2 * rr;
                s = f * rr;
                c = g * rr;
                g = d[i + 1] - p;
                const float r2 = fmaf(d[i] - g, s, 2.0f * c * b);
                p = s * r2;
                d[i + 1] = g + p;
                g = fmaf(c, r2, -b);
            }
            d[l] -= p;
            e[l] = g;
            e[D - 1] = 0.0f;
        }
    }
}

// Eigenvalues of a symmetric D x D matrix (upper storage of A, destroyed), all in FP64 -- about 4x fewer FP64 pipe slots
// than cyclic Jacobi for d = 6.  Callers fall back to Jacobi (relative accuracy) for ill-conditioned matrices.
template <int D>
__device__ __forceinline__ void tridiag_ql_eigvals(double (&A)[D][D], double (&d)[D]) {
    double e[D];
    householder_tridiag<D>(A, d, e);
    ql_tri_eigvals<D, double>(d, e);
}

// 1/x to ~2^-40 relative: hardware seed + one Newton step (the divides below only scale small corrections).
__device__ __forceinline__ double rcp_fast40(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return fma(fma(-x, r, 1.0), r, r);
}

// Mixed-precision eigenvalues of the symmetric positive-definite tridiagonal (d, e) (both kept): the data-dependent QL
// iteration runs in FP32 on the FP32 pipe -- it only has to land each eigenvalue inside the basin of its own root -- and
// FP64 does two Newton steps per eigenvalue on the characteristic polynomial, evaluated with its derivative by the
// three-term recurrence (no divisions, fixed trip count, three independent chains in flight).  Work is done on T / trace(T)
// so that neither precision can over- or underflow.  Validation: every second Newton correction must be >= 1000x smaller
// than the first (quadratic convergence -> remaining error <= 1e-6 of the second correction) and the eigenvalues must sum
// to the trace to 1e-13 (two starts that fell to the same root would miss it by their gap).  Returns false when either
// test fails (near-multiple eigenvalues, FP32 non-convergence, non-positive trace, NaN): the caller then runs the FP64 QL.
template <int D>
__device__ __forceinline__ bool tridiag_eigvals_mixed(const double (&d)[D], const double (&e)[D], double (&ev)[D]) {
    double tr = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) tr += d[k];
    if (!(tr > 0.0) || !(tr < 1e300)) return false;
    const double rt = rcp_nr2(tr);
    double ds[D], e2[D];
    float df[D], ef[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        ds[k] = d[k] * rt;
        const double es = e[k] * rt;
        e2[k] = es * es;
        df[k] = (float)ds[k];
        ef[k] = (float)es;
    }
    ql_tri_eigvals_start_f32<D>(df, ef);
    bool ok = true;
    double sum = 0.0;
    constexpr int G = 3;   // eigenvalues refined together
#pragma unroll
    for (int g0 = 0; g0 < D; g0 += G) {
        double x[G], d1[G];
#pragma unroll
        for (int u = 0; u < G; ++u) x[u] = (g0 + u < D) ? (double)df[g0 + u] : 0.0;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            double p1[G], p2[G], q1[G], q2[G];   // p_k, p_{k-1} and derivatives
#pragma unroll
            for (int u = 0; u < G; ++u) { p2[u] = 1.0; q2[u] = 0.0; p1[u] = ds[0] - x[u]; q1[u] = -1.0; }
#pragma unroll
            for (int k = 1; k < D; ++k) {
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    const double t = ds[k] - x[u];
                    const double pn = fma(t, p1[u], -e2[k - 1] * p2[u]);
                    const double qn = fma(t, q1[u], fma(-e2[k - 1], q2[u], -p1[u]));
                    p2[u] = p1[u]; q2[u] = q1[u]; p1[u] = pn; q1[u] = qn;
                }
            }
#pragma unroll
            for (int u = 0; u < G; ++u) {
                const double dlt = p1[u] * rcp_fast40(q1[u]);
                x[u] -= dlt;
                if (it == 0) d1[u] = fabs(dlt);
                else if (g0 + u < D) ok = ok && (fabs(dlt) <= fmax(1e-3 * d1[u], 1e-15));
            }
        }
#pragma unroll
        for (int u = 0; u < G; ++u)
            if (g0 + u < D) { ev[g0 + u] = x[u] * tr; sum += x[u]; }
    }
    return ok && (fabs(sum - 1.0) <= 1e-13);
}

// ---- packed FP32 pairs (sm_100a FFMA2 / FADD2: two FP32 operations per issue slot) --------------------------------------
// The affine-invariant pair kernel is ISSUE-bound, and its FP32 QL start is a serial dependency chain.  Running the chains
// of TWO pairs (two rows of the same column) in the two halves of 64-bit registers halves the FP32 instruction count per
// pair and doubles the instruction-level parallelism of that phase.  Only the reciprocal square roots stay scalar (MUFU),
// as bare rsqrt.approx.ftz / rcp.approx.ftz (rsqrtf() without -use_fast_math adds a denormal guard of four instructions).
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) {
    return ((f32x2_t)__float_as_uint(hi) << 32) | (f32x2_t)__float_as_uint(lo);
}
__device__ __forceinline__ float f2_lo(f32x2_t a) { return __uint_as_float((unsigned)a); }
__device__ __forceinline__ float f2_hi(f32x2_t a) { return __uint_as_float((unsigned)(a >> 32)); }
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) { f32x2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t f2_sub(f32x2_t a, f32x2_t b) { f32x2_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float rsq_approx_f32(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_approx_f32(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ f32x2_t f2_rsq(f32x2_t a) { return f2_pack(rsq_approx_f32(f2_lo(a)), rsq_approx_f32(f2_hi(a))); }
__device__ __forceinline__ f32x2_t f2_rcp(f32x2_t a) { return f2_pack(rcp_approx_f32(f2_lo(a)), rcp_approx_f32(f2_hi(a))); }
// |a| with the sign of b, per half
__device__ __forceinline__ f32x2_t f2_copysign(f32x2_t a, f32x2_t b) {
    return (a & 0x7fffffff7fffffffull) | (b & 0x8000000080000000ull);
}
// |e| <= 2^-16 (da + db) in both halves.  da, db are diagonal entries of an SPD tridiagonal (positive, and they stay positive
// under the orthogonal similarity transforms of QL), so no absolute values are needed on them.  NaN halves count as "converged" (their lanes fail the FP64 validation anyway).
__device__ __forceinline__ bool f2_negligible16(f32x2_t e, f32x2_t da, f32x2_t db) {
    const f32x2_t t = f2_sub(f2_mul(e & 0x7fffffff7fffffffull, f2_pack(65536.0f, 65536.0f)), f2_add(da, db));
    return !(f2_lo(t) > 0.0f || f2_hi(t) > 0.0f);
}

// ql_tri_eigvals_start_f32 for two tridiagonals at once (low / high halves).  Both run the same number of iterations (the
// maximum of the two); an extra implicit-shift sweep on an already converged block is a valid similarity transform.
template <int D>
__device__ __forceinline__ void ql_tri_eigvals_start_f32x2(f32x2_t (&d)[D], f32x2_t (&e)[D]) {
    const f32x2_t one = f2_pack(1.0f, 1.0f), half = f2_pack(0.5f, 0.5f);
#pragma unroll
    for (int l = 0; l < D - 1; ++l) {
#pragma unroll 1   // rolled: six unrolled copies per l made a 70 KB loop body, beyond the 32 KB L1.5 instruction cache
        for (int iter = 0; iter < 6; ++iter) {
            if (f2_negligible16(e[l], d[l], d[l + 1])) break;
            const f32x2_t dl = f2_mul(f2_sub(d[l + 1], d[l]), half);
            const f32x2_t e2 = f2_mul(e[l], e[l]);
            const f32x2_t hh = f2_fma(dl, dl, e2);
            const f32x2_t rt = f2_mul(hh, f2_rsq(hh));
            const f32x2_t den = f2_add(dl, f2_copysign(rt, dl));
            f32x2_t g = f2_add(f2_sub(d[D - 1], d[l]), f2_mul(e2, f2_rcp(den)));
            f32x2_t s = one, c = one, p = 0ull;
#pragma unroll
            for (int i = D - 2; i >= l; --i) {
                const f32x2_t f = f2_mul(s, e[i]), b = f2_mul(c, e[i]);
                const f32x2_t h2 = f2_fma(f, f, f2_mul(g, g));
                const f32x2_t rr = f2_rsq(h2);
                e[i + 1] = f2_mul(h2, rr);
                s = f2_mul(f, rr);
                c = f2_mul(g, rr);
                g = f2_sub(d[i + 1], p);
                const f32x2_t r2 = f2_fma(f2_sub(d[i], g), s, f2_mul(f2_add(c, c), b));
                p = f2_mul(s, r2);
                d[i + 1] = f2_add(g, p);
                g = f2_sub(f2_mul(c, r2), b);
            }
            d[l] = f2_sub(d[l], p);
            e[l] = g;
            e[D - 1] = 0ull;
        }
    }
}

// The FP64 half of tridiag_eigvals_mixed on its own: two Newton steps per eigenvalue of the trace-normalised tridiagonal
// (ds diagonal, e2 squared off-diagonals) from the starting values x0, with the same validation.  ev = eigenvalues / trace.
template <int D>
__device__ __forceinline__ bool tridiag_newton_refine(const double (&ds)[D], const double (&e2)[D], const float (&x0)[D],
                                                      double (&ev)[D]) {
    bool ok = true;
    double sum = 0.0;
    constexpr int G = 3;
#pragma unroll
    for (int g0 = 0; g0 < D; g0 += G) {
        double x[G], d1[G];
#pragma unroll
        for (int u = 0; u < G; ++u) { x[u] = (g0 + u < D) ? (double)x0[g0 + u] : 0.0; d1[u] = 0.0; }
#pragma unroll 1   // rolled to keep the caller's loop body inside the instruction cache
        for (int it = 0; it < 2; ++it) {
            double p1[G], p2[G], q1[G], q2[G];
#pragma unroll
            for (int u = 0; u < G; ++u) { p2[u] = 1.0; q2[u] = 0.0; p1[u] = ds[0] - x[u]; q1[u] = -1.0; }
#pragma unroll
            for (int k = 1; k < D; ++k) {
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    const double t = ds[k] - x[u];
                    const double pn = fma(t, p1[u], -e2[k - 1] * p2[u]);
                    const double qn = fma(t, q1[u], fma(-e2[k - 1], q2[u], -p1[u]));
                    p2[u] = p1[u]; q2[u] = q1[u]; p1[u] = pn; q1[u] = qn;
                }
            }
#pragma unroll
            for (int u = 0; u < G; ++u) {
                const double dlt = p1[u] * rcp_fast40(q1[u]);
                x[u] -= dlt;
                if (it == 0) d1[u] = fabs(dlt);
                else if (g0 + u < D) ok = ok && (fabs(dlt) <= fmax(1e-3 * d1[u], 1e-15));
            }
        }
#pragma unroll
        for (int u = 0; u < G; ++u)
            if (g0 + u < D) { ev[g0 + u] = x[u]; sum += x[u]; }
    }
    return ok && (fabs(sum - 1.0) <= 1e-13);
}

// Lower Cholesky factor of the symmetric matrix S (upper storage read), L lower (L[i][j], j <= i).
// Returns 0, or k (1-based) if pivot k is not positive.
template <int D>
__device__ __forceinline__ int chol_lower(const double (&S)[D][D], double (&L)[D][D]) {
    int info = 0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double d = S[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d = fma(-L[j][k], L[j][k], d);
        if (!(d > 0.0) && info == 0) info = j + 1;
        const double r = rsqrt(d);
        L[j][j] = d * r;
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
            double v = S[j][i];
#pragma unroll
            for (int k = 0; k < j; ++k) v = fma(-L[i][k], L[j][k], v);
            L[i][j] = v * r;
        }
    }
    return info;
}

// W = L^-1 for lower-triangular L (W lower).
template <int D>
__device__ __forceinline__ void tri_inverse_lower(const double (&L)[D][D], double (&W)[D][D]) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
        W[j][j] = 1.0 / L[j][j];
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
            double v = 0.0;
#pragma unroll
            for (int k = j; k < i; ++k) v = fma(L[i][k], W[k][j], v);
            W[i][j] = -v / L[i][i];
        }
    }
}

}  // namespace gabo
