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

// One implicit-shift QL iteration (Wilkinson shift, sweep i = D-2 .. l) on the block [l, D-1] of two tridiagonals held in the
// low / high halves of d, e.  FP32, STARTING VALUES only: no search for interior splits, no zero-rotation guard -- the FP64
// refinement validates what it is given.
template <int D, int L>
__device__ __forceinline__ void ql_sweep_f32x2(f32x2_t (&d)[D], f32x2_t (&e)[D]) {
    const f32x2_t one = f2_pack(1.0f, 1.0f), half = f2_pack(0.5f, 0.5f);
    const f32x2_t dl = f2_mul(f2_sub(d[L + 1], d[L]), half);
    const f32x2_t e2 = f2_mul(e[L], e[L]);
    const f32x2_t hh = f2_fma(dl, dl, e2);
    const f32x2_t rt = f2_mul(hh, f2_rsq(hh));
    const f32x2_t den = f2_add(dl, f2_copysign(rt, dl));
    f32x2_t g = f2_add(f2_sub(d[D - 1], d[L]), f2_mul(e2, f2_rcp(den)));
    f32x2_t s = one, c = one, p = 0ull;
#pragma unroll
    for (int i = D - 2; i >= L; --i) {
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
    d[L] = f2_sub(d[L], p);
    e[L] = g;
    e[D - 1] = 0ull;
}

template <int D, int L>
struct QlLevels {
    // one packed pair
    static __device__ __forceinline__ void run(f32x2_t (&da)[D], f32x2_t (&ea)[D]) {
#pragma unroll 1   // rolled: six unrolled copies per level made a 70 KB loop body, beyond the 32 KB L1.5 instruction cache
        for (int iter = 0; iter < 6; ++iter) {
            if (f2_negligible16(ea[L], da[L], da[L + 1])) break;
            ql_sweep_f32x2<D, L>(da, ea);
        }
        QlLevels<D, L + 1>::run(da, ea);
    }
    // two packed pairs as two independent dependency chains in one instruction stream
    static __device__ __forceinline__ void run(f32x2_t (&da)[D], f32x2_t (&ea)[D], f32x2_t (&db)[D], f32x2_t (&eb)[D]) {
#pragma unroll 1
        for (int iter = 0; iter < 6; ++iter) {
            if (f2_negligible16(ea[L], da[L], da[L + 1]) && f2_negligible16(eb[L], db[L], db[L + 1])) break;
            ql_sweep_f32x2<D, L>(da, ea);
            ql_sweep_f32x2<D, L>(db, eb);
        }
        QlLevels<D, L + 1>::run(da, ea, db, eb);
    }
};
template <int D>
struct QlLevels<D, D - 1> {
    static __device__ __forceinline__ void run(f32x2_t (&)[D], f32x2_t (&)[D]) {}
    static __device__ __forceinline__ void run(f32x2_t (&)[D], f32x2_t (&)[D], f32x2_t (&)[D], f32x2_t (&)[D]) {}
};

// Eigenvalue STARTS of two tridiagonals (low / high halves), levels l = 0 .. D-2, at most 6 iterations per level.  Both
// halves run the same number of iterations (the maximum of the two); an extra implicit-shift sweep on an already converged
// block is a valid similarity transform.
template <int D>
__device__ __forceinline__ void ql_tri_eigvals_start_f32x2(f32x2_t (&d)[D], f32x2_t (&e)[D]) {
    QlLevels<D, 0>::run(d, e);
}
// The same for FOUR tridiagonals as two packed pairs that advance together (twice the instruction-level parallelism, twice
// the code; measured no faster than one pair per trip at four CTAs per SM, kept for reference).
template <int D>
__device__ __forceinline__ void ql_tri_eigvals_start_f32x2_dual(f32x2_t (&da)[D], f32x2_t (&ea)[D], f32x2_t (&db)[D],
                                                                f32x2_t (&eb)[D]) {
    QlLevels<D, 0>::run(da, ea, db, eb);
}

// Mixed-precision eigenvalues, FP64 half: refinement of the starting values x0 on the characteristic polynomial of the
// trace-normalised tridiagonal (ds diagonal, e2 squared off-diagonals), evaluated by the three-term recurrence (no
// divisions, fixed trip count, three independent chains in flight).  Step 1 is a Newton step (p and p'); step 2 re-evaluates
// p only and reuses 1/p' of step 1 (the chord step: 3 instead of 5 FP64 operations per recurrence row and no second
// reciprocal).  With e0 the error of the start and kappa = p''/p', the corrections are dlt1 ~ e0, dlt2 ~ kappa e0^2 / 2 and
// the error left after step 2 is kappa^2 e0^3 / 2 ~ 2 dlt2^2 / dlt1, which is what gets validated: every eigenvalue must
// satisfy dlt2^2 <= 1e-15 |dlt1| (error <= 2e-15 of the trace), and the eigenvalues must sum to the trace to 1e-13 (two
// starts that fell to the same root miss it by their gap).  Returns false otherwise (near-multiple eigenvalues, FP32
// non-convergence, NaN): the caller then takes the all-FP64 slow path.  ev = eigenvalues / trace.
// Measured on 4e5 pairs of the C5 workload (tools/spd_newton_emul.py): 6e-5 of the pairs fail the validation; the
// accepted eigenvalues are within 2e-14 relative of LAPACK's.
template <int D>
__device__ __forceinline__ bool tridiag_newton_refine(const double (&ds)[D], const double (&e2)[D], const float (&x0)[D],
                                                      double (&ev)[D]) {
    bool ok = true;
    double sum = 0.0;
    constexpr int G = 3;
#pragma unroll
    for (int g0 = 0; g0 < D; g0 += G) {
        double x[G], d1[G], rq[G];
#pragma unroll
        for (int u = 0; u < G; ++u) x[u] = (g0 + u < D) ? (double)x0[g0 + u] : 0.0;
        {   // step 1: p and p'
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
                rq[u] = rcp_fast40(q1[u]);
                const double dlt = p1[u] * rq[u];
                x[u] -= dlt;
                d1[u] = fabs(dlt);
            }
        }
        {   // step 2: p only, chord step with the derivative of step 1
            double p1[G], p2[G];
#pragma unroll
            for (int u = 0; u < G; ++u) { p2[u] = 1.0; p1[u] = ds[0] - x[u]; }
#pragma unroll
            for (int k = 1; k < D; ++k) {
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    const double t = ds[k] - x[u];
                    const double pn = fma(t, p1[u], -e2[k - 1] * p2[u]);
                    p2[u] = p1[u]; p1[u] = pn;
                }
            }
#pragma unroll
            for (int u = 0; u < G; ++u) {
                const double dlt = p1[u] * rq[u];
                x[u] -= dlt;
                if (g0 + u < D) ok = ok && (dlt * dlt <= 1e-15 * d1[u]);
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
