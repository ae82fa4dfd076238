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

// Eigenvalues of a symmetric D x D matrix (upper storage of A, destroyed) by Householder tridiagonalisation followed by
// the implicit-shift QL iteration, eigenvalues only -- about 4x fewer FP64 pipe slots than cyclic Jacobi for d = 6.
// Everything is unrolled over compile-time indices so d[], e[] and A stay in registers; the active block always starts
// at the compile-time row l, and its data-dependent bottom m only predicates the unrolled sweep steps.
// Absolute accuracy ~ eps * ||A||; callers fall back to Jacobi (relative accuracy) for ill-conditioned matrices.
template <int D>
__device__ __forceinline__ void tridiag_ql_eigvals(double (&A)[D][D], double (&d)[D]) {
    double e[D];   // e[i] couples d[i] and d[i+1]; e[D-1] is a dummy slot
#pragma unroll
    for (int k = 0; k < D - 2; ++k) {
        double sigma = 0.0;
#pragma unroll
        for (int i = k + 1; i < D; ++i) sigma = fma(A[k][i], A[k][i], sigma);
        const double x0 = A[k][k + 1];
        double tail = 0.0;
#pragma unroll
        for (int i = k + 2; i < D; ++i) tail = fma(A[k][i], A[k][i], tail);
        d[k] = A[k][k];
        if (tail > 0.0) {
            const double nrm = sigma * rsqrt(sigma);
            const double alpha = -copysign(nrm, x0);
            double v[D];
#pragma unroll
            for (int i = k + 1; i < D; ++i) v[i] = A[k][i];
            v[k + 1] = x0 - alpha;
            const double beta = 1.0 / fma(-alpha, x0, sigma);   // 2 / (v.v), v.v = 2 (sigma - alpha x0)
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
                for (int j = i; j < D; ++j) A[i][j] -= fma(v[i], pv[j], pv[i] * v[j]);
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

    // Convergence tests on the exponents with INTEGER arithmetic (an FP64 compare occupies the FP64 pipe that bounds this
    // kernel): e is negligible against its neighbours when exponent(e) + 49 <= exponent(max(|d_i|, |d_i+1|)), i.e.
    // |e| < 2^-48 * max(...) ~ 3.6e-15 * max: neglected coupling far inside the 1e-12 budget.  (hi word >> 20 = exponent.)
    auto negligible = [](double ee, double da, double db) {
        const int he = __double2hiint(ee) & 0x7fffffff;
        const int ha = __double2hiint(da) & 0x7fffffff, hb = __double2hiint(db) & 0x7fffffff;
        const int hm = ha > hb ? ha : hb;
        return he + (49 << 20) <= hm;
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
            double dm = d[D - 1];
#pragma unroll
            for (int mm = l + 1; mm < D - 1; ++mm)
                if (m == mm) dm = d[mm];
            // Wilkinson shift: e^2 / (delta + sign(delta) sqrt(delta^2 + e^2)), delta = (d[l+1] - d[l]) / 2
            const double dl = 0.5 * (d[l + 1] - d[l]);
            const double e2 = e[l] * e[l];
            const double hh = fma(dl, dl, e2);
            double g = dm - d[l] + e2 / (dl + copysign(hh * rsqrt(hh), dl));
            double s = 1.0, c = 1.0, p = 0.0;
            bool aborted = false;
#pragma unroll
            for (int i = D - 2; i >= l; --i) {
                if (i < m && !aborted) {
                    const double f = s * e[i], b = c * e[i];
                    const double h2 = fma(f, f, g * g);
                    if (((__double2hiint(h2) & 0x7fffffff) | __double2loint(h2)) == 0) {
                        d[i + 1] -= p;
                        aborted = true;
                    } else {
                        const double rr = rsqrt(h2);
                        e[i + 1] = h2 * rr;
                        s = f * rr;
                        c = g * rr;
                        g = d[i + 1] - p;
                        const double r2 = fma(d[i] - g, s, 2.0 * c * b);
                        p = s * r2;
                        d[i + 1] = g + p;
                        g = fma(c, r2, -b);
                    }
                }
            }
            if (!aborted) {
                d[l] -= p;
                e[l] = g;
            }
#pragma unroll
            for (int mm = l + 1; mm < D; ++mm)
                if (m == mm) e[mm] = 0.0;
        }
    }
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
