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
