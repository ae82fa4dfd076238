// Batched small-N Gaussian-process evaluation: log marginal likelihood AND its hyper-parameter gradients for B hyper-parameter
// sets in ONE launch, one CTA per set, everything in shared memory.
//
// Why: configs 1-3 of BASELINE.json (examples/kernels/sphere/sphere_kernels.py:125-139, examples/bo_sphere/benchmark_examples/
// bo_sphere_ackley_manifold.py:137-160, examples/bo_spd/benchmark_examples/bo_spd_ackley_manifold.py:159-187) run at N = 5..100+:
// the reference's GPR.optimize evaluates GPflow 0.5's build_likelihood [third party] (Gram -> Cholesky -> solve -> log-density)
// and its TensorFlow gradient hundreds of times per BO iteration.  At that size a chain of tiled kernels is pure launch latency,
// so the whole evaluation is fused, and the unit of parallelism is the hyper-parameter set (multi-start / grid / line-search
// candidates), not the matrix tile.
//
// Every kernel of the path is variance * exp(-beta g(x, x')) (eig_sweep.cu), so all B Gram matrices are elementwise powers of ONE
// log-Gram G = log(K_beta0 / variance0) (gabo_log_gram_f64): Ky_b = v_b exp(r_b G) + s2_b I, r_b = beta_b / beta0.
//
// One loop over the columns j, one CTA barrier per column, does three things at once (augmented Cholesky):
//   rows 0..n-1        : L         (left-looking column Cholesky; every thread re-derives the pivot from the broadcast row j)
//   rows n..n+p-1      : V = L^-1 (Y - c)      (the rows of [Ky ; (Y-c)^T] below the square)
//   rows n+p..2n+p-1   : L^-1                  (the rows of [Ky ; I] below the square; column c of L^-1 is kept in ROW c of the
//                                               unused upper triangle, its diagonal 1/L_cc in rdiag)
// then alpha = L^-T V, tr(Ky^-1) = ||L^-1||_F^2 and sum_{i>j} (alpha alpha^T - p Ky^-1)_ij K_ij G_ij by dot products over rows of
// that upper triangle.  The row stride is odd, so "thread = row" accesses are bank-conflict free and the pivot row is a
// broadcast.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "gabo_common.cuh"

namespace gabo {
namespace {

constexpr int GPS_MAX_N = GABO_GP_SMALL_MAX_N;
constexpr int GPS_MAX_P = GABO_GP_SMALL_MAX_P;
constexpr double kLog2Pi = 1.8378770664093454835606594728112;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA, identical in every thread (fixed order).  red: 32 doubles.
__device__ __forceinline__ double cta_sum(double v, double* red) {
    v = warp_sum_d(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}

struct GpSmallParams {
    const double* G; int64_t ldg;
    const double* Y; int64_t ldy;
    const double* theta;          // [B][4]: ratio, variance, sigma2, mean constant
    double* out;                  // [B][5]: loglik, d/d ratio, d/d variance, d/d sigma2, d/d mean constant
    int32_t* info;                // [B]
    double* Lout; int64_t ldl, l_stride;   // optional: lower factor per set
    double* Vout;                 // optional: [B][n][p]
    int n, p, want_grad;
};

__global__ void gp_small_batch_kernel(const GpSmallParams P) {
    extern __shared__ __align__(16) unsigned char gps_smem[];
    const int n = P.n, p = P.p;
    const int ld = (n | 1);                       // odd row stride (in doubles)
    double* const A = reinterpret_cast<double*>(gps_smem);      // [n + p][ld]: L below, (L^-1)^T above, V in the rows n..n+p-1
    double* const rdiag = A + (size_t)(n + p) * ld;             // [n] 1 / L_cc
    double* const dl = rdiag + n;                               // [n] L_cc (moved onto the diagonal of A after the loop)
    double* const alpha = dl + n;                               // [p][n]
    double* const red = alpha + (size_t)p * n;                  // [32]

    const int b = blockIdx.x;
    const int t = threadIdx.x;
    const double ratio = P.theta[4 * b + 0], var = P.theta[4 * b + 1], s2 = P.theta[4 * b + 2], cmean = P.theta[4 * b + 3];

    // ---- Ky (lower) and the right-hand sides -------------------------------------------------------------------------
    for (int idx = t; idx < n * n; idx += blockDim.x) {
        const int i = idx / n, j = idx - i * n;
        if (j < i) A[i * ld + j] = var * exp(ratio * P.G[(int64_t)i * P.ldg + j]);
        else if (j == i) A[i * ld + j] = var + s2;
    }
    for (int idx = t; idx < n * p; idx += blockDim.x) {
        const int j = idx / p, c = idx - j * p;
        A[(n + c) * ld + j] = P.Y[(int64_t)j * P.ldy + c] - cmean;
    }
    __syncthreads();

    // ---- augmented left-looking Cholesky --------------------------------------------------------------------------------
    // thread roles: t < n: matrix row t;  n <= t < n+p: right-hand side t-n;  n+p <= t < 2n+p: column t-n-p of L^-1
    const bool is_row = t < n, is_rhs = (t >= n) && (t < n + p), is_inv = P.want_grad && (t >= n + p) && (t < 2 * n + p);
    const int ic = t - n - p;                                    // inverse column handled by an is_inv thread
    const double* myrow = is_inv ? (A + ic * ld) : (A + t * ld);   // rows of A addressed by this thread's dot products
    int info = 0;
    for (int j = 0; j < n; ++j) {
        const double* rj = A + j * ld;
        // pivot (every thread, same order => bitwise the same value) and this thread's own dot product.  Two plain loops --
        // pivot only up to ks, pivot + own row from ks on -- so that both unroll with their loads in flight (entries k <= ic of
        // an inverse column are zero / rdiag, and a thread with nothing to produce for this column only needs the pivot)
        double piv0 = 0.0, piv1 = 0.0, d0 = 0.0, d1 = 0.0;
        const bool active = (is_row && t >= j) || is_rhs || (is_inv && ic < j);
        int ks = active ? (is_inv ? ic + 1 : 0) : j;
        if (ks > j) ks = j;
        int k = 0;
#pragma unroll 4
        for (; k + 1 < ks; k += 2) {
            const double a0 = rj[k], a1 = rj[k + 1];
            piv0 = fma(a0, a0, piv0);
            piv1 = fma(a1, a1, piv1);
        }
        if (k < ks) {                                            // odd prefix: one pivot term, keeps the pairing below aligned
            const double a0 = rj[k];
            piv0 = fma(a0, a0, piv0);
            ++k;
            if (k < j) {
                const double a1 = rj[k];
                piv1 = fma(a1, a1, piv1);
                d1 = fma(myrow[k], a1, d1);
                ++k;
            }
        }
#pragma unroll 4
        for (; k + 1 < j; k += 2) {
            const double a0 = rj[k], a1 = rj[k + 1];
            const double m0 = myrow[k], m1 = myrow[k + 1];
            piv0 = fma(a0, a0, piv0);
            piv1 = fma(a1, a1, piv1);
            d0 = fma(m0, a0, d0);
            d1 = fma(m1, a1, d1);
        }
        if (k < j) {
            const double a0 = rj[k];
            piv0 = fma(a0, a0, piv0);
            d0 = fma(myrow[k], a0, d0);
        }
        const double piv = rj[j] - (piv0 + piv1);
        if (!(piv > 0.0)) { info = j + 1; break; }               // uniform: every thread computed the same pivot
        const double rdj = rsqrt(piv), dj = piv * rdj;
        double res = 0.0;
        bool write = false;
        if (is_row && t > j) { res = (myrow[j] - (d0 + d1)) * rdj; write = true; }
        else if (is_rhs) { res = (myrow[j] - (d0 + d1)) * rdj; write = true; }
        else if (is_inv && ic < j) { res = -(fma(rdiag[ic], rj[ic], d0 + d1)) * rdj; write = true; }
        // nothing written here is read by another thread during this iteration (A[j][j] keeps the input value: the factor's
        // diagonal goes to dl), so one barrier per column is enough
        if (write) {
            if (is_inv) A[ic * ld + j] = res;                    // (L^-1)[j][ic], kept at row ic of the upper triangle
            else A[t * ld + j] = res;
        }
        if (t == j) { dl[j] = dj; rdiag[j] = rdj; }
        __syncthreads();
    }
    if (info == 0 && is_row) A[t * ld + t] = dl[t];
    __syncthreads();
    if (info != 0) {
        if (t == 0) {
            P.info[b] = info;
            const double nanv = __longlong_as_double(0x7ff8000000000000ll);
            for (int q = 0; q < 5; ++q) P.out[5 * b + q] = nanv;
        }
        return;
    }

    // ---- log marginal likelihood ------------------------------------------------------------------------------------
    double part = 0.0;
    if (is_row) part = -(double)p * log(A[t * ld + t]);
    if (is_rhs) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) s = fma(myrow[j], myrow[j], s);
        part = -0.5 * s;
    }
    const double ll = cta_sum(part, red) - 0.5 * (double)n * (double)p * kLog2Pi;
    double vtv = 0.0;
    {
        double s = 0.0;
        if (is_rhs) for (int j = 0; j < n; ++j) s = fma(myrow[j], myrow[j], s);
        vtv = cta_sum(s, red);
    }
    if (P.Lout != nullptr) {
        double* Lb = P.Lout + (int64_t)b * P.l_stride;
        for (int idx = t; idx < n * n; idx += blockDim.x) {
            const int i = idx / n, j = idx - i * n;
            if (j <= i) Lb[(int64_t)i * P.ldl + j] = A[i * ld + j];
        }
    }
    if (P.Vout != nullptr) {
        double* Vb = P.Vout + (int64_t)b * n * p;
        for (int idx = t; idx < n * p; idx += blockDim.x) {
            const int j = idx / p, c = idx - j * p;
            Vb[idx] = A[(n + c) * ld + j];
        }
    }
    if (!P.want_grad) {
        if (t == 0) {
            P.info[b] = 0;
            P.out[5 * b + 0] = ll;
            const double nanv = __longlong_as_double(0x7ff8000000000000ll);
            for (int q = 1; q < 5; ++q) P.out[5 * b + q] = nanv;
        }
        return;
    }

    // ---- alpha = L^-T V, tr(Ky^-1) ------------------------------------------------------------------------------------
    double tr_part = 0.0, aa_part = 0.0, asum_part = 0.0;
    if (is_row) {
        const double rd = rdiag[t];
        double s = rd * rd;
        for (int i = t + 1; i < n; ++i) { const double x = myrow[i]; s = fma(x, x, s); }
        tr_part = s;
        for (int c = 0; c < p; ++c) {
            const double* Vc = A + (n + c) * ld;
            double a = rd * Vc[t];
            for (int i = t + 1; i < n; ++i) a = fma(myrow[i], Vc[i], a);
            alpha[c * n + t] = a;
            aa_part = fma(a, a, aa_part);
            asum_part += a;
        }
    }
    const double tr_kinv = cta_sum(tr_part, red);
    const double aa = cta_sum(aa_part, red);
    const double asum = cta_sum(asum_part, red);   // (also orders the alpha writes before the reads below)

    // ---- S = sum_{i>j} (sum_c alpha_ic alpha_jc - p Kinv_ij) K_ij G_ij ;  dll/dratio = S ---------------------------------
    // thread pair (i, half): row i, columns j of one parity; Kinv_ij = rdiag_i Linv[i][j] + sum_{k>i} Linv[k][i] Linv[k][j],
    // with Linv[k][c] at A[c][k]; k runs downwards from n-1 so that lanes (consecutive i) share k: A[j][k] is a broadcast
    double s_part = 0.0;
    {
        const int i = t % n, half = t / n;
        if (half < 2 && t < 2 * n) {
            const double* ri = A + i * ld;
            const double rdi = rdiag[i];
            for (int j = half; j < i; j += 2) {
                const double* rjj = A + j * ld;
                double k0a = rdi * rjj[i], k1a = 0.0;
                int k = n - 1;
                for (; k - 1 > i; k -= 2) { k0a = fma(ri[k], rjj[k], k0a); k1a = fma(ri[k - 1], rjj[k - 1], k1a); }
                if (k > i) k0a = fma(ri[k], rjj[k], k0a);
                const double kinv = k0a + k1a;
                double aat = 0.0;
                for (int c = 0; c < p; ++c) aat = fma(alpha[c * n + i], alpha[c * n + j], aat);
                const double g = P.G[(int64_t)j * P.ldg + i];      // G is symmetric: lanes (consecutive i) read one row
                const double kij = var * exp(ratio * g);
                s_part = fma((aat - (double)p * kinv) * kij, g, s_part);
            }
        }
    }
    const double S = cta_sum(s_part, red);
    if (t == 0) {
        P.info[b] = 0;
        P.out[5 * b + 0] = ll;
        P.out[5 * b + 1] = S;
        P.out[5 * b + 2] = 0.5 * ((vtv - s2 * aa) - (double)p * ((double)n - s2 * tr_kinv)) / var;
        P.out[5 * b + 3] = 0.5 * (aa - (double)p * tr_kinv);
        P.out[5 * b + 4] = asum;
    }
}

}  // namespace
}  // namespace gabo

extern "C" int gabo_gp_small_batch_f64(int64_t n, int64_t p, const double* G, int64_t ldg, const double* Y, int64_t ldy, int64_t B,
                                       const double* theta, int want_grad, double* out, int32_t* info, double* Lout, int64_t ldl,
                                       int64_t l_stride, double* Vout, gabo_stream_t stream) {
    using namespace gabo;
    if (n < 1 || n > GPS_MAX_N) return -1;
    if (p < 1 || p > GPS_MAX_P) return -2;
    if (G == nullptr) return -3;
    if (ldg < n) return -4;
    if (Y == nullptr) return -5;
    if (ldy < p) return -6;
    if (B < 0) return -7;
    if (B == 0) return 0;
    if (theta == nullptr) return -8;
    if (out == nullptr) return -10;
    if (info == nullptr) return -11;
    if (Lout != nullptr && (ldl < n || (B > 1 && l_stride < n * ldl))) return -13;
    GpSmallParams P;
    P.G = G; P.ldg = ldg; P.Y = Y; P.ldy = ldy; P.theta = theta; P.out = out; P.info = info;
    P.Lout = Lout; P.ldl = ldl; P.l_stride = l_stride; P.Vout = Vout;
    P.n = (int)n; P.p = (int)p; P.want_grad = want_grad != 0;
    const int ld = ((int)n | 1);
    const size_t smem = ((size_t)(n + p) * ld + 2 * (size_t)n + (size_t)p * n + 32) * sizeof(double);
    if (smem > 227 * 1024) return -2;
    int threads = (int)(2 * n + p);
    if (!want_grad) threads = (int)(n + p);
    threads = ((threads + 31) / 32) * 32;
    if (threads < 64) threads = 64;
    static_assert(2 * GPS_MAX_N + GPS_MAX_P <= 1024, "one thread per augmented row");
    GABO_CUDA_TRY(cudaFuncSetAttribute(gp_small_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gp_small_batch_kernel<<<(unsigned)B, threads, smem, (cudaStream_t)stream>>>(P);
    GABO_LAUNCH_CHECK();
    return 0;
}
