// Extreme eigenvalues of a FAMILY of kernel matrices K_b = variance * exp(r_b * G), b = 0..B-1, sharing one "log-Gram"
// G = log(K_ref / variance) <= 0 -- the device-side form of the reference's parameter-selection experiments
// (examples/kernels/sphere/sphere_gaussian_kernel_parameters.py:83-104, examples/kernels/spd/spd_gaussian_kernel_parameters.py:
// 108-117: for each beta of a grid, build the full Gram and take the minimum of numpy.linalg.eig).
//
// Every kernel of the path is variance * exp(-beta * g(x, y)) with g = d^2 (Gaussian kinds) or d (Laplace kinds), so the Gram
// of ANY beta is an elementwise power of the Gram of one reference beta0: K_beta = variance * (K_beta0 / variance)^(beta/beta0).
// The expensive pairwise geometry (acos, or a 6 x 6 generalised eigenproblem per pair) is therefore done ONCE per point set by
// the existing Gram kernels; this file never stores a second N x N matrix.  One CTA per (matrix, beta) runs Lanczos with full
// reorthogonalisation on the OPERATOR v -> K_b v, whose entries are re-formed on the fly (one lean FP64 exp per entry and
// step, gabo_math.cuh), then finds the extreme eigenvalues of its tridiagonal matrix by Sturm-count multisection.  With
// m = n steps (the default for the reference's n = 500) the tridiagonal matrix is an orthogonal similarity of K_b, i.e. the
// result has the accuracy of a dense symmetric eigensolver (O(eps * ||K||)); with m < n the smallest Ritz value is an upper
// bound of lambda_min and `resid` its residual estimate.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "gabo_common.cuh"
#include "gabo_math.cuh"

namespace gabo {
namespace {

constexpr int LZ_THREADS = 512;
constexpr int LZ_WARPS = LZ_THREADS / 32;

__global__ void log_ratio_kernel(const double* __restrict__ K, int64_t ldk, double* __restrict__ G, int64_t ldg, int64_t n1,
                                 int64_t n2, double inv_variance, double scale) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n2) return;
    for (int64_t i = blockIdx.y; i < n1; i += gridDim.y) {
        const double g = log(K[i * ldk + j] * inv_variance);
        G[i * ldg + j] = g > 0.0 ? 0.0 : scale * g;  // K <= variance for every kind of the path; rounding may leave +1 ulp
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA, same value (bitwise) in every thread.  `red` holds LZ_WARPS doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < LZ_WARPS; ++i) t += red[i];
    return t;
}

__device__ __forceinline__ double hash_uniform(unsigned long long s) {   // splitmix64 -> (-1, 1)
    s += 0x9E3779B97F4A7C15ull;
    s = (s ^ (s >> 30)) * 0xBF58476D1CE4E5B9ull;
    s = (s ^ (s >> 27)) * 0x94D049BB133111EBull;
    s ^= s >> 31;
    return (double)(long long)(s >> 11) * (1.0 / 4503599627370496.0) - 1.0;
}

// Number of eigenvalues of the symmetric tridiagonal (al[0..m), be[0..m-1)) that are < x (Sturm sequence of the LDL^T pivots).
__device__ __forceinline__ int sturm_count(const double* al, const double* be, int m, double x, double pivmin) {
    int cnt = 0;
    double q = al[0] - x;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += q < 0.0;
    for (int i = 1; i < m; ++i) {
        q = al[i] - x - be[i - 1] * be[i - 1] / q;
        if (fabs(q) < pivmin) q = -pivmin;
        cnt += q < 0.0;
    }
    return cnt;
}

// k-th smallest eigenvalue (k = 0: minimum, k = m-1: maximum) of the tridiagonal by LZ_THREADS-way multisection of [lo, hi].
// All threads return the same value.  `ired` holds one int.
__device__ double tridiag_kth_eig(const double* al, const double* be, int m, int k, double lo, double hi, double pivmin,
                                  int* ired) {
    for (int round = 0; round < 8; ++round) {
        const double w = (hi - lo) / (double)(LZ_THREADS + 1);
        const double x = lo + w * (double)(threadIdx.x + 1);
        const int c = sturm_count(al, be, m, x, pivmin);          // #eig < x, non-decreasing in threadIdx.x
        __syncthreads();
        if (threadIdx.x == 0) *ired = LZ_THREADS;                   // first thread whose point has count > k
        __syncthreads();
        if (c > k) atomicMin(ired, (int)threadIdx.x);
        __syncthreads();
        const int f = *ired;
        // eigenvalue k lies in [x_{f-1}, x_f]
        const double nlo = lo + w * (double)f, nhi = (f == LZ_THREADS) ? hi : lo + w * (double)(f + 1);
        lo = nlo;
        hi = nhi;
        if (!(hi - lo > 4.0 * 2.3e-16 * fmax(fabs(lo), fabs(hi)))) break;
    }
    return 0.5 * (lo + hi);
}

// One CTA = Lanczos for one (matrix t, ratio b).
//   G      [n_mat][n][ldg]  log-Gram (<= 0, -inf allowed)
//   Qws    [n_mat*nb][m_max + 1][n]  Lanczos vectors (workspace, L2 / HBM)
//   alpha/beta_out [n_mat*nb][m_max] the tridiagonal (diagnostic, may be NULL)
//   out    [n_mat*nb][4]: theta_min, theta_max, resid (|beta_m s_m| of the minimal Ritz pair; 0 when the Krylov space closed
//          or m == n), steps taken
__global__ void __launch_bounds__(LZ_THREADS, 1)
lanczos_sweep_kernel(const double* __restrict__ G, int64_t ldg, int64_t g_stride, int n, const double* __restrict__ ratios,
                     int nb, double variance, int m_max, double tol, double* __restrict__ Qws, double* __restrict__ alpha_out,
                     double* __restrict__ beta_out, double* __restrict__ out, unsigned long long seed) {
    extern __shared__ double sm[];
    const int npad = (n + 3) & ~3;
    double* v = sm;                      // current Lanczos vector q_k
    double* h = v + npad;                // projections (m_max + 1)
    double* al = h + (m_max + 1);        // diagonal of T
    double* be = al + m_max;             // off-diagonal of T
    double* tab = be + m_max;            // 2^(j/32)
    double* red = tab + 32;              // LZ_WARPS partials
    int* ired = reinterpret_cast<int*>(red + LZ_WARPS);

    const int b = blockIdx.x, t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double r = ratios[b];
    const double* Gm = G + (int64_t)t * g_stride;
    const int64_t slot = (int64_t)t * nb + b;
    double* Q = Qws + slot * (int64_t)(m_max + 1) * n;
    if (tid < 32) tab[tid] = kExp2Tab[tid];

    // q_0: pseudo-random, normalised
    double ss = 0.0;
    for (int i = tid; i < n; i += LZ_THREADS) {
        const double x = hash_uniform(seed + (unsigned long long)slot * 0x100000001B3ull + (unsigned long long)i);
        v[i] = x;
        ss += x * x;
    }
    ss = block_sum(ss, red);
    {
        const double s = rsqrt(ss);
        for (int i = tid; i < n; i += LZ_THREADS) {
            v[i] *= s;
            Q[i] = v[i];
        }
    }
    __syncthreads();

    int m = 0;
    bool closed = false;
    double anorm = 0.0;                  // running bound of ||T||
    for (int k = 0; k < m_max; ++k) {
        double* w = Q + (int64_t)(k + 1) * n;
        // ---- w = K_b q_k : one warp per row, four independent exps per lane in flight
        for (int i = warp; i < n; i += LZ_WARPS) {
            const double* g = Gm + (int64_t)i * ldg;
            double acc = 0.0;
            for (int j0 = 0; j0 < n; j0 += 128) {
                double x[4], e[4], vj[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 32 * u + lane;
                    const bool in = j < n;
                    x[u] = in ? r * g[j] : 0.0;
                    vj[u] = in ? v[j] : 0.0;
                }
                exp_neg_fast_n<4>(x, e, tab);
#pragma unroll
                for (int u = 0; u < 4; ++u) acc = fma(e[u], vj[u], acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) w[i] = variance * acc;
        }
        __syncthreads();
        // ---- full reorthogonalisation against q_0..q_k (classical Gram-Schmidt, repeated when it cancelled: DGKS criterion)
        double alpha = 0.0, nrm2_before = 0.0;
        for (int i = tid; i < n; i += LZ_THREADS) nrm2_before += w[i] * w[i];
        nrm2_before = block_sum(nrm2_before, red);
        double nrm2 = nrm2_before;
        for (int pass = 0; pass < 3; ++pass) {
            for (int j = warp; j <= k; j += LZ_WARPS) {
                const double* qj = Q + (int64_t)j * n;
                double d = 0.0;
                for (int i = lane; i < n; i += 32) d = fma(qj[i], w[i], d);
                d = warp_sum(d);
                if (lane == 0) h[j] = d;
            }
            __syncthreads();
            alpha += h[k];
            double s2 = 0.0;
            for (int i = tid; i < n; i += LZ_THREADS) {
                double wi = w[i];
                for (int j = 0; j <= k; ++j) wi = fma(-h[j], Q[(int64_t)j * n + i], wi);
                w[i] = wi;
                s2 += wi * wi;
            }
            const double after = block_sum(s2, red);     // (barriers inside also order the w updates before the next pass)
            const bool again = after < 0.5 * nrm2;       // lost more than 1/sqrt(2) of the norm: project once more
            nrm2 = after;
            if (!again) break;
        }
        const double beta = sqrt(nrm2);
        if (tid == 0) {
            al[k] = alpha;
            be[k] = beta;
        }
        m = k + 1;
        anorm = fmax(anorm, fabs(alpha) + beta + (k > 0 ? be[k - 1] : 0.0));
        // Krylov space closed (or all n directions used): T is exact
        if (m == n || !(beta > 64.0 * 2.3e-16 * anorm)) {
            closed = true;
            __syncthreads();
            break;
        }
        {
            const double s = 1.0 / beta;
            for (int i = tid; i < n; i += LZ_THREADS) {
                const double q = w[i] * s;
                w[i] = q;
                v[i] = q;
            }
        }
        __syncthreads();
        // optional early exit on the minimal Ritz pair (tol > 0): every 32 steps
        if (tol > 0.0 && (m & 31) == 0 && m < m_max) {
            double glo = al[0], ghi = al[0];
            for (int i = 0; i < m; ++i) {
                const double rad = (i > 0 ? fabs(be[i - 1]) : 0.0) + (i + 1 < m ? fabs(be[i]) : 0.0);
                glo = fmin(glo, al[i] - rad);
                ghi = fmax(ghi, al[i] + rad);
            }
            const double piv = 1e-300 + 2.3e-16 * 2.3e-16 * anorm * anorm;
            const double th = tridiag_kth_eig(al, be, m, 0, glo, ghi, piv, ired);
            // last component of the unit eigenvector of T_m for th by the three-term recurrence (rescaled)
            double s0 = 1.0, s1 = 0.0, nn = 1.0;
            s1 = -(al[0] - th) * s0 / be[0];
            nn += s1 * s1;
            for (int i = 1; i + 1 < m; ++i) {
                const double s2 = -((al[i] - th) * s1 + be[i - 1] * s0) / be[i];
                s0 = s1;
                s1 = s2;
                nn += s2 * s2;
                if (nn > 1e200) { s0 *= 1e-100; s1 *= 1e-100; nn *= 1e-200; }
            }
            const double resid = fabs(be[m - 1] * s1) * rsqrt(nn);
            if (resid <= tol * anorm) break;
        }
    }
    __syncthreads();
    // ---- extreme eigenvalues of T_m
    double glo = al[0], ghi = al[0];
    for (int i = 0; i < m; ++i) {
        const double rad = (i > 0 ? fabs(be[i - 1]) : 0.0) + (i + 1 < m ? fabs(be[i]) : 0.0);
        glo = fmin(glo, al[i] - rad);
        ghi = fmax(ghi, al[i] + rad);
    }
    const double tn = fmax(fabs(glo), fabs(ghi));
    const double piv = 1e-300 + 2.3e-16 * 2.3e-16 * tn * tn;
    const double thmin = tridiag_kth_eig(al, be, m, 0, glo, ghi, piv, ired);
    const double thmax = tridiag_kth_eig(al, be, m, m - 1, glo, ghi, piv, ired);
    double resid = 0.0;
    if (!closed && m > 1) {
        double s0 = 1.0, s1 = -(al[0] - thmin) / be[0], nn = 1.0 + s1 * s1;
        for (int i = 1; i + 1 < m; ++i) {
            const double s2 = -((al[i] - thmin) * s1 + be[i - 1] * s0) / be[i];
            s0 = s1;
            s1 = s2;
            nn += s2 * s2;
            if (nn > 1e200) { s0 *= 1e-100; s1 *= 1e-100; nn *= 1e-200; }
        }
        resid = fabs(be[m - 1] * s1) * rsqrt(nn);
    }
    if (tid == 0) {
        out[slot * 4 + 0] = thmin;
        out[slot * 4 + 1] = thmax;
        out[slot * 4 + 2] = resid;
        out[slot * 4 + 3] = (double)m;
    }
    if (alpha_out != nullptr)
        for (int i = tid; i < m_max; i += LZ_THREADS) {
            alpha_out[slot * m_max + i] = i < m ? al[i] : 0.0;
            beta_out[slot * m_max + i] = i < m ? be[i] : 0.0;
        }
}

inline size_t lanczos_smem_bytes(int n, int m_max) {
    const int npad = (n + 3) & ~3;
    return sizeof(double) * ((size_t)npad + (m_max + 1) + 2 * (size_t)m_max + 32 + LZ_WARPS + 2);
}

}  // namespace
}  // namespace gabo

// G = scale * min(0, log(K / variance)).  In place allowed (G == K, ldg == ldk).
extern "C" int gabo_log_gram_f64(const double* K, int64_t ldk, double* G, int64_t ldg, int64_t n1, int64_t n2, double variance,
                                 double scale, gabo_stream_t stream) {
    if (n1 < 0) return -5;
    if (n2 < 0) return -6;
    if (n1 == 0 || n2 == 0) return 0;
    if (K == nullptr) return -1;
    if (G == nullptr) return -3;
    if (ldk < n2) return -2;
    if (ldg < n2) return -4;
    if (!(variance > 0.0)) return -7;
    dim3 grid((unsigned)gabo::ceil_div64(n2, 256), (unsigned)(n1 < 32768 ? n1 : 32768));
    gabo::log_ratio_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(K, ldk, G, ldg, n1, n2, 1.0 / variance, scale);
    GABO_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t gabo_lanczos_sweep_workspace_bytes(int64_t n, int n_mat, int n_ratios, int m_max) {
    if (n <= 0 || n_mat <= 0 || n_ratios <= 0 || m_max <= 0) return 0;
    return sizeof(double) * (size_t)n_mat * (size_t)n_ratios * (size_t)(m_max + 1) * (size_t)n;
}

extern "C" int gabo_lanczos_sweep_f64(const double* G, int64_t ldg, int64_t g_stride, int64_t n, int n_mat, const double* ratios,
                                      int n_ratios, double variance, int m_max, double tol, unsigned long long seed, double* out,
                                      double* alpha_out, double* beta_out, void* workspace, size_t workspace_bytes,
                                      gabo_stream_t stream) {
    if (G == nullptr) return -1;
    if (n <= 0 || n > 16384) return -4;
    if (ldg < n) return -2;
    if (n_mat <= 0 || n_mat > 65535) return -5;
    if (n_mat > 1 && g_stride < n * ldg) return -3;
    if (ratios == nullptr) return -6;
    if (n_ratios <= 0) return -7;
    if (!(variance > 0.0)) return -8;
    if (m_max <= 0 || m_max > 2048) return -9;
    if (m_max > n) m_max = (int)n;
    if (out == nullptr) return -12;
    if ((alpha_out == nullptr) != (beta_out == nullptr)) return -13;
    if (workspace == nullptr || workspace_bytes < gabo_lanczos_sweep_workspace_bytes(n, n_mat, n_ratios, m_max)) return -15;
    const size_t smem = gabo::lanczos_smem_bytes((int)n, m_max);
    if (smem > 227 * 1024) return -9;
    GABO_CUDA_TRY(cudaFuncSetAttribute(gabo::lanczos_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)n_ratios, (unsigned)n_mat);
    gabo::lanczos_sweep_kernel<<<grid, gabo::LZ_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        G, ldg, g_stride, (int)n, ratios, n_ratios, variance, m_max, tol, static_cast<double*>(workspace), alpha_out, beta_out, out,
        seed);
    GABO_LAUNCH_CHECK();
    return 0;
}
