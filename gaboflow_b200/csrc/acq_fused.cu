// Fused acquisition for candidate batches on the sphere: for each of M candidates ONE CTA does
//     k = K(X, x*)  ->  a = L^-1 k  ->  mean = a.V + m,  var = Kdiag - a.a  ->  Expected Improvement  (->  Euclidean gradients)
// in one launch, reading the cached inverse factor L^-1 (lower triangular, formed once per factorisation) instead of running
// Gram + TRSM + predict + EI (+ TRSM^T + gradient) as five to seven launches with fresh intermediates per call.
//
// This is the inner loop of the reference's acquisition optimisation: MultistartedOptimizer scores 100-1000 anchor points per BO
// iteration (BO_utils/manifold_optimization.py:402-431, `objective(100 pts)` :427) and the manifold conjugate gradient then asks
// for (-EI, -dEI/dx) one candidate at a time (:250-271) -- each call a fresh TF session run through GPR.build_predict and
// GPflowOpt's ExpectedImprovement [third party].  Training sets there are tens to hundreds of points: launch-latency territory.
//
// Arithmetic is the reference's: A = L^-1 K(X,X*), mean = A^T V + m, var = Kdiag(X*) - colsum(A^2) (GPR.build_predict), EI with
// the variance floored at 1e-6 (GPflowOpt), gradients as in gabo_sphere_posterior_grad_f64.
#include "gabo_common.cuh"

namespace gabo {
namespace {

constexpr int AQ_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                                   // red may still be read from the previous call
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < AQ_THREADS / 32; ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(AQ_THREADS) sphere_acq_fused_kernel(int kind, int64_t n, int dim, const double* __restrict__ X,
                                                                      int64_t ldx, const double* __restrict__ Xc, int64_t ldc,
                                                                      const double* __restrict__ Linv, int64_t ldl,
                                                                      const double* __restrict__ V, const double* __restrict__ alpha,
                                                                      double beta, double variance, double mean_const, double f_min,
                                                                      double* __restrict__ mean_out, double* __restrict__ var_out,
                                                                      double* __restrict__ ei_out, double* __restrict__ gmean,
                                                                      double* __restrict__ gvar, double* __restrict__ dei) {
    extern __shared__ double sm_acq[];
    double* sk = sm_acq;              // k_i, later reused for c_i alpha_i
    double* sa = sm_acq + n;          // a = L^-1 k
    double* sc = sm_acq + 2 * n;      // gradient coefficients c_i, later c_i b_i
    __shared__ double sxc[64];
    __shared__ double red[AQ_THREADS / 32];
    __shared__ double gred[4][64][2];
    const int64_t m = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < dim) sxc[tid] = Xc[m * ldc + tid];
    __syncthreads();
    // 1. cross-Gram column and the gradient coefficients  c_i = k_i beta p d^(p-1) / sin d
    for (int64_t i = tid; i < n; i += AQ_THREADS) {
        double t = 0.0;
        for (int j = 0; j < dim; ++j) t = fma(X[i * ldx + j], sxc[j], t);
        t = fmin(fmax(t, -1.0), 1.0);                                   // sphere_utils_tf.py:59
        const double d = acos(t);                                       // sphere_utils_tf.py:60
        double k, c;
        if (kind == GABO_SPHERE_GAUSSIAN) {
            k = variance * exp(-beta * d * d);                          // kernels_sphere_tf.py:83-88
            c = 2.0 * beta * k * ((d > 1e-8) ? d / fmax(sin(d), 1e-12) : 1.0);
        } else {
            k = variance * exp(-beta * d);                              // kernels_sphere_tf.py:199-204
            c = beta * k / fmax(sin(d), 1e-12);
        }
        sk[i] = k;
        sc[i] = c;
    }
    __syncthreads();
    // 2. a = L^-1 k : warp per row, lanes along the row (coalesced), rows dealt to the 8 warps
    for (int64_t r = warp; r < n; r += AQ_THREADS / 32) {
        double s = 0.0;
        const double* row = Linv + r * ldl;
        for (int64_t j = lane; j <= r; j += 32) s = fma(row[j], sk[j], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sa[r] = s;
    }
    __syncthreads();
    // 3. posterior mean / variance (GPR.build_predict), 4. Expected Improvement (GPflowOpt)
    double pm = 0.0, pv = 0.0;
    for (int64_t i = tid; i < n; i += AQ_THREADS) {
        pm = fma(sa[i], V[i], pm);
        pv = fma(sa[i], sa[i], pv);
    }
    const double mu = block_sum(pm, red) + mean_const;
    const double var = variance - block_sum(pv, red);
    const double vf = fmax(var, 1e-6);
    const double sd = sqrt(vf);
    const double z = (f_min - mu) / sd;
    const double cdf = 0.5 * erfc(-z * 0.70710678118654752440);
    const double phi = exp(-0.5 * z * z) * 0.39894228040143267794;
    if (tid == 0) {
        mean_out[m] = mu;
        var_out[m] = var;
        ei_out[m] = (f_min - mu) * cdf + vf * (phi / sd);
    }
    if (gmean == nullptr) return;
    // 5. b = L^-T a = Ky^-1 k : thread per column (coalesced along the row for every r), then the weighted sums over the points
    __syncthreads();
    for (int64_t j = tid; j < n; j += AQ_THREADS) {
        double s = 0.0;
        for (int64_t r = j; r < n; ++r) s = fma(Linv[r * ldl + j], sa[r], s);
        const double c = sc[j];
        sc[j] = c * s;                   // c_j b_j
        sk[j] = c * alpha[j];            // c_j alpha_j   (k itself is no longer needed)
    }
    __syncthreads();
    const int j = tid & 63, grp = tid >> 6;
    double a1 = 0.0, a2 = 0.0;
    if (j < dim) {
        for (int64_t i = grp; i < n; i += 4) {
            const double x = X[i * ldx + j];
            a1 = fma(sk[i], x, a1);
            a2 = fma(sc[i], x, a2);
        }
    }
    gred[grp][j][0] = a1;
    gred[grp][j][1] = a2;
    __syncthreads();
    if (grp == 0 && j < dim) {
        const double g1 = (gred[0][j][0] + gred[1][j][0]) + (gred[2][j][0] + gred[3][j][0]);
        const double g2 = -2.0 * ((gred[0][j][1] + gred[1][j][1]) + (gred[2][j][1] + gred[3][j][1]));
        gmean[m * dim + j] = g1;
        gvar[m * dim + j] = g2;
        if (dei != nullptr) dei[m * dim + j] = -cdf * g1 + ((var > 1e-6) ? phi / (2.0 * sd) * g2 : 0.0);
    }
}

}  // namespace
}  // namespace gabo

// Sphere kernels.  X [n, dim] training points, Xc [M, dim] candidates, Linv [n, n] = L^-1 (lower, row-major), V [n] = L^-1 (y - m),
// alpha [n] = Ky^-1 (y - m) (only read when gradients are requested).  Outputs mean / var / ei [M]; gmean / gvar / dei [M, dim]
// (all three NULL: no gradients; dei alone may be NULL).  n <= 9000 (three n-vectors in shared memory), dim <= 64.
extern "C" int gabo_sphere_acq_fused_f64(int kind, int64_t n, int64_t M, int dim, const double* X, int64_t ldx, const double* Xc,
                                         int64_t ldc, const double* Linv, int64_t ldl, const double* V, const double* alpha,
                                         double beta, double variance, double mean_const, double fmin, double* mean, double* var,
                                         double* ei, double* gmean, double* gvar, double* dei, gabo_stream_t stream) {
    using namespace gabo;
    if (kind != GABO_SPHERE_GAUSSIAN && kind != GABO_SPHERE_LAPLACE) return -1;
    if (n < 1 || n > 9000) return -2;
    if (M < 0) return -3;
    if (dim < 1 || dim > 64) return -4;
    if (M == 0) return 0;
    if (X == nullptr) return -5;
    if (ldx < dim) return -6;
    if (Xc == nullptr) return -7;
    if (ldc < dim) return -8;
    if (Linv == nullptr) return -9;
    if (ldl < n) return -10;
    if (V == nullptr) return -11;
    if (!(beta > 0.0)) return -13;
    if (!(variance > 0.0)) return -14;
    if (mean == nullptr) return -17;
    if (var == nullptr) return -18;
    if (ei == nullptr) return -19;
    if ((gmean == nullptr) != (gvar == nullptr)) return -20;
    if (gmean != nullptr && alpha == nullptr) return -12;
    if (gmean == nullptr && dei != nullptr) return -22;
    const size_t smem = (size_t)3 * n * sizeof(double);
    static size_t attr_smem = 48 * 1024;
    if (smem > attr_smem) {
        GABO_CUDA_TRY(cudaFuncSetAttribute(sphere_acq_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    sphere_acq_fused_kernel<<<(unsigned)M, AQ_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        kind, n, dim, X, ldx, Xc, ldc, Linv, ldl, V, alpha, beta, variance, mean_const, fmin, mean, var, ei, gmean, gvar, dei);
    GABO_LAUNCH_CHECK();
    return 0;
}
