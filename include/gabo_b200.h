/* gabo_b200.h -- C ABI of the B200-native (sm_100a) GaBOflow hot path.
 *
 * The reference (NoemieJaquier/GaBOflow) has no FFI: its boundary for this path is the Python
 * kernel-class contract (K / Kdiag on a gpflow.kernels.Kern subclass) plus GPflow 0.5's GPR
 * linear algebra.  Each entry point below names the reference code it replaces (paths relative
 * to the reference root).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - every pointer is a caller-owned DEVICE pointer unless the name ends in _host;
 *   - matrices are row-major (NumPy / torch default), leading dimension in ELEMENTS;
 *   - points are rows; SPD inputs arrive as Mandel vectors [n, d(d+1)/2]
 *     (SPD_utils_tf.py:11-65 ordering: main diagonal, then super-diagonals 1..d-1, off-diagonal
 *     slots scaled by sqrt(2));
 *   - no internal DEVICE-MEMORY allocation: scratch is passed in, sized by the matching *_workspace_bytes();
 *   - stream is a cudaStream_t passed as void*; calls are asynchronous on that stream and thread-safe.
 *     Process-wide state, all of it mutex- or atomic-guarded: the diagnostic launch counter, the optional
 *     trailing-update profiler (gabo_profile_trailing_*), and a per-device pool of one high-priority side
 *     stream + two events per concurrent gabo_potrf_f64 call (created on first use, reused afterwards, so
 *     that call creates no CUDA objects in steady state and can be captured into a CUDA graph);
 *   - return value: 0 ok; -k = argument k (1-based) is invalid; GABO_ERR_CUDA-e = CUDA runtime
 *     error e at launch.  Numerical failures (non-SPD input, failed Cholesky pivot) are reported
 *     LAPACK-style through device-side info words, never by aborting (the reference's TF ops raise
 *     InvalidArgumentError on a failed Cholesky and GPflowOpt retries with more jitter).
 *   - there is no CPU fallback anywhere behind this ABI.
 */
#ifndef GABO_B200_H
#define GABO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GABO_ABI_VERSION 2
#define GABO_ERR_CUDA (-1000)
#define GABO_GP_SMALL_MAX_N 160 /* gabo_gp_small_batch_f64: training points per model */
#define GABO_GP_SMALL_MAX_P 8   /* ... and output columns */
#define GABO_MAX_PEERS 8      /* ranks of one NVSwitch node that exchange through peer memory */

typedef void* gabo_stream_t; /* cudaStream_t */

/* Point-set Gram kinds (one fused tile kernel; csrc/points_gram.cu). */
enum gabo_points_kind {
    GABO_SPHERE_GAUSSIAN = 0, /* variance*exp(-beta*acos(clamp(x.y))^2)  kernels_sphere_tf.py:59-88 + sphere_utils_tf.py:11-60 */
    GABO_SPHERE_LAPLACE = 1,  /* variance*exp(-beta*acos(clamp(x.y)))    kernels_sphere_tf.py:179-204 */
    GABO_SQDIST_GAUSSIAN = 2  /* variance*exp(-beta*||x-y||^2): SpdFrobeniusGaussianKernel.K (kernels_spd_tf.py:495-533)
                                 on the Mandel vectors, SpdLogEuclideanGaussianKernel.K (:603-641) on Mandel(logm) */
};

/* SPD pair kinds (csrc/spd_gram.cu). */
enum gabo_spd_kind {
    GABO_SPD_AI_GAUSSIAN = 0, /* variance*exp(-beta*d_AI^2)   kernels_spd_tf.py:214-248 + SPD_utils_tf.py:95-139 */
    GABO_SPD_AI_LAPLACE = 1,  /* variance*exp(-beta*d_AI)     kernels_spd_tf.py:368-400 */
    GABO_SPD_STEIN = 2        /* variance*det(A B)^beta/det((A+B)/2)^beta, as coded at kernels_spd_tf.py:98-112 */
};

/* Which part of a symmetric Gram (X2 == NULL) is produced. */
enum gabo_uplo {
    GABO_FULL = 0,        /* every entry computed (what K(X) returns in the reference) */
    GABO_LOWER = 1,       /* tiles on/below the diagonal only (input of gabo_potrf_f64); upper part untouched */
    GABO_SYMM_MIRROR = 2  /* lower tiles computed once and mirror-stored: full, exactly symmetric matrix; needs row0==0,row1==n */
};

int gabo_abi_version(void);
/* Diagnostic only: number of CUDA kernels this library has launched in this process so far. */
long long gabo_launch_count(void);
/* Fills SM count and compute capability of the current device. */
int gabo_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Static, human-readable text for a status code returned by any call below. */
const char* gabo_status_string(int status);

/* ---- Point-set Gram: sphere Gaussian / Laplace, squared-distance Gaussian -------------------------
 * Replaces sphere_distance_tf (sphere_utils_tf.py:11-60) + SphereGaussianKernel.K / SphereLaplaceKernel.K
 * (kernels_sphere_tf.py:59-88, 179-204) and the pair stage of the Frobenius / Log-Euclidean kernels.
 * K[(i-row0)*ldk + j] = k(X[i], X2[j]) for row0 <= i < row1, 0 <= j < n2.  X2 == NULL means X2 = X
 * (kernels_sphere_tf.py:75-76); in that case the diagonal is exactly `variance`.
 * row0 must be a multiple of 128 (row-block sharding across GPUs uses 128-row granularity). */
size_t gabo_points_gram_workspace_bytes(int64_t n1, int64_t n2, int dim);
int gabo_points_gram_f64(int kind, const double* X, int64_t n1, int64_t ldx,
                         const double* X2, int64_t n2, int64_t ldx2, int dim,
                         double beta, double variance,
                         double* K, int64_t ldk, int64_t row0, int64_t row1, int uplo,
                         void* workspace, size_t workspace_bytes, gabo_stream_t stream);
/* FP32 mode of the same kernel (north_star: 1e-5 relative). */
int gabo_points_gram_f32(int kind, const float* X, int64_t n1, int64_t ldx,
                         const float* X2, int64_t n2, int64_t ldx2, int dim,
                         float beta, float variance,
                         float* K, int64_t ldk, int64_t row0, int64_t row1, int uplo,
                         void* workspace, size_t workspace_bytes, gabo_stream_t stream);

/* Kdiag = fill([n], variance)   (kernels_sphere_tf.py:105,221; kernels_spd_tf.py:265,417,550,658). */
int gabo_kdiag_const_f64(int64_t n, double variance, double* out, gabo_stream_t stream);

/* ---- SPD per-point preparation ---------------------------------------------------------------------
 * Replaces vector_to_symmetric_matrix_tf (SPD_utils_tf.py:11-65) and hoists the per-PAIR inverse / sqrtm /
 * logm of the reference (SPD_utils_tf.py:129, kernels_spd_tf.py:634) to once per point:
 *   S[i]      d x d symmetric matrix of Mandel row i                    (out_S, [n,d*d], may be NULL)
 *   W[i]      inverse of the lower Cholesky factor L[i] of S[i]         (out_W, [n,d*d], may be NULL)
 *   logdet[i] log det S[i]                                              (out_logdet, [n], may be NULL)
 *   logm[i]   Mandel vector of the matrix logarithm (Jacobi eigen)      (out_logm_mandel, [n,m], may be NULL)
 *   Lt        the factors L[i], packed lower (entry (a,k), k <= a, at a(a+1)/2 + k) and TRANSPOSED:
 *             Lt[(a(a+1)/2 + k)*ldt + i], ldt >= n -- the column operand of the affine-invariant pair kernel, laid
 *             out so that a warp reads one entry of 32 consecutive points in one 256 B request (out_Lt, may be NULL)
 *   info[i]   0, or k>0 if the k-th Cholesky pivot of S[i] is not positive (outputs for i are NaN)
 * 2 <= d <= 8. */
int gabo_spd_prep_f64(const double* Xmandel, int64_t n, int64_t ldx, int d,
                      double* out_S, double* out_W, double* out_logdet, double* out_logm_mandel,
                      double* out_Lt, int64_t ldt, int32_t* info, gabo_stream_t stream);

/* ---- SPD pair Gram -----------------------------------------------------------------------------------
 * Affine-invariant kinds replace affine_invariant_distance_tf (SPD_utils_tf.py:95-139):
 *   d_AI(i,j)^2 = sum_k log^2 lambda_k( W1[i] S2[j] W1[i]^T ), formed as (W1[i] L2[j])(W1[i] L2[j])^T from the triangular
 *   factors; eigenvalues by Householder + FP32 QL starts + validated FP64 Newton refinement (FP64 QL / cyclic Jacobi
 *   when the refinement does not validate or the spectrum is wide).
 * Stein replaces kernels_spd_tf.py:98-112: exp(beta*(logdet1[i]+logdet2[j]-logdet((S1[i]+S2[j])/2))).
 * W1 / S1 / logdet1 describe X (rows), S2 / L2t / logdet2 describe X2 (columns); `symmetric` != 0 states that
 * both describe the same point set (enables uplo modes and the exact diagonal for the AI kinds).
 * Unused inputs for a kind may be NULL (AI: S1, S2, logdet1, logdet2; Stein: W1, L2t). */
int gabo_spd_gram_f64(int kind, int d,
                      const double* W1, const double* S1, const double* logdet1, int64_t n1,
                      const double* S2, const double* L2t, int64_t ldt2, const double* logdet2, int64_t n2,
                      int symmetric, double beta, double variance,
                      double* K, int64_t ldk, int64_t row0, int64_t row1, int uplo,
                      gabo_stream_t stream);
/* gabo_spd_gram_cyclic_f64: the symmetric SPD Gram K(X, X) of the kinds above, distributed over the GPUs of one node by
 *   512-row panels, 1-D block-cyclic (panel P on rank P % world, local slot P / world -- the layout the distributed
 *   Cholesky consumes).  Every rank passes the same per-point data (replicated: 4.7 MB at N = 16384, d = 6); rank `rank`
 *   evaluates its round-robin share of the lower 32 x 128 blocks -- so the triangular work is balanced whatever the panel
 *   ownership -- and stores each block AND its mirror image straight into the HBM of the ranks that own those rows
 *   (K_bases_host: HOST array of `world` device pointers, entry r = base of rank r's [local_rows, ldk] buffer as mapped in
 *   this process: own memory or CUDA IPC peer mappings, i.e. NVLink stores issued from the epilogue; no collective).
 *   Callers synchronise (stream + barrier, or a flag handshake) before reading their rows. */
int gabo_spd_gram_cyclic_f64(int kind, int d, const double* W1, const double* S1, const double* logdet1,
                             const double* L1t, int64_t ldt, int64_t n, double beta, double variance,
                             double* const* K_bases_host, int64_t ldk, int world, int rank, gabo_stream_t stream);
/* SpdSteinGaussianKernel.Kdiag = diag_part(K(X)) (kernels_spd_tf.py:129) = variance*exp(beta*logdet[i]). */
int gabo_spd_stein_kdiag_f64(const double* logdet, int64_t n, double beta, double variance,
                             double* out, gabo_stream_t stream);

/* ---- GP linear algebra (GPflow 0.5 GPR.build_likelihood / build_predict; call sites
 *      examples/kernels/sphere/sphere_kernels.py:131-139, examples/kernels/spd/spd_kernels.py:146-170) ----
 * gabo_potrf_f64: A <- chol(A + sigma2*I), lower triangle of the row-major n x n matrix, in place
 *   (tf.cholesky of K + sigma^2 I).  Blocked right-looking; trailing updates on the FP64 DMMA pipe.
 *   *info (device int32) = 0, or k>0 if the leading minor of order k is not positive definite.
 *   The strictly upper triangle is neither read nor written. */
size_t gabo_potrf_workspace_bytes(int64_t n);
int gabo_potrf_f64(int64_t n, double* A, int64_t lda, double sigma2, int32_t* info,
                   void* workspace, size_t workspace_bytes, gabo_stream_t stream);
/* gabo_trsm_f64: B <- L^-1 B (trans == 0) or L^-T B (trans != 0); L lower n x n row-major, B [n, nrhs]
 *   row-major, in place (tf.matrix_triangular_solve).  The workspace holds the inverses of the 128 x 128
 *   diagonal blocks of L, after which the substitution is a chain of DMMA products. */
size_t gabo_trsm_workspace_bytes(int64_t n, int64_t nrhs);
int gabo_trsm_f64(int trans, int64_t n, int64_t nrhs, const double* L, int64_t ldl,
                  double* B, int64_t ldb, void* workspace, size_t workspace_bytes, gabo_stream_t stream);
/* gabo_trsm_winv_f64: the same solve with the inverses of L's 128 x 128 diagonal blocks supplied
 *   (Winv_all [ceil(n/128)][128][128], e.g. from gabo_potrf_diag_f64): no inversion kernel.  `scratch`: 256 B of device
 *   memory (gabo_trsm_winv_workspace_bytes) for the progress counter of the one-right-hand-side kernel; may be NULL, in
 *   which case nrhs == 1 runs as the launch chain.
 * With nrhs == 1 both entry points run the whole substitution as ONE persistent kernel: each 128-row block of x is owned by a
 * resident CTA that streams its block row (or column) of L and consumes the earlier blocks as their owners publish them. */
size_t gabo_trsm_winv_workspace_bytes(void);
int gabo_trsm_winv_f64(int trans, int64_t n, int64_t nrhs, const double* L, int64_t ldl, const double* Winv_all,
                       double* B, int64_t ldb, void* scratch, size_t scratch_bytes, gabo_stream_t stream);
/* out[0] = sum_i log L_ii  (half the log-determinant of K + sigma^2 I). */
int gabo_logdet_f64(int64_t n, const double* L, int64_t ldl, double* out, gabo_stream_t stream);
/* out[0] = -0.5*n*p*log(2 pi) - p*sum_i log L_ii - 0.5*||alpha||_F^2, alpha = L^-1 (y - m) already solved,
 *   [n,p] row-major (GPflow 0.5 densities.multivariate_normal summed over the p output columns). */
int gabo_gp_loglik_f64(int64_t n, int64_t p, const double* L, int64_t ldl,
                       const double* alpha, int64_t ldalpha, double* out, gabo_stream_t stream);
/* Rank-1 growth of a cached factor by one observation (what the BO loop does every iteration instead of rebuilding the GPflow
 *   model: examples/bo_sphere/benchmark_examples/bo_sphere_ackley_manifold.py:137-160).  With Ky' = [[Ky, k], [k^T, kss]]:
 *   l = L^-1 k is given ([n] with stride ldl; gabo_trsm_f64 with one right-hand side), kss = k(x*,x*) + sigma^2.
 *   d = sqrt(kss - l.l);  Lrow[0..n) = l, Lrow[n] = d (the new row of the factor);  vnew[c] = (ynew[c] - mean_const - l.V[:,c]) / d
 *   (the new entry of V = L^-1 (y - m), [p]).  *info = 1 and nothing written if kss - l.l <= 0 (extended matrix not PD). */
int gabo_gp_append_row_f64(int64_t n, int64_t p, const double* l, int64_t ldl, double kss, const double* V, int64_t ldv,
                           const double* ynew, double mean_const, double* Lrow, double* vnew, int* info, gabo_stream_t stream);
/* Posterior at M test points from A = L^-1 K(X,X*) [n,M] and V = L^-1 (y-m) [n,p]:
 *   mean[M,p] = A^T V + mean_const ;  var[M] = kdiag[M] - colsum(A*A)   (GPR.build_predict, full_cov False). */
/* gabo_gp_grad_beta_f64: derivative of the log marginal likelihood w.r.t. the kernel's beta (the
 *   reference gets it from TensorFlow autodiff through GPflow's GPR.build_likelihood).  Every kernel of the path is
 *   K = variance * exp(beta * g), hence dK/dbeta = K log(K/variance) / beta and
 *   out[0] = 1/(2 beta) sum_ij (alpha alpha^T - p Kinv)_ij K_ij log(K_ij/variance), read from the LOWER triangles of K
 *   (noise-free Gram) and negKinv = -(K + sigma^2 I)^-1; alpha [n, p] = (K + sigma^2 I)^-1 (y - m).  Deterministic. */
size_t gabo_gp_grad_beta_workspace_bytes(void);
int gabo_gp_grad_beta_f64(int64_t n, int64_t p, const double* K, int64_t ldk, const double* negKinv, int64_t ldi,
                          const double* alpha, int64_t lda, double variance, double beta, double* out,
                          void* workspace, size_t workspace_bytes, gabo_stream_t stream);
int gabo_gp_predict_f64(int64_t n, int64_t M, int64_t p, const double* A, int64_t lda,
                        const double* V, int64_t ldv, const double* kdiag, double mean_const,
                        double* mean_out, double* var_out, gabo_stream_t stream);
/* C <- C - A B^T on the lower triangle only when `lower` != 0 (SYRK-shaped when A == B), else full:
 *   C [m,n], A [m,k], B [n,k] row-major.  The DMMA trailing-update kernel of the Cholesky, exported
 *   because the multi-GPU block-cyclic factorisation and predict_f_full_cov (cov = K** - A^T A) reuse it. */
int gabo_gemm_nt_sub_f64(int64_t m, int64_t n, int64_t k, const double* A, int64_t lda,
                         const double* B, int64_t ldb, double* C, int64_t ldc, int lower,
                         gabo_stream_t stream);

/* gabo_sphere_posterior_grad_f64: Euclidean gradients of the posterior mean / variance and of Expected Improvement with
 *   respect to each of M candidate points on the sphere (what GPflowOpt's evaluate_with_gradients hands to
 *   BO_utils/manifold_optimization.py:260-271).  Kx [n, M] = K(X, X*) (noise-free cross-Gram), alpha [n] = Ky^-1 (y - m),
 *   Bm [n, M] = Ky^-1 Kx, mean / var [M] the posterior at the candidates.  gmean, gvar, dei: [M, dim]; dei may be NULL
 *   (then mean / var / fmin are ignored).  dim <= 64.  kind: GABO_SPHERE_GAUSSIAN or GABO_SPHERE_LAPLACE. */
int gabo_sphere_posterior_grad_f64(int kind, int64_t n, int64_t M, int dim, const double* X, int64_t ldx,
                                   const double* Kx, int64_t ldk, const double* alpha, const double* Bm, int64_t ldb,
                                   double beta, double variance, const double* mean, const double* var, double fmin,
                                   double* gmean, double* gvar, double* dei, gabo_stream_t stream);
/* gabo_spd_ai_posterior_grad_f64: the same for candidate SPD matrices under the affine-invariant kernels
 *   (examples/bo_spd/benchmark_examples/bo_spd_ackley_manifold.py:159-187 through manifold_optimization.py:260-271).  For
 *   M_i = W1[i] X W1[i]^T = V Lambda V^T the Euclidean gradient of d_AI(A_i, X)^2 with respect to the symmetric matrix X is
 *   2 sum_k (log lambda_k / lambda_k) u_k u_k^T, u_k = W1[i]^T v_k (= 2 X^-1 Log(X A_i^-1)); Mandel coordinates scale the
 *   off-diagonal slots by sqrt(2).  W1 [n, d*d] inverse Cholesky factors of the training matrices (gabo_spd_prep_f64), S2
 *   [M, d*d] the candidates, alpha [n] = Ky^-1 (y - m), Bm [n, M] = Ky^-1 K(X, X*), mean / var [M] the posterior at the
 *   candidates.  gmean, gvar, dei: [M, d(d+1)/2]; dei may be NULL (then mean / var / fmin are ignored).  2 <= d <= 6. */
int gabo_spd_ai_posterior_grad_f64(int kind, int d, int64_t n, int64_t M, const double* W1, const double* S2,
                                   const double* alpha, const double* Bm, int64_t ldb, double beta, double variance,
                                   const double* mean, const double* var, double fmin, double* gmean, double* gvar,
                                   double* dei, gabo_stream_t stream);
/* gabo_sphere_acq_fused_f64: the acquisition evaluation of the reference's optimisers in ONE launch --
 *   MultistartedOptimizer scores 100-1000 anchors per BO iteration (manifold_optimization.py:402-431) and the manifold CG asks
 *   for (-EI, -dEI/dx) one candidate at a time (:250-271).  Per candidate: k = K(X, x*) -> a = L^-1 k -> mean = a.V + mean_const,
 *   var = variance - a.a (GPR.build_predict) -> EI (GPflowOpt) -> optionally the Euclidean gradients of mean / var / EI.
 *   Linv [n, n] = L^-1 (lower triangular, row-major; e.g. gabo_trsm_f64 on the identity, once per factorisation),
 *   V [n] = L^-1 (y - m), alpha [n] = Ky^-1 (y - m) (read only with gradients).  mean / var / ei: [M]; gmean / gvar / dei:
 *   [M, dim], all NULL for values only (dei alone may be NULL).  n <= 9000, dim <= 64, sphere kinds. */
int gabo_sphere_acq_fused_f64(int kind, int64_t n, int64_t M, int dim, const double* X, int64_t ldx, const double* Xc,
                              int64_t ldc, const double* Linv, int64_t ldl, const double* V, const double* alpha,
                              double beta, double variance, double mean_const, double fmin, double* mean, double* var,
                              double* ei, double* gmean, double* gvar, double* dei, gabo_stream_t stream);
/* Expected Improvement over M candidates from the posterior mean / variance (GPflowOpt ExpectedImprovement
 * [third party]; consumed by the reference at BO_utils/manifold_optimization.py:250-271,427):
 *   out = (fmin - mean) Phi(z) + var N(fmin; mean, sqrt(var)), var floored at 1e-6, z = (fmin - mean)/sqrt(var). */
int gabo_expected_improvement_f64(int64_t M, const double* mean, const double* var, double fmin, double* out,
                                  gabo_stream_t stream);

/* ---- FP64 sphere Gram with integer-tensor-core inner products (csrc/points_gram_i8.cu) --------------------------------------
 * Same contract and arguments as gabo_points_gram_f64 for kind = GABO_SPHERE_GAUSSIAN / GABO_SPHERE_LAPLACE and dim <= 64
 * (sphere_utils_tf.py:11-60 + kernels_sphere_tf.py:59-88, 179-204), FP64 in and out.  The inner products x.y are formed on the
 * 5th-generation tensor cores in 8-bit integer arithmetic (tcgen05.mma kind::i8, S32 accumulators in TMEM) from an exact
 * 8 x 7-bit slicing of 56-bit fixed-point coordinates; 36 of the 64 slice products are kept: |error of x.y| < 6e-15 (max |x|)^2
 * worst case at dim = 52, i.e. < ~1e-14 beta relative on a Gaussian Gram entry (north_star tolerance: 1e-12).  The FP64 pipe,
 * which DMMA shares with DFMA on B200, is left to the clamp / acos / exp epilogue.  Coordinates must be finite.
 * cyc_world > 0: block-cyclic row sharding for one NVSwitch node -- this rank computes the FULL rows of its 512-row panels
 * (panel P on rank P % cyc_world, local slot P / cyc_world) into its local [local_rows, ldk] buffer K; uplo must be GABO_FULL,
 * n1 a multiple of 512, row0 / row1 are ignored.  No data leaves the GPU.
 * workspace: gabo_points_gram_i8_workspace_bytes(n1, n2) bytes, 128-byte aligned. */
size_t gabo_points_gram_i8_workspace_bytes(int64_t n1, int64_t n2);
int gabo_points_gram_i8_f64(int kind, const double* X, int64_t n1, int64_t ldx, const double* X2, int64_t n2, int64_t ldx2,
                            int dim, double beta, double variance, double* K, int64_t ldk, int64_t row0, int64_t row1, int uplo,
                            int cyc_world, int cyc_rank, void* workspace, size_t workspace_bytes, gabo_stream_t stream);

/* ---- small-N batched GP evaluation (csrc/gp_small.cu) -----------------------------------------------------------------
 * The reference's small configurations (examples/kernels/sphere/sphere_kernels.py:125-139, examples/bo_sphere/benchmark_examples/
 * bo_sphere_ackley_manifold.py:137-160, examples/bo_spd/benchmark_examples/bo_spd_ackley_manifold.py:159-187) evaluate GPflow 0.5's
 * GPR.build_likelihood [third party] and its gradient hundreds of times per BO iteration at N = 5..100+.  This entry point
 * evaluates B hyper-parameter sets in ONE launch (one CTA per set, Gram -> Cholesky -> solve -> log-density -> gradients fused
 * in shared memory):   Ky_b = v_b exp(r_b G) + s2_b I,   G [n, ldg] = log(K_beta0 / variance0) (gabo_log_gram_f64, G_ii = 0),
 * theta [B, 4] = {r = beta/beta0, v, s2, constant mean c};  Y [n, ldy], p columns.
 *   out [B, 5] = { log marginal likelihood, d/dr, d/dv, d/ds2, d/dc } (the derivatives are NaN unless want_grad);
 *   info [B]   = 0, or k if the k-th pivot of Ky_b is not positive (out[b] is then NaN);
 *   Lout (may be NULL): lower Cholesky factor of set b at Lout + b*l_stride, row stride ldl;  Vout (may be NULL): [B, n, p] L^-1 (Y - c).
 * 1 <= n <= GABO_GP_SMALL_MAX_N, 1 <= p <= GABO_GP_SMALL_MAX_P (and (n + p)(n | 1) + (p + 2) n + 32 doubles <= 227 KB). */
int gabo_gp_small_batch_f64(int64_t n, int64_t p, const double* G, int64_t ldg, const double* Y, int64_t ldy, int64_t B,
                            const double* theta, int want_grad, double* out, int32_t* info, double* Lout, int64_t ldl,
                            int64_t l_stride, double* Vout, gabo_stream_t stream);

/* ---- SPD exponential / logarithm maps (affine-invariant metric), batched ------------------------------------------
 * Replace expmap_tf / logmap_tf (SPD_utils_tf.py:142-167, 170-195) and the NumPy / Mandel twins (SPD_utils.py:104-177):
 *   expmap: out[i] = S_i expm(S_i^-1 U[i]) = L f(W U W^T) L^T with f = exp;   logmap: out[i] = S_i logm(S_i^-1 X[i]), f = log
 *   (S_i = L L^T, W = L^-1; symmetric eigenproblem in real FP64 -- the reference goes through complex64).
 * mandel == 0: every matrix is [d*d] row-major (ld >= d*d); mandel != 0: Mandel vectors of length d(d+1)/2 (SPD_utils_tf.py:11-92).
 * base_stride == 0: one base point for all n inputs (the reference's signature); otherwise base + i*base_stride.
 * info[i] (may be NULL): 0, k in 1..d if the k-th Cholesky pivot of the base is not positive, d+1+k if (logmap) eigenvalue k of
 * W X W^T is not positive; the outputs of such points are NaN.  2 <= d <= 8. */
int gabo_spd_expmap_f64(int d, const double* U, int64_t ldu, const double* base, int64_t base_stride, int64_t n, int mandel,
                        double* out, int64_t ldo, int32_t* info, gabo_stream_t stream);
int gabo_spd_logmap_f64(int d, const double* X, int64_t ldx, const double* base, int64_t base_stride, int64_t n, int mandel,
                        double* out, int64_t ldo, int32_t* info, gabo_stream_t stream);
/* symmetric_matrix_to_vector_tf (SPD_utils_tf.py:68-92; NumPy twin SPD_utils.py:57-76): [n, d*d] matrices -> Mandel vectors
 * [n, d(d+1)/2] (diagonal first, then super-diagonals in order, off-diagonal slots times sqrt(2)); the upper triangle is read.
 * (The inverse, vector_to_symmetric_matrix_tf, is gabo_spd_prep_f64's out_S.) */
int gabo_sym_to_mandel_f64(int d, const double* S, int64_t lds, int64_t n, double* out, int64_t ldo, gabo_stream_t stream);

/* ---- beta sweeps: extreme eigenvalues of K_beta for a grid of beta (csrc/eig_sweep.cu) ----------------------------------
 * Device-side form of examples/kernels/sphere/sphere_gaussian_kernel_parameters.py:83-104 and
 * examples/kernels/spd/spd_gaussian_kernel_parameters.py:108-117 (Gram + numpy.linalg.eig minimum per beta and trial).
 * Every kernel of the path is variance*exp(-beta*g), so K_beta = variance*exp(ratio*G) with G = log(K_beta0/variance) and
 * ratio = beta/beta0: ONE Gram per point set serves the whole beta grid.
 * gabo_log_gram_f64: G = scale * min(0, log(K/variance)) elementwise (in place allowed): scale = 1 gives the operand of the
 *   sweep; scale = -1 (Laplace kinds, beta = 1) or -1 (Gaussian kinds: squared distance) recovers the distance matrix of
 *   sphere_distance_tf (sphere_utils_tf.py:11-60) / affine_invariant_distance_tf (SPD_utils_tf.py:95-139) from a Gram.
 * gabo_lanczos_sweep_f64: for each of n_mat log-Grams (G + t*g_stride, [n, ldg]) and each of n_ratios ratios (device array),
 *   one CTA runs Lanczos with full reorthogonalisation on the operator v -> K v (entries re-formed on the fly, no second
 *   n x n matrix), at most m_max steps (clamped to n; m_max == n gives dense-eigensolver accuracy, O(eps*||K||)), then takes
 *   the extreme eigenvalues of the tridiagonal matrix by Sturm-count multisection.
 *   out [n_mat*n_ratios, 4] (matrix-major): lambda_min estimate (an upper bound when the run stopped early), lambda_max
 *   estimate, residual estimate of the minimal Ritz pair (0 if exact), steps taken.  tol > 0 enables an early exit (checked
 *   every 32 steps) once that residual <= tol*||T||.  alpha_out / beta_out: optional [n_mat*n_ratios, m_max] tridiagonal
 *   coefficients (both or neither).  n <= 16384, m_max <= 2048.  workspace: gabo_lanczos_sweep_workspace_bytes. */
int gabo_log_gram_f64(const double* K, int64_t ldk, double* G, int64_t ldg, int64_t n1, int64_t n2, double variance,
                      double scale, gabo_stream_t stream);
size_t gabo_lanczos_sweep_workspace_bytes(int64_t n, int n_mat, int n_ratios, int m_max);
int gabo_lanczos_sweep_f64(const double* G, int64_t ldg, int64_t g_stride, int64_t n, int n_mat, const double* ratios,
                           int n_ratios, double variance, int m_max, double tol, unsigned long long seed, double* out,
                           double* alpha_out, double* beta_out, void* workspace, size_t workspace_bytes, gabo_stream_t stream);

/* Diagnostic: element-wise evaluation of the lean device math that the Gram epilogues inline (csrc/gabo_math.cuh), through
 * the same batched code paths: which = 0 acos(clamp(x,-1,1)) (sphere_utils_tf.py:59-60), 1 scale*exp(x) for x <= 0
 * (kernels_sphere_tf.py:88), 2 log(x) for positive normal x (SPD_utils_tf.py:136).  For accuracy tests only. */
int gabo_math_probe_f64(int which, const double* x, int64_t n, double scale, double* out, gabo_stream_t stream);

/* Diagnostic: in-situ timing of the dominant kernel.  gabo_profile_trailing_updates(1) makes gabo_potrf_f64 bracket each big
 * trailing-update launch (gemm_kernel, k = 512) with CUDA events on its stream; gabo_profile_trailing_read sums them
 * (launch count, device ms, algorithmic flops, and the longest launch); (0) switches it off and drops the records. */
int gabo_profile_trailing_updates(int enable);
int gabo_profile_trailing_read(int64_t* launches, double* total_ms, double* total_flops, double* max_ms, double* max_flops);

/* ---- Building blocks of the multi-GPU factorisation (1-D block-cyclic by row panels) --------
 * The host side (gaboflow_b200/dist.py) broadcasts the factored diagonal block and all-gathers the panel over NCCL;
 * these calls do the local work of each rank.
 * gabo_trsm_right_f64: A[m, nb] <- A * L^-T, L lower nb x nb (panel rows owned by this rank). */
/* Enables loads/stores from kernels on the current device to memory of `peer_device` (cudaDeviceEnablePeerAccess);
 * returns -1 if the two devices have no peer path. */
int gabo_enable_peer_access(int peer_device);
/* Peer-shared buffers for gabo_points_gram_cyclic_f64: cudaMalloc'd on the current device (gabo_shared_alloc/free),
 * exported as a 64-byte CUDA IPC handle (gabo_ipc_export) and opened by the other ranks of the node on THEIR device
 * with lazy peer access (gabo_ipc_open / gabo_ipc_close), so that kernels there may store into it over NVLink. */
int gabo_shared_alloc(size_t bytes, void** out);
int gabo_shared_free(void* ptr);
int gabo_ipc_export(void* ptr, unsigned char* handle64);
int gabo_ipc_open(const unsigned char* handle64, void** out);
int gabo_ipc_close(void* ptr);
size_t gabo_trsm_right_workspace_bytes(int64_t m, int64_t nb);
/* gabo_points_gram_cyclic_f64: symmetric Gram K(X,X) distributed by 512-row panels, 1-D block-cyclic (panel P on rank
 *   P % world, local slot P / world), with the symmetry exploited ACROSS GPUs: this rank computes only the lower tiles
 *   of its own row panels and stores every off-diagonal tile's mirror image directly into the HBM of the rank owning
 *   that row panel (peer-mapped pointers over NVLink/NVSwitch; the transfer overlaps the tile math, no collective).
 *   K_bases_host: HOST array of `world` device pointers, entry r = base of rank r's [local_rows, ldk] buffer as mapped
 *   in this process (entry `rank` is the local buffer; the others come from cudaIpcOpenMemHandle).  n must be a multiple
 *   of 512, dim <= 52.  Callers synchronise the stream and barrier across ranks before reading their rows. */
int gabo_points_gram_cyclic_f64(int kind, const double* X, int64_t n, int64_t ldx, int dim, double beta, double variance,
                                double* const* K_bases_host, int64_t ldk, int world, int rank,
                                void* workspace, size_t workspace_bytes, gabo_stream_t stream);
/* gabo_points_gram_cyclic_hybrid_f64: as above, but recompute_frac256 / 256 of the panel pairs owned by DIFFERENT ranks
 *   (selected by a hash of the pair) are RECOMPUTED by the owner of the row panel instead of being mirrored to it over
 *   NVLink (0: mirror all).  Producing an entry costs ~3 ps of FP64 pipe, moving it ~10 ps of NVLink: recomputing part of the
 *   upper triangle balances the two.  Workspace: gabo_points_gram_cyclic_workspace_bytes(n, dim). */
size_t gabo_points_gram_cyclic_workspace_bytes(int64_t n, int dim);
int gabo_points_gram_cyclic_hybrid_f64(int kind, const double* X, int64_t n, int64_t ldx, int dim, double beta,
                                       double variance, double* const* K_bases_host, int64_t ldk, int world, int rank,
                                       int recompute_frac256, void* workspace, size_t workspace_bytes, gabo_stream_t stream);
/* gabo_points_gram_cyclic_staged_f64: the same distributed symmetric Gram, but no SM ever stores to a peer: mirror images
 *   bound for other ranks are written (transposed) into a LOCAL staging buffer, laid out so that everything a chunk of this
 *   rank's row tiles owes one peer is a single dense 2-D block; the kernel counts finished tiles per chunk and raises a flag
 *   when a chunk is complete, and copy-engine transfers (cudaMemcpy2DAsync on per-peer side streams, each released by
 *   cuStreamWaitValue32 on that flag) move the blocks over NVLink while the kernel computes on.  Chunks: one per local panel,
 *   the last `split` panels one per 128-row tile.  recompute_frac256 / 256 of each (row panel, destination rank) block list --
 *   its LOWEST slots, so that the mirrored remainder stays one contiguous row range -- is recomputed by the destination rank
 *   instead (the recomputed tiles run last and cover the tail of the copies); the next direct_frac256 / 256 of the slots are
 *   stored into the peer's HBM by the kernel itself, as in the hybrid function, so that the two transports -- SM stores and copy
 *   engines, each of which saturates well below NVLink's rate in an 8-GPU all-to-all -- work side by side.  On return the caller's stream waits for this
 *   rank's outgoing copies; completion across ranks is the caller's flag handshake, as for the functions above.
 *   stage: gabo_points_gram_cyclic_stage_bytes(n, world, rank, recompute_frac256, direct_frac256, split) bytes of local device memory;
 *   progress: >= 2 * (local panels + 3 * split) 32-bit words of local device memory, owned by this call sequence;
 *   epoch: strictly increasing over successive calls on the same `progress` (> 0).  Bit-identical to the other two. */
size_t gabo_points_gram_cyclic_stage_bytes(int64_t n, int world, int rank, int recompute_frac256, int direct_frac256, int split);
int gabo_points_gram_cyclic_staged_f64(int kind, const double* X, int64_t n, int64_t ldx, int dim, double beta, double variance,
                                       double* const* K_bases_host, int64_t ldk, int world, int rank, int recompute_frac256,
                                       int direct_frac256, int split, double* stage, size_t stage_bytes, uint32_t* progress,
                                       int64_t progress_words, uint32_t epoch, void* workspace, size_t workspace_bytes,
                                       gabo_stream_t stream);
int gabo_trsm_right_f64(int64_t m, int64_t nb, const double* L, int64_t ldl, double* A, int64_t lda,
                        void* workspace, size_t workspace_bytes, gabo_stream_t stream);
/* gabo_potrf_diag_f64: factor one diagonal block (nb <= 512) in place AND return the inverses of its 128 x 128 diagonal
 *   blocks, Winv_all [ceil(nb/128)][128][128]; the owner of a panel broadcasts factor + inverses, so that
 * gabo_trsm_right_winv_f64 (same as gabo_trsm_right_f64 with the inverses supplied) has no inversion on its path. */
int gabo_potrf_diag_f64(int64_t nb, double* A, int64_t lda, double sigma2, int32_t* info, double* Winv_all,
                        void* workspace, size_t workspace_bytes, gabo_stream_t stream);
int gabo_trsm_right_winv_f64(int64_t m, int64_t nb, const double* L, int64_t ldl, const double* Winv_all,
                             double* A, int64_t lda, void* workspace, size_t workspace_bytes, gabo_stream_t stream);
/* gabo_trsm_right_scatter_f64: gabo_trsm_right_winv_f64 that ALSO delivers the solved block column to every rank: each
 *   solved 128-column block is stored, in global row order, into the n_dst (<= GABO_MAX_PEERS) block-column buffers
 *   P_dst_host[d] [rows below the panel][ldp] -- this rank's own and the peers' (CUDA IPC mappings; NVLink stores issued
 *   as the blocks are produced, no pack / all-gather / reorder).  Local row i lands on row
 *   p_row_off + (i / cyc_nb) * cyc_world * cyc_nb + i % cyc_nb.  ldp even, buffers 16-byte aligned. */
int gabo_trsm_right_scatter_f64(int64_t m, int64_t nb, const double* L, int64_t ldl, const double* Winv_all,
                                double* A, int64_t lda, double* const* P_dst_host, int n_dst, int64_t ldp,
                                int64_t p_row_off, int cyc_nb, int cyc_world, void* workspace, size_t workspace_bytes,
                                gabo_stream_t stream);
/* Stream-ordered flags between the ranks of a node.  gabo_flag_signal: after everything earlier in `stream` is visible
 *   system-wide, store `value` into the n 32-bit flags flags_host[i] (HOST array of device pointers, typically one flag in
 *   each peer's memory).  gabo_flag_wait_geq: `stream` stalls (cuStreamWaitValue32, no SM occupied) until *flag >= value;
 *   `flag` must be in this device's memory.  gabo_copy_async: cudaMemcpyAsync between local / IPC-mapped buffers. */
int gabo_flag_signal(uint32_t* const* flags_host, int n, uint32_t value, gabo_stream_t stream);
int gabo_flag_wait_geq(const uint32_t* flag, uint32_t value, gabo_stream_t stream);
/* gabo_push_signal_f64: payload + flag in one launch -- copies `count` doubles from src into each of the n buffers
 *   dst_host[i] (peers' memory) and then, after a system-scope fence, stores `value` into the n flags flags_host[i].  Used by
 *   the distributed substitution to publish each solved 512-block of x. */
int gabo_push_signal_f64(const double* src, int64_t count, double* const* dst_host, uint32_t* const* flags_host, int n,
                         uint32_t value, gabo_stream_t stream);
int gabo_copy_async(void* dst, const void* src, size_t bytes, gabo_stream_t stream);
/* gabo_potrf_cyclic_p2p_f64: the distributed right-looking Cholesky of the leading n x n block (lower triangle stored
 *   block-cyclically: global panel p of nb rows on rank p % world, local slot p / world; A_loc [local_rows, >= n]) as ONE host
 *   call per rank: peer-memory panel exchange, one panel of look-ahead, no collective.  Main stream: wait for the rows of block
 *   column kp from every rank -> U1(kp) (next block column) -> U2(kp) (rest).  Side stream (released by U1): the owner factors
 *   diagonal block kp + 1 (gabo_potrf_diag_f64), pushes factor + block inverses to the peers with copy engines and raises their
 *   flag; every rank solves its rows of block column kp + 1 (gabo_trsm_right_scatter_f64: the result lands in the block-column
 *   buffer of EVERY rank by NVLink stores) and raises the peers' flags.
 *   P_host / bc_host: HOST arrays of world * 3 device pointers (buffer b of rank r at [3 r + b]): block-column buffers [n][nb]
 *   and diagonal-block buffers (factor [nb][nb] + inverses [nb/128][128][128]); flag_host [world]: base of every rank's 32-word
 *   flag area ([0] diagonal block arrived, [1 + r] rows of rank r arrived); values base + kp + 1.  info_dev [ceil(n/nb)]: LAPACK
 *   info per diagonal block (owner's entries).  winv_store: NULL or [local slots][nb/128][128][128].  ws_diag
 *   (gabo_potrf_workspace_bytes(nb)) and ws_trsm (gabo_trsm_right_workspace_bytes(local_rows, nb)) are used on the side
 *   stream.  ev_u1 / ev_side: events from gabo_event_create.  Same arithmetic, per entry and in order, as the single-GPU
 *   factorisation of each rank's rows. */
int gabo_event_create(void** ev);
int gabo_event_destroy(void* ev);
int gabo_potrf_cyclic_p2p_f64(int64_t n, int nb, int world, int rank, double* A_loc, int64_t lda, double sigma2,
                              double* const* P_host, double* const* bc_host, uint32_t* const* flag_host, uint32_t base,
                              int32_t* info_dev, double* winv_store, void* ws_diag, size_t ws_diag_bytes, void* ws_trsm,
                              size_t ws_trsm_bytes, gabo_stream_t main_stream, gabo_stream_t side_stream, void* ev_u1, void* ev_side);
/* gabo_potrf_cyclic_chain_f64: the same factorisation (same arithmetic per entry, same buffers and flags) sequenced along its
 *   critical chain  D(k) -> push to the owner of k+1 -> that owner solves its own nb rows of block column k -> updates its
 *   diagonal block -> D(k+1):  the next owner does exactly that on the high-priority stream (and receives the diagonal block and
 *   its flag before the other ranks); the rest of every rank's row solve runs on a second internal high-priority stream, the
 *   update of the next block column is released by ONE nb x nb block (flag word [9]) instead of every rank's rows, and only the
 *   big trailing update waits for all rows.  The owners may run up to `world` panels ahead of a rank that owns none of them, so
 *   P_host / bc_host hold nbuf >= world + 1 (>= 3) rotating buffers per rank (buffer b of rank r at [nbuf r + b]).
 *   ws_diag: max(gabo_potrf_workspace_bytes(nb), gabo_trsm_right_workspace_bytes(nb, nb)) bytes. */
int gabo_potrf_cyclic_chain_f64(int64_t n, int nb, int world, int rank, double* A_loc, int64_t lda, double sigma2, int nbuf,
                                double* const* P_host, double* const* bc_host, uint32_t* const* flag_host, uint32_t base,
                                int32_t* info_dev, double* winv_store, void* ws_diag, size_t ws_diag_bytes, void* ws_trsm,
                                size_t ws_trsm_bytes, gabo_stream_t main_stream, gabo_stream_t side_stream);
/* gabo_solve_cyclic_p2p_f64: the forward substitution L x = b of the block-cyclically distributed factor (panel p of nb rows on
 *   rank p % world, local slot p / world; A_loc [local_rows, >= n], B_loc [local_rows, nrhs] contiguous, in place) as ONE host
 *   call per rank, no collective: at step kp the owner solves its block (inverses of the 128-blocks of the diagonal block in
 *   winv_host[kp], device pointers in a HOST array of ceil(n/nb) entries, only the owner's are read) and publishes it into slot
 *   kp of every peer's x area (xarea_host[r], HOST array of `world` device pointers; nb * nrhs doubles per slot) together with
 *   the flag value base + kp + 1 (flag_host[r], one 32-bit word in each rank's memory); the peers' streams wait for the flag.
 *   The owner of panel kp + 1 updates those rows first, solves and publishes, then updates the rest of its rows.
 *   scratch: >= 256 B of device memory.  Successive solves must be separated by a rank-wide synchronisation and use
 *   increasing `base` values. */
int gabo_solve_cyclic_p2p_f64(int64_t n, int nb, int world, int rank, const double* A_loc, int64_t lda, double* B_loc,
                              int64_t ldb, int64_t nrhs, const double* const* winv_host, double* const* xarea_host,
                              uint32_t* const* flag_host, uint32_t base, void* scratch, size_t scratch_bytes,
                              gabo_stream_t stream);
/* gabo_syrk_cyclic_f64: C[m,n] -= A[m,k] B[n,k]^T restricted to GLOBAL col <= GLOBAL row, where local row r of C is
 *   global row row_base + ((r / cyc_nb) * cyc_world + cyc_rank) * cyc_nb + r % cyc_nb (cyc_nb == 0: row_base + r) and
 *   local column c is global column col_base + c. */
int gabo_syrk_cyclic_f64(int64_t m, int64_t n, int64_t k, const double* A, int64_t lda, const double* B, int64_t ldb,
                         double* C, int64_t ldc, int64_t row_base, int64_t col_base, int cyc_nb, int cyc_world,
                         int cyc_rank, gabo_stream_t stream);
/* gabo_solve_update_f64: Y[m,nrhs] -= Lp[m,k] X[k,nrhs] (substitution update of the rows a rank owns). */
int gabo_solve_update_f64(int64_t m, int64_t k, int64_t nrhs, const double* Lp, int64_t ldl, const double* X, int64_t ldx,
                          double* Y, int64_t ldy, gabo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GABO_B200_H */
