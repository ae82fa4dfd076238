// SPD-manifold kernels (sm_100a): per-point preparation and the pair Gram for the affine-invariant and Stein kernels.
//
// Replaces  vector_to_symmetric_matrix_tf (reference SPD_utils_tf.py:11-65),
//           affine_invariant_distance_tf (SPD_utils_tf.py:95-139: per pair matrix_inverse + sqrtm + 2 matmul + svd),
//           SpdAffineInvariant{Gaussian,Laplace}Kernel.K (kernels_spd_tf.py:214-248, 368-400),
//           SpdSteinGaussianKernel.K / Kdiag (kernels_spd_tf.py:70-129), and the per-pair logm of
//           SpdLogEuclideanGaussianKernel.K (kernels_spd_tf.py:634) which becomes a per-point Mandel(logm).
//
// Data layout in HBM: S, W as [n, d*d] row-major doubles (W lower triangular, zeros above), logdet [n],
// logm as Mandel [n, m].  Pair kernels: a CTA owns a 32-row x 128-column block of K; thread = one column j with
// S_j held in registers, looping over the 32 rows whose operand (W_i or S_i) sits in shared memory and is read as a
// warp broadcast.  Affine-invariant kinds (spd_ai_gram_kernel): per pair the d x d congruence W_i S_j W_i^T and its
// Householder tridiagonalisation in FP64, eigenvalue STARTS by an implicit-shift QL in FP32 run for TWO rows at once in
// packed FFMA2 registers, two FP64 Newton steps per eigenvalue on the characteristic polynomial (validated; an all-FP64
// QL / cyclic Jacobi slow path otherwise), lean log / exp.  Stein (spd_gram_kernel): LDL^T of S_i + S_j in registers.
// Row stores are 256 B contiguous per warp; the mirrored block goes through a padded shared-memory transpose.
#include <cstdlib>
#include "spd_small.cuh"
#include "gabo_math.cuh"

namespace gabo {
namespace {

constexpr int TI = 32;    // rows per CTA
constexpr int TJ = 128;   // columns per CTA == threads per CTA

// ---------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void spd_prep_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, double* __restrict__ outS,
                                double* __restrict__ outW, double* __restrict__ outLogdet,
                                double* __restrict__ outLogm, int32_t* __restrict__ info) {
    constexpr int M = D * (D + 1) / 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double inv_sqrt2 = 0.70710678118654752440;
    const double sqrt2 = 1.41421356237309504880;
    double S[D][D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = r; c < D; ++c) {
            double v = X[i * ldx + mandel_slot<D>(r, c)];
            if (c != r) v *= inv_sqrt2;   // SPD_utils_tf.py:59: off-diagonal slots are stored times sqrt(2)
            S[r][c] = v;
            S[c][r] = v;
        }
    if (outS) {
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
            for (int c = 0; c < D; ++c) outS[i * D * D + r * D + c] = S[r][c];
    }
    double L[D][D];
    const int bad = chol_lower<D>(S, L);
    info[i] = bad;
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    if (outW) {
        double W[D][D];
        tri_inverse_lower<D>(L, W);
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
            for (int c = 0; c < D; ++c) outW[i * D * D + r * D + c] = bad ? nanv : (c <= r ? W[r][c] : 0.0);
    }
    if (outLogdet) {
        double prod = 1.0;
#pragma unroll
        for (int k = 0; k < D; ++k) prod *= L[k][k];
        double ld = 0.0;
        if (prod > 1e-140 && prod < 1e140) {
            ld = 2.0 * log(prod);
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) ld += 2.0 * log(L[k][k]);
        }
        outLogdet[i] = bad ? nanv : ld;
    }
    if (outLogm) {
        double V[D][D];
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
            for (int c = 0; c < D; ++c) V[r][c] = (r == c) ? 1.0 : 0.0;
        jacobi_eig<D, true>(S, V);
        double lw[D];
#pragma unroll
        for (int k = 0; k < D; ++k) lw[k] = log(S[k][k]);   // NaN for a non-positive eigenvalue, as tf.log would give
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
            for (int c = r; c < D; ++c) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) v = fma(V[r][k] * lw[k], V[c][k], v);
                if (c != r) v *= sqrt2;   // symmetric_matrix_to_vector_tf, SPD_utils_tf.py:87-89
                outLogm[i * M + mandel_slot<D>(r, c)] = v;
            }
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct SpdGramParams {
    const double* R;        // row operand [n1, d*d]: W (AI kinds) or S (Stein)
    const double* Rld;      // logdet of row points (Stein) or nullptr
    const double* Cm;       // column operand S2 [n2, d*d]
    const double* Cld;      // logdet of column points (Stein) or nullptr
    double* K;
    int64_t ldk, n1, n2, row0, row1;
    int symmetric, uplo;
    double beta, variance;
};

// M = W S W^T for one pair: W lower (packed triangle wrow: entry (a,k), k <= a, at a(a+1)/2 + k), S symmetric (upper storage).
template <int D>
__device__ __forceinline__ void ai_congruence(const double* __restrict__ wrow, const double (&B)[D][D], double (&Mx)[D][D]) {
#pragma unroll
    for (int a = 0; a < D; ++a) {
        double T[D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k <= a; ++k) v = fma(wrow[a * (a + 1) / 2 + k], GABO_SYM(B, k, c), v);
            T[c] = v;
        }
#pragma unroll
        for (int b = 0; b <= a; ++b) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k <= b; ++k) v = fma(T[k], wrow[b * (b + 1) / 2 + k], v);
            Mx[b][a] = v;   // upper storage
        }
    }
}

// Out-of-line slow path of the affine-invariant pair (kept out of the hot loop: ~3 k instructions that would otherwise sit
// between the fast path's blocks in the instruction cache and inflate its register allocation).  Taken when the
// mixed-precision eigenvalues do not validate (near-multiple eigenvalues) or the spectrum is wide: all-FP64 QL on the
// tridiagonal, then cyclic Jacobi (relative accuracy) if lambda_min <= 0.02 lambda_max.  S2 travels by value (packed upper
// triangle) so that the caller's copy can stay in registers.  Returns sum_k log^2 lambda_k.
template <int D>
struct SymPacked { double v[D * (D + 1) / 2]; };

template <int D>
__device__ __noinline__ double ai_sumlog2_slow(const double* __restrict__ wrow, const SymPacked<D> Sp) {
    double B[D][D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = r; c < D; ++c) B[r][c] = Sp.v[r * D - r * (r - 1) / 2 + (c - r)];
    double Mx[D][D], ev[D];
    ai_congruence<D>(wrow, B, Mx);
    tridiag_ql_eigvals<D>(Mx, ev);
    double emin = ev[0], emax = ev[0];
#pragma unroll
    for (int k = 1; k < D; ++k) { emin = fmin(emin, ev[k]); emax = fmax(emax, ev[k]); }
    if (!(emin > 0.02 * emax)) {
        ai_congruence<D>(wrow, B, Mx);
        double dummy[1][1];
        jacobi_eig<D, false>(Mx, dummy);
#pragma unroll
        for (int k = 0; k < D; ++k) ev[k] = Mx[k][k];
    }
    double s2 = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const double lg = log(ev[k]);
        s2 = fma(lg, lg, s2);
    }
    return s2;
}

template <int D, int KIND>
__global__ void __launch_bounds__(TJ, (D <= 6) ? 4 : 2) spd_gram_kernel(const SpdGramParams p) {
    constexpr int DD = D * D;
    constexpr int MP = D * (D + 1) / 2;   // packed lower triangle: entry (a,k), k <= a, at a(a+1)/2 + k
    __shared__ double sR[TI][MP];
    __shared__ double sRld[TI];
    __shared__ double sOut[TI][TJ + 1];

    const int tid = threadIdx.x;
    const int64_t i0 = p.row0 + (int64_t)blockIdx.y * TI;
    const int64_t j0 = (int64_t)blockIdx.x * TJ;
    // symmetric modes: blocks strictly above the diagonal are not needed
    if (p.uplo != GABO_FULL && j0 > i0 + TI - 1) return;
    const bool mirror = (p.uplo == GABO_SYMM_MIRROR);

    for (int e = tid; e < TI * MP; e += TJ) {
        const int r = e / MP, idx = e - r * MP;
        int a = 0;
        while ((a + 1) * (a + 2) / 2 <= idx) ++a;
        const int k = idx - a * (a + 1) / 2;
        const int64_t gi = i0 + r;
        sR[r][idx] = (gi < p.row1) ? p.R[gi * DD + a * D + k] : 0.0;   // lower triangle of W_i, or of the symmetric S_i
    }
    if (KIND == GABO_SPD_STEIN && tid < TI) sRld[tid] = (i0 + tid < p.row1) ? p.Rld[i0 + tid] : 0.0;

    const int64_t j = j0 + tid;
    const bool jok = j < p.n2;
    double B[D][D];   // S2[j], upper storage used
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = r; c < D; ++c) B[r][c] = jok ? p.Cm[j * DD + r * D + c] : (r == c ? 1.0 : 0.0);
    double cld = 0.0;
    if (KIND == GABO_SPD_STEIN && jok) cld = p.Cld[j];
    __syncthreads();

    const int rows_here = (int)((p.row1 - i0) < (int64_t)TI ? (p.row1 - i0) : (int64_t)TI);
    for (int r = 0; r < rows_here; ++r) {
        const int64_t gi = i0 + r;
        // in the mirrored mode the strictly upper entries are produced by their transposes
        const bool skip = mirror && (j > gi);
        double val = 0.0;
        if (jok && !skip) {
            if (KIND == GABO_SPD_STEIN) {
                // log det(S_i + S_j) by an LDL^T elimination in registers (no square roots)
                double A[D][D];
#pragma unroll
                for (int a = 0; a < D; ++a)
#pragma unroll
                    for (int b = a; b < D; ++b) A[a][b] = sR[r][b * (b + 1) / 2 + a] + B[a][b];
                double piv[D];
                double prod = 1.0;
                bool bad = false;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    piv[k] = A[k][k];
                    bad = bad || !(piv[k] > 0.0);
                    prod *= piv[k];
                    const double rp = 1.0 / piv[k];
#pragma unroll
                    for (int a = k + 1; a < D; ++a) {
                        const double l = A[k][a] * rp;
#pragma unroll
                        for (int b = a; b < D; ++b) A[a][b] = fma(-l, A[k][b], A[a][b]);
                    }
                }
                double lgdet;
                if (prod > 1e-280 && prod < 1e280) {
                    lgdet = log(prod);
                } else {   // pivot product out of range: sum the logs instead
                    lgdet = 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) lgdet += log(piv[k]);
                }
                const double ldsum = lgdet - (double)D * 0.69314718055994530942;   // log det((A+B)/2)
                // kernels_spd_tf.py:104-112: det(X X2)^beta / det((X+X2)/2)^beta
                val = p.variance * exp(p.beta * (sRld[r] + cld - ldsum));
                if (bad) val = __longlong_as_double(0x7ff8000000000000LL);
            } else {
                if (p.symmetric && gi == j) {
                    val = p.variance;   // d_AI(X,X) = 0 exactly
                } else {
                    // M = W_i S_j W_i^T, Householder (FP64), then FP32 QL + FP64 Newton refinement of the eigenvalues; the
                    // out-of-line slow path when that does not validate or the spectrum is wide (ai_sumlog2_slow)
                    double Mx[D][D], ev[D], td[D], te[D];
                    ai_congruence<D>(sR[r], B, Mx);
                    householder_tridiag<D>(Mx, td, te);
                    bool fast = tridiag_eigvals_mixed<D>(td, te, ev);
                    double emin = ev[0], emax = ev[0];
#pragma unroll
                    for (int k = 1; k < D; ++k) { emin = fmin(emin, ev[k]); emax = fmax(emax, ev[k]); }
                    fast = fast && (emin > 0.02 * emax);
                    double s2 = 0.0;
                    if (fast) {
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            const double lg = log(ev[k]);   // SPD_utils_tf.py:136-139
                            s2 = fma(lg, lg, s2);
                        }
                    } else {
                        SymPacked<D> Sp;
#pragma unroll
                        for (int rr = 0; rr < D; ++rr)
#pragma unroll
                            for (int cc = rr; cc < D; ++cc) Sp.v[rr * D - rr * (rr - 1) / 2 + (cc - rr)] = B[rr][cc];
                        s2 = ai_sumlog2_slow<D>(sR[r], Sp);
                    }
                    if (KIND == GABO_SPD_AI_GAUSSIAN) val = p.variance * exp(-p.beta * s2);          // kernels_spd_tf.py:241-248
                    else val = p.variance * exp(-p.beta * sqrt(s2));                                 // kernels_spd_tf.py:396-400
                }
            }
            st_cs(p.K + (gi - p.row0) * p.ldk + j, val);
        }
        if (mirror) sOut[r][tid] = val;
    }
    if (mirror) {
        __syncthreads();
        const int lane = tid & 31, warp = tid >> 5;
        for (int jl = warp; jl < TJ; jl += TJ / 32) {
            const int64_t gj = j0 + jl;          // becomes the row of the mirrored entry
            const int64_t gi = i0 + lane;        // becomes the column
            if (gj < p.n2 && lane < rows_here && gj < gi) st_cs(p.K + gj * p.ldk + gi, sOut[lane][jl]);
        }
    }
}

// ---- affine-invariant pair kernel, second generation ----------------------------------------------------------------------
// Measured on the first version (profiles/r01_spd_gram_mixed_ncu.txt): issue-bound at 3.3 k warp instructions per pair-warp,
// 41 % of them the FP32 QL start (rsqrtf()/__fdividef() denormal guards, convergence tests, a 6x unrolled iteration loop)
// and a 67 KB loop body that does not fit the 32 KB L1.5 instruction cache (stall_no_instruction 1.3 per issue).  Here:
//   * two rows per trip; their FP32 QL chains share packed FFMA2 instructions (ql_tri_eigvals_start_f32x2);
//   * the FP64 phases are ROLLED over the two rows (results pass through a per-thread shared-memory stash), the QL
//     iteration loop is rolled, log/exp are the lean versions of gabo_math.cuh: loop body ~25 KB;
//   * the mirror transpose buffer holds 8 rows (64 B mirrored segments = whole sectors), 38 KB of shared memory per CTA.
constexpr int AI_TH = 8;           // rows per mirror-transpose phase
constexpr int AI_SLD = TJ + 1;     // padded leading dimension of the transpose buffer

template <int D>
struct AiSmem {
    static constexpr int MP = D * (D + 1) / 2;
    static constexpr int NST = 2 * D;                       // stash per (row of the pair, thread): ds[D], es[D-1], trace
    static constexpr size_t w_off = 0;                                              // [TI][MP]
    static constexpr size_t stash_off = w_off + sizeof(double) * TI * MP;           // [2][NST][TJ]
    static constexpr size_t out_off = stash_off + sizeof(double) * 2 * NST * TJ;    // [AI_TH][AI_SLD]
    static constexpr size_t tab_off = out_off + sizeof(double) * AI_TH * AI_SLD;    // [32]
    static constexpr size_t bytes = tab_off + sizeof(double) * 32;
};

// slow path of one pair with S_j handed over as a packed upper triangle (see ai_sumlog2_slow)
template <int D, int KIND>
__global__ void __launch_bounds__(TJ, (D <= 6) ? 4 : 2) spd_ai_gram_kernel(const SpdGramParams p) {
    constexpr int DD = D * D;
    constexpr int MP = D * (D + 1) / 2;
    constexpr int NST = AiSmem<D>::NST;
    extern __shared__ __align__(16) unsigned char ai_smem_raw[];
    double (*sW)[MP] = reinterpret_cast<double (*)[MP]>(ai_smem_raw + AiSmem<D>::w_off);
    double* const sStash = reinterpret_cast<double*>(ai_smem_raw + AiSmem<D>::stash_off);
    double (*sOut)[AI_SLD] = reinterpret_cast<double (*)[AI_SLD]>(ai_smem_raw + AiSmem<D>::out_off);
    double* const sTab = reinterpret_cast<double*>(ai_smem_raw + AiSmem<D>::tab_off);

    const int tid = threadIdx.x;
    const int64_t i0 = p.row0 + (int64_t)blockIdx.y * TI;
    const int64_t j0 = (int64_t)blockIdx.x * TJ;
    if (p.uplo != GABO_FULL && j0 > i0 + TI - 1) return;   // symmetric modes: blocks strictly above the diagonal
    const bool mirror = (p.uplo == GABO_SYMM_MIRROR);

    for (int e = tid; e < TI * MP; e += TJ) {
        const int r = e / MP, idx = e - r * MP;
        int a = 0;
        while ((a + 1) * (a + 2) / 2 <= idx) ++a;
        const int k = idx - a * (a + 1) / 2;
        const int64_t gi = i0 + r;
        sW[r][idx] = (gi < p.row1) ? p.R[gi * DD + a * D + k] : (a == k ? 1.0 : 0.0);   // W_i (identity past the last row)
    }
    if (tid < 32) sTab[tid] = kExp2Tab[tid] * p.variance;

    const int64_t j = j0 + tid;
    const bool jok = j < p.n2;
    double B[D][D];   // S2[j], upper storage used
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = r; c < D; ++c) B[r][c] = jok ? p.Cm[j * DD + r * D + c] : (r == c ? 1.0 : 0.0);
    __syncthreads();

    const double nbeta = -p.beta, variance = p.variance;
    const int rows_here = (int)((p.row1 - i0) < (int64_t)TI ? (p.row1 - i0) : (int64_t)TI);
    double* const stash0 = sStash + tid;   // entry k of row u of the pair: stash0[(u * NST + k) * TJ]

#pragma unroll 1
    for (int hb = 0; hb < TI; hb += AI_TH) {
        if (hb >= rows_here) break;
#pragma unroll 1
        for (int r = hb; r < hb + AI_TH; r += 2) {
            // lanes that have a pair to evaluate in row r / r + 1 (the rest computes along on benign data and is discarded)
            bool act[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int64_t gi = i0 + r + u;
                act[u] = jok && (r + u < rows_here) && !(mirror && j > gi) && !(p.symmetric && gi == j);
            }
            double val[2] = {0.0, 0.0};
            if (__any_sync(0xffffffffu, act[0] || act[1])) {
                // phase A (FP64, rolled over the two rows): congruence, Householder, trace normalisation -> stash
#pragma unroll 1
                for (int u = 0; u < 2; ++u) {
                    double Mx[D][D], td[D], te[D];
                    ai_congruence<D>(sW[r + u], B, Mx);
                    householder_tridiag<D>(Mx, td, te);
                    double tr = 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) tr += td[k];
                    const double rt = rcp_nr2(tr);
                    double* st = stash0 + (size_t)u * NST * TJ;
#pragma unroll
                    for (int k = 0; k < D; ++k) st[k * TJ] = td[k] * rt;
#pragma unroll
                    for (int k = 0; k < D - 1; ++k) st[(D + k) * TJ] = te[k] * rt;
                    st[(2 * D - 1) * TJ] = tr;
                }
                // phase B (FP32, both rows in packed halves): implicit-shift QL for the starting values
                f32x2_t dp[D], ep[D];
#pragma unroll
                for (int k = 0; k < D; ++k) dp[k] = f2_pack((float)stash0[k * TJ], (float)stash0[(NST + k) * TJ]);
#pragma unroll
                for (int k = 0; k < D - 1; ++k) ep[k] = f2_pack((float)stash0[(D + k) * TJ], (float)stash0[(NST + D + k) * TJ]);
                ep[D - 1] = 0ull;
                ql_tri_eigvals_start_f32x2<D>(dp, ep);
                // phase C (FP64, rolled): Newton refinement, validation, sum of squared logs
                double arg[2];
#pragma unroll 1
                for (int u = 0; u < 2; ++u) {
                    const double* st = stash0 + (size_t)u * NST * TJ;
                    double ds[D], e2[D], ev[D];
                    float x0[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        ds[k] = st[k * TJ];
                        x0[k] = u ? f2_hi(dp[k]) : f2_lo(dp[k]);
                    }
#pragma unroll
                    for (int k = 0; k < D - 1; ++k) { const double es = st[(D + k) * TJ]; e2[k] = es * es; }
                    e2[D - 1] = 0.0;
                    const double tr = st[(2 * D - 1) * TJ];
                    bool fast = tridiag_newton_refine<D>(ds, e2, x0, ev);
                    double emin = ev[0], emax = ev[0];
#pragma unroll
                    for (int k = 1; k < D; ++k) { emin = fmin(emin, ev[k]); emax = fmax(emax, ev[k]); }
                    fast = fast && (emin > 0.02 * emax) && (tr > 1e-280) && (tr < 1e280);
                    double s2 = 0.0;
                    if (fast) {
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            const double lg = log_pos_normal(ev[k] * tr);   // SPD_utils_tf.py:136-139
                            s2 = fma(lg, lg, s2);
                        }
                    } else if (u ? act[1] : act[0]) {
                        SymPacked<D> Sp;
#pragma unroll
                        for (int rr = 0; rr < D; ++rr)
#pragma unroll
                            for (int cc = rr; cc < D; ++cc) Sp.v[rr * D - rr * (rr - 1) / 2 + (cc - rr)] = B[rr][cc];
                        s2 = ai_sumlog2_slow<D>(sW[r + u], Sp);
                    }
                    const double a = (KIND == GABO_SPD_AI_GAUSSIAN) ? nbeta * s2 : nbeta * sqrt(s2);   // kernels_spd_tf.py:241-248 / 396-400
                    if (u) arg[1] = a; else arg[0] = a;
                }
                exp_neg_scaled_fast_n<2>(arg, val, sTab, variance);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int64_t gi = i0 + r + u;
                if (p.symmetric && gi == j) val[u] = variance;   // d_AI(X,X) = 0 exactly
                const bool skip = mirror && (j > gi);
                if (jok && (r + u < rows_here) && !skip) st_cs(p.K + (gi - p.row0) * p.ldk + j, val[u]);
                if (mirror) sOut[r + u - hb][tid] = val[u];
            }
        }
        if (mirror) {
            // transposed write of this 8-row strip: 8 lanes per mirrored row, 64 B contiguous
            __syncthreads();
            const int lane = tid & 31, warp = tid >> 5;
            const int cr = lane & 7;
            for (int jl = warp * 4 + (lane >> 3); jl < TJ; jl += 16) {
                const int64_t gj = j0 + jl;          // becomes the row of the mirrored entry
                const int64_t gi = i0 + hb + cr;     // becomes the column
                if (gj < p.n2 && hb + cr < rows_here && gj < gi) st_cs(p.K + gj * p.ldk + gi, sOut[cr][jl]);
            }
            __syncthreads();
        }
    }
}

template <int D, int KIND>
int launch_spd_ai(const SpdGramParams& p, dim3 grid, cudaStream_t stream) {
    static bool attr_set = false;   // idempotent; a race only repeats the call
    if (!attr_set) {
        GABO_CUDA_TRY(cudaFuncSetAttribute(spd_ai_gram_kernel<D, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)AiSmem<D>::bytes));
        attr_set = true;
    }
    spd_ai_gram_kernel<D, KIND><<<grid, TJ, AiSmem<D>::bytes, stream>>>(p);
    return 0;
}

template <int D>
int launch_spd_gram(int kind, const SpdGramParams& p, cudaStream_t stream) {
    dim3 grid((unsigned)ceil_div64(p.n2, TJ), (unsigned)ceil_div64(p.row1 - p.row0, TI));
    static const bool ai_v1 = (getenv("GABO_SPD_AI_V1") != nullptr);   // A/B switch for profiling: first-generation kernel
    int rc = 0;
    if (kind == GABO_SPD_STEIN) spd_gram_kernel<D, GABO_SPD_STEIN><<<grid, TJ, 0, stream>>>(p);
    else if (ai_v1 && kind == GABO_SPD_AI_GAUSSIAN) spd_gram_kernel<D, GABO_SPD_AI_GAUSSIAN><<<grid, TJ, 0, stream>>>(p);
    else if (ai_v1) spd_gram_kernel<D, GABO_SPD_AI_LAPLACE><<<grid, TJ, 0, stream>>>(p);
    else if (kind == GABO_SPD_AI_GAUSSIAN) rc = launch_spd_ai<D, GABO_SPD_AI_GAUSSIAN>(p, grid, stream);
    else rc = launch_spd_ai<D, GABO_SPD_AI_LAPLACE>(p, grid, stream);
    if (rc) return rc;
    GABO_LAUNCH_CHECK();
    return 0;
}

__global__ void stein_kdiag_kernel(const double* __restrict__ logdet, int64_t n, double beta, double variance,
                                   double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = variance * exp(beta * logdet[i]);
}

}  // namespace
}  // namespace gabo

extern "C" int gabo_spd_prep_f64(const double* Xmandel, int64_t n, int64_t ldx, int d, double* out_S, double* out_W,
                                 double* out_logdet, double* out_logm_mandel, int32_t* info, gabo_stream_t stream) {
    using namespace gabo;
    if (n < 0) return -2;
    if (d < 2 || d > 8) return -4;
    if (ldx < d * (d + 1) / 2) return -3;
    if (n == 0) return 0;
    if (Xmandel == nullptr) return -1;
    if (info == nullptr) return -9;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int threads = 64;
    const unsigned grid = (unsigned)ceil_div64(n, threads);
#define GABO_PREP_CASE(DV)                                                                                     \
    case DV:                                                                                                   \
        spd_prep_kernel<DV><<<grid, threads, 0, s>>>(Xmandel, n, ldx, out_S, out_W, out_logdet, out_logm_mandel, info); \
        break;
    switch (d) {
        GABO_PREP_CASE(2) GABO_PREP_CASE(3) GABO_PREP_CASE(4) GABO_PREP_CASE(5)
        GABO_PREP_CASE(6) GABO_PREP_CASE(7) GABO_PREP_CASE(8)
    }
#undef GABO_PREP_CASE
    GABO_LAUNCH_CHECK();
    return 0;
}

extern "C" int gabo_spd_gram_f64(int kind, int d, const double* W1, const double* S1, const double* logdet1, int64_t n1,
                                 const double* S2, const double* logdet2, int64_t n2, int symmetric, double beta,
                                 double variance, double* K, int64_t ldk, int64_t row0, int64_t row1, int uplo,
                                 gabo_stream_t stream) {
    using namespace gabo;
    if (kind < GABO_SPD_AI_GAUSSIAN || kind > GABO_SPD_STEIN) return -1;
    if (d < 2 || d > 8) return -2;
    if (n1 < 0) return -6;
    if (n2 < 0) return -9;
    if (kind == GABO_SPD_STEIN) {
        if (S1 == nullptr) return -4;
        if (logdet1 == nullptr) return -5;
        if (logdet2 == nullptr) return -8;
    } else if (W1 == nullptr) {
        return -3;
    }
    if (S2 == nullptr) return -7;
    if (K == nullptr) return -13;
    if (ldk < n2) return -14;
    if (row0 < 0 || row0 > n1 || row0 % 32 != 0) return -15;
    if (row1 < row0 || row1 > n1) return -16;
    if (uplo < GABO_FULL || uplo > GABO_SYMM_MIRROR) return -17;
    if (uplo != GABO_FULL && !symmetric) return -17;
    if (uplo == GABO_SYMM_MIRROR && !(row0 == 0 && row1 == n1)) return -17;
    if (symmetric && n1 != n2) return -10;
    if (row1 == row0 || n2 == 0) return 0;
    SpdGramParams p;
    p.R = (kind == GABO_SPD_STEIN) ? S1 : W1;
    p.Rld = logdet1; p.Cm = S2; p.Cld = logdet2; p.K = K; p.ldk = ldk; p.n1 = n1; p.n2 = n2;
    p.row0 = row0; p.row1 = row1; p.symmetric = symmetric ? 1 : 0; p.uplo = uplo; p.beta = beta; p.variance = variance;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (d) {
#ifndef GABO_SPD_ONLY_D6   /* development builds: compile the d = 6 pair kernels only (the full set takes ~100 s) */
        case 2: return launch_spd_gram<2>(kind, p, s);
        case 3: return launch_spd_gram<3>(kind, p, s);
        case 4: return launch_spd_gram<4>(kind, p, s);
        case 5: return launch_spd_gram<5>(kind, p, s);
        case 7: return launch_spd_gram<7>(kind, p, s);
        case 8: return launch_spd_gram<8>(kind, p, s);
#endif
        default: return launch_spd_gram<6>(kind, p, s);
    }
}

extern "C" int gabo_spd_stein_kdiag_f64(const double* logdet, int64_t n, double beta, double variance, double* out,
                                        gabo_stream_t stream) {
    if (n < 0) return -2;
    if (n == 0) return 0;
    if (logdet == nullptr) return -1;
    if (out == nullptr) return -5;
    gabo::stein_kdiag_kernel<<<(unsigned)gabo::ceil_div64(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        logdet, n, beta, variance, out);
    GABO_LAUNCH_CHECK();
    return 0;
}
