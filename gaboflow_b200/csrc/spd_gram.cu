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
                                double* __restrict__ outLogm, double* __restrict__ outLt, int64_t ldt,
                                int32_t* __restrict__ info) {
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
    if (outLt) {   // packed lower factor, transposed: entry (a,k) of point i at outLt[(a(a+1)/2 + k) * ldt + i] (coalesced per entry)
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int k = 0; k <= a; ++k) outLt[(int64_t)(a * (a + 1) / 2 + k) * ldt + i] = bad ? nanv : L[a][k];
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
    const double* Cm;       // Stein: column operand S2 [n2, d*d]
    const double* Lt;       // AI kinds: column operand, packed lower Cholesky factors, transposed [MP][ldt]
    const double* Cld;      // logdet of column points (Stein) or nullptr
    double* K;
    int64_t ldk, ldt, n1, n2, row0, row1;
    int symmetric, uplo;
    double beta, variance;
    // multi-GPU mode (cyc_world > 0; symmetric, mirrored): K is distributed by 512-row panels, 1-D block-cyclic -- global row
    // gi lives on rank (gi / 512) % cyc_world at local row (gi / 512 / cyc_world) * 512 + gi % 512 of peer[rank] (this
    // GPU's own buffer or an NVLink peer mapping).  The lower 32 x 128 blocks are dealt to the ranks round-robin
    // ((blockIdx.x + blockIdx.y) % cyc_world), so the work is balanced whatever the panel ownership, and every block is
    // stored -- rows and mirror image -- straight into the HBM of the ranks that own those rows.
    int cyc_world, cyc_rank;
    double* peer[GABO_MAX_PEERS];
};

constexpr int SPD_CYC_NB = 512;   // panel height of the block-cyclic layout (gaboflow_b200/dist.py)

// start of global row gi of K
__device__ __forceinline__ double* spd_row_ptr(const SpdGramParams& p, int64_t gi) {
    if (p.cyc_world > 0) {
        const int64_t P = gi / SPD_CYC_NB;
        return p.peer[P % p.cyc_world] + ((P / p.cyc_world) * SPD_CYC_NB + (gi - P * SPD_CYC_NB)) * p.ldk;
    }
    return p.K + (gi - p.row0) * p.ldk;
}
__device__ __forceinline__ bool spd_block_is_mine(const SpdGramParams& p) {
    return p.cyc_world <= 0 || (int)((blockIdx.x + blockIdx.y) % (unsigned)p.cyc_world) == p.cyc_rank;
}

// M = (W L)(W L)^T for one pair: W = chol(S_i)^-1 and L = chol(S_j), both lower triangular and packed (entry (a,k), k <= a,
// at a(a+1)/2 + k).  G = W L is lower triangular again, so the congruence W S_j W^T costs 2 * d(d+1)(d+2)/6 = 112 FMA at
// d = 6 (182 through the dense S_j) and never holds more than two triangles in registers.  Mx: upper storage.
template <int D>
__device__ __forceinline__ void ai_congruence(const double* __restrict__ wrow, const double (&Lj)[D * (D + 1) / 2],
                                              double (&Mx)[D][D]) {
    double G[D][D];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int c = 0; c <= a; ++c) {
            double v = wrow[a * (a + 1) / 2 + c] * Lj[c * (c + 1) / 2 + c];
#pragma unroll
            for (int k = c + 1; k <= a; ++k) v = fma(wrow[a * (a + 1) / 2 + k], Lj[k * (k + 1) / 2 + c], v);
            G[a][c] = v;
        }
#pragma unroll
    for (int b = 0; b < D; ++b)
#pragma unroll
        for (int a = b; a < D; ++a) {
            double v = G[a][0] * G[b][0];
#pragma unroll
            for (int k = 1; k <= b; ++k) v = fma(G[a][k], G[b][k], v);
            Mx[b][a] = v;
        }
}

// Out-of-line slow path of the affine-invariant pair (kept out of the hot loop: ~3 k instructions).  Taken when the
// mixed-precision eigenvalues do not validate (near-multiple eigenvalues) or the spectrum is wide: all-FP64 QL on the
// tridiagonal, then cyclic Jacobi (relative accuracy) if lambda_min <= 0.02 lambda_max.  L_j travels by value (packed)
// so that the caller's copy can stay in registers.  Returns sum_k log^2 lambda_k.
template <int D>
struct TriPacked { double v[D * (D + 1) / 2]; };

template <int D>
__device__ __noinline__ double ai_sumlog2_slow(const double* __restrict__ wrow, const TriPacked<D> Lp) {
    double Mx[D][D], ev[D];
    ai_congruence<D>(wrow, Lp.v, Mx);
    tridiag_ql_eigvals<D>(Mx, ev);
    double emin = ev[0], emax = ev[0];
#pragma unroll
    for (int k = 1; k < D; ++k) { emin = fmin(emin, ev[k]); emax = fmax(emax, ev[k]); }
    if (!(emin > 0.02 * emax)) {
        ai_congruence<D>(wrow, Lp.v, Mx);
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

// ---- Stein pair kernel: thread = one column j with S_j in registers, 32 rows per CTA -------------------------------------------
template <int D>
__global__ void __launch_bounds__(TJ, (D <= 6) ? 4 : 2) spd_stein_gram_kernel(const SpdGramParams p) {
    constexpr int DD = D * D;
    constexpr int MP = D * (D + 1) / 2;   // packed lower triangle: entry (a,k), k <= a, at a(a+1)/2 + k
    __shared__ double sR[TI][MP];
    __shared__ double sRld[TI];
    __shared__ double sOut[TI][TJ + 1];

    const int tid = threadIdx.x;
    const int64_t i0 = p.row0 + (int64_t)blockIdx.y * TI;
    const int64_t j0 = (int64_t)blockIdx.x * TJ;
    if (p.uplo != GABO_FULL && j0 > i0 + TI - 1) return;   // symmetric modes: blocks strictly above the diagonal
    if (!spd_block_is_mine(p)) return;
    const bool mirror = (p.uplo == GABO_SYMM_MIRROR);

    for (int e = tid; e < TI * MP; e += TJ) {
        const int r = e / MP, idx = e - r * MP;
        int a = 0;
        while ((a + 1) * (a + 2) / 2 <= idx) ++a;
        const int k = idx - a * (a + 1) / 2;
        const int64_t gi = i0 + r;
        sR[r][idx] = (gi < p.row1) ? p.R[gi * DD + a * D + k] : 0.0;   // lower triangle of the symmetric S_i
    }
    if (tid < TI) sRld[tid] = (i0 + tid < p.row1) ? p.Rld[i0 + tid] : 0.0;

    const int64_t j = j0 + tid;
    const bool jok = j < p.n2;
    double B[D][D];   // S2[j], upper storage used
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = r; c < D; ++c) B[r][c] = jok ? p.Cm[j * DD + r * D + c] : (r == c ? 1.0 : 0.0);
    const double cld = jok ? p.Cld[j] : 0.0;
    __syncthreads();

    const int rows_here = (int)((p.row1 - i0) < (int64_t)TI ? (p.row1 - i0) : (int64_t)TI);
    for (int r = 0; r < rows_here; ++r) {
        const int64_t gi = i0 + r;
        const bool skip = mirror && (j > gi);   // in the mirrored mode the strictly upper entries are produced by their transposes
        double val = 0.0;
        if (jok && !skip) {
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
            st_cs(spd_row_ptr(p, gi) + j, val);
        }
        if (mirror) sOut[r][tid] = val;
    }
    if (mirror) {
        __syncthreads();
        const int lane = tid & 31, warp = tid >> 5;
        for (int jl = warp; jl < TJ; jl += TJ / 32) {
            const int64_t gj = j0 + jl;          // becomes the row of the mirrored entry
            const int64_t gi = i0 + lane;        // becomes the column
            if (gj < p.n2 && lane < rows_here && gj < gi) st_cs(spd_row_ptr(p, gj) + gi, sOut[lane][jl]);
        }
    }
}

// ---- affine-invariant pair kernel ----------------------------------------------------------------------------------------
// History, measured at N = 16384, d = 6 (profiles/r01_spd_gram_mixed_ncu.txt, profiles/r02_spd_ai_*.txt):
//   v1  one row per trip, scalar FP32 QL (rsqrtf()/__fdividef() denormal guards, six unrolled iterations per level): 3.3 k warp
//       instructions per pair, a 67 KB loop body that misses the 32 KB L1.5 instruction cache (stall_no_instruction 1.3 per
//       issue), S_j held in 42 registers: 20.3 ms.
//   v2  two rows per trip with their QL chains in packed FFMA2 halves, the FP64 phases ROLLED over the rows of a trip (results
//       pass through a per-thread shared-memory stash), rolled QL iterations, lean log / exp: 2.3 k instructions, 33 KB: 14.3 ms.
//   v3  four rows per trip (two packed chains for instruction-level parallelism in the QL phase): MUFU latency hidden, but a
//       46 KB loop body and more local-memory traffic: 14.5 ms.
//   v4  (this one) two rows per trip; the congruence W_i S_j W_i^T is accumulated as a sum of outer products of the columns of
//       G = W_i L_j (L_j = chol(S_j), lower: 112 instead of 182 FMA), with the columns of L_j read from L1 / L2 when needed
//       (transposed layout: coalesced) instead of S_j living in 42 registers through phases that never use it.
constexpr int AI_NR = 2;           // rows per trip
constexpr int AI_TH = 8;           // rows per mirror-transpose phase
constexpr int AI_SLD = TJ + 1;     // padded leading dimension of the mirror transpose buffer

template <int D>
struct AiSmem {
    static constexpr int MP = D * (D + 1) / 2;
    static constexpr int NST = 2 * D;                       // stash per (row of the trip, thread): ds[D], es[D-1], trace
    static constexpr size_t w_off = 0;                                                  // [TI][MP]
    static constexpr size_t stash_off = w_off + sizeof(double) * TI * MP;               // [AI_NR][NST][TJ]
    static constexpr size_t out_off = stash_off + sizeof(double) * AI_NR * NST * TJ;    // [AI_TH][AI_SLD]
    static constexpr size_t tab_off = out_off + sizeof(double) * AI_TH * AI_SLD;        // [32]
    static constexpr size_t bytes = tab_off + sizeof(double) * 32;
};

// a load that the compiler may neither hoist out of the row loop nor keep live across it
__device__ __forceinline__ double ld_global_keep(const double* ptr) {
    double v;
    asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(ptr));
    return v;
}

// M = (W L)(W L)^T accumulated over the columns g_c of G = W L: M += g_c g_c^T.  w: packed lower W_i in shared memory;
// lt: entry (k,c) of L_j at lt[(k(k+1)/2 + c) * ldt] (global, L1-resident).  Mx: upper storage.
template <int D>
__device__ __forceinline__ void ai_congruence_cols(const double* __restrict__ w, const double* lt, int64_t ldt,
                                                   double (&Mx)[D][D]) {
#pragma unroll
    for (int c = 0; c < D; ++c) {
        double lc[D], g[D];
#pragma unroll
        for (int k = c; k < D; ++k) lc[k] = ld_global_keep(lt + (int64_t)(k * (k + 1) / 2 + c) * ldt);
#pragma unroll
        for (int a = c; a < D; ++a) {
            double v = w[a * (a + 1) / 2 + c] * lc[c];
#pragma unroll
            for (int k = c + 1; k <= a; ++k) v = fma(w[a * (a + 1) / 2 + k], lc[k], v);
            g[a] = v;
        }
#pragma unroll
        for (int b = c; b < D; ++b)
#pragma unroll
            for (int a = b; a < D; ++a) Mx[b][a] = (c == 0) ? g[a] * g[b] : fma(g[a], g[b], Mx[b][a]);
    }
}

template <int D, int KIND>
__global__ void __launch_bounds__(TJ, (D <= 6) ? 5 : 2) spd_ai_gram_kernel(const SpdGramParams p) {
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
    if (!spd_block_is_mine(p)) return;
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
    __syncthreads();

    const int64_t j = j0 + tid;
    const bool jok = j < p.n2;
    const double* const ltj = p.Lt + (jok ? j : 0);        // lanes past the last column compute along on column 0
    const double nbeta = -p.beta, variance = p.variance;
    const int exp_slow_hi = exp_slow_threshold_hi(variance);
    const int rows_here = (int)((p.row1 - i0) < (int64_t)TI ? (p.row1 - i0) : (int64_t)TI);
    double* const stash0 = sStash + tid;   // entry k of row u of the trip: stash0[(u * NST + k) * TJ]

#pragma unroll 1
    for (int hb = 0; hb < TI; hb += AI_TH) {
        if (hb >= rows_here) break;
#pragma unroll 1
        for (int r = hb; r < hb + AI_TH; r += AI_NR) {
            // lanes that have a pair to evaluate in row r / r + 1 (the rest computes along on benign data and is discarded)
            bool act[AI_NR];
#pragma unroll
            for (int u = 0; u < AI_NR; ++u) {
                const int64_t gi = i0 + r + u;
                act[u] = jok && (r + u < rows_here) && !(mirror && j > gi) && !(p.symmetric && gi == j);
            }
            double val[AI_NR] = {0.0, 0.0};
            if (__any_sync(0xffffffffu, act[0] || act[1])) {
                // phase A (FP64, rolled over the two rows): congruence, Householder, trace normalisation -> stash
#pragma unroll 1
                for (int u = 0; u < AI_NR; ++u) {
                    double Mx[D][D], td[D], te[D];
                    ai_congruence_cols<D>(sW[r + u], ltj, p.ldt, Mx);
                    householder_tridiag<D>(Mx, td, te);
                    double t = 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) t += td[k];
                    const double rt = rcp_nr2(t);
                    double* st = stash0 + (size_t)u * NST * TJ;
#pragma unroll
                    for (int k = 0; k < D; ++k) st[k * TJ] = td[k] * rt;
#pragma unroll
                    for (int k = 0; k < D - 1; ++k) st[(D + k) * TJ] = te[k] * rt;
                    st[(2 * D - 1) * TJ] = t;
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
                double arg[AI_NR];
#pragma unroll 1
                for (int u = 0; u < AI_NR; ++u) {
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
                    const double t = st[(2 * D - 1) * TJ];
                    bool fast = tridiag_newton_refine<D>(ds, e2, x0, ev);
                    // every eigenvalue above 1 % of the trace (the refinement is accurate to ~1e-16 of the TRACE; a smaller
                    // eigenvalue needs the relative accuracy of the slow path); tested on the exponent words
                    bool wide = false;
#pragma unroll
                    for (int k = 0; k < D; ++k) wide = wide || (__double2hiint(ev[k]) < 0x3f847ae1);
                    fast = fast && !wide && (t > 1e-280) && (t < 1e280);
                    double s2 = 0.0;
                    if (fast) {
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            const double lg = log_pos_normal(ev[k] * t);   // SPD_utils_tf.py:136-139
                            s2 = fma(lg, lg, s2);
                        }
                    } else if (u ? act[1] : act[0]) {
                        TriPacked<D> Lp;
#pragma unroll
                        for (int idx = 0; idx < MP; ++idx) Lp.v[idx] = ltj[(int64_t)idx * p.ldt];
                        s2 = ai_sumlog2_slow<D>(sW[r + u], Lp);
                    }
                    const double a = (KIND == GABO_SPD_AI_GAUSSIAN) ? nbeta * s2 : nbeta * sqrt(s2);   // kernels_spd_tf.py:241-248 / 396-400
                    if (u) arg[1] = a; else arg[0] = a;
                }
                exp_neg_scaled_fast_n<AI_NR>(arg, val, sTab, variance, exp_slow_hi);
            }
#pragma unroll
            for (int u = 0; u < AI_NR; ++u) {
                const int64_t gi = i0 + r + u;
                if (p.symmetric && gi == j) val[u] = variance;   // d_AI(X,X) = 0 exactly
                const bool skip = mirror && (j > gi);
                if (jok && (r + u < rows_here) && !skip) st_cs(spd_row_ptr(p, gi) + j, val[u]);
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
                if (gj < p.n2 && hb + cr < rows_here && gj < gi) st_cs(spd_row_ptr(p, gj) + gi, sOut[cr][jl]);
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
    int rc = 0;
    if (kind == GABO_SPD_STEIN) spd_stein_gram_kernel<D><<<grid, TJ, 0, stream>>>(p);
    else if (kind == GABO_SPD_AI_GAUSSIAN) rc = launch_spd_ai<D, GABO_SPD_AI_GAUSSIAN>(p, grid, stream);
    else rc = launch_spd_ai<D, GABO_SPD_AI_LAPLACE>(p, grid, stream);
    if (rc) return rc;
    GABO_LAUNCH_CHECK();
    return 0;
}

int dispatch_spd_gram(int kind, int d, const SpdGramParams& p, cudaStream_t s) {
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

// ---- candidate gradients of the affine-invariant kernels ------------------------------------------------------
// For k_i(X) = variance exp(-beta d_i^p), d_i^2 = sum_k log^2 lambda_k(M_i), M_i = W_i X W_i^T = V Lambda V^T  (W_i = chol(A_i)^-1):
//   d(d_i^2) = 2 tr( log(A_i^-1 X) X^-1 dX )  and  log(A_i^-1 X) X^-1 = W_i^T V diag(log lambda_k / lambda_k) V^T W_i,
// so the Euclidean gradient of d_i^2 with respect to the symmetric matrix X is  G_i = 2 sum_k (log lambda_k / lambda_k) u_k u_k^T,
// u_k = W_i^T v_k  (equal to 2 X^-1 Log(X A_i^-1), which is symmetric).  In Mandel coordinates (SPD_utils_tf.py:27-63) the
// diagonal slots get G_rr and the off-diagonal slots sqrt(2) G_rc.  dk_i/dx = -beta k_i g_i (Gaussian), -beta k_i g_i / (2 d_i)
// (Laplace).  The posterior gradients are weighted sums over the training points, as for the sphere kernels
// (gabo_sphere_posterior_grad_f64):  d mean/dx = sum_i alpha_i dk_i/dx,  d var/dx = -2 sum_i (Ky^-1 k)_i dk_i/dx.
// One CTA per candidate, a thread per training point (cyclic Jacobi WITH eigenvectors in registers), fixed-order reduction.
// The reference gets these from TensorFlow autodiff and hands them to the manifold optimiser (manifold_optimization.py:260-271).
template <int D, int KIND>
__global__ void __launch_bounds__(128) spd_ai_post_grad_kernel(int64_t n, const double* __restrict__ W1, const double* __restrict__ S2,
                                                               const double* __restrict__ alpha, const double* __restrict__ Bm,
                                                               int64_t ldb, double beta, double variance,
                                                               const double* __restrict__ mean, const double* __restrict__ var,
                                                               double fmin, double* __restrict__ gmean, double* __restrict__ gvar,
                                                               double* __restrict__ dei) {
    constexpr int DD = D * D;
    constexpr int MP = D * (D + 1) / 2;
    __shared__ double red[2][MP][128 + 1];
    const int64_t m = blockIdx.x;
    const int tid = threadIdx.x;
    double B[D][D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) B[r][c] = S2[m * DD + r * D + c];
    double a1[MP], a2[MP];
#pragma unroll
    for (int s = 0; s < MP; ++s) { a1[s] = 0.0; a2[s] = 0.0; }
    for (int64_t i = tid; i < n; i += 128) {
        double Wm[D][D];
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
            for (int c = 0; c < D; ++c) Wm[r][c] = (c <= r) ? W1[i * DD + r * D + c] : 0.0;
        // M = W B W^T (upper storage)
        double T[D][D], Mx[D][D], V[D][D];
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k <= a; ++k) v = fma(Wm[a][k], B[k][c], v);
                T[a][c] = v;
            }
#pragma unroll
        for (int b = 0; b < D; ++b)
#pragma unroll
            for (int a = b; a < D; ++a) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k <= b; ++k) v = fma(T[a][k], Wm[b][k], v);
                Mx[b][a] = v;
            }
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
            for (int c = 0; c < D; ++c) V[r][c] = (r == c) ? 1.0 : 0.0;
        jacobi_eig<D, true>(Mx, V);
        double s2 = 0.0, ck[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const double lg = log(Mx[k][k]);
            s2 = fma(lg, lg, s2);
            ck[k] = lg / Mx[k][k];
        }
        // U = W^T V (column k = u_k)
        double U[D][D];
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int k = 0; k < D; ++k) {
                double v = 0.0;
#pragma unroll
                for (int b = a; b < D; ++b) v = fma(Wm[b][a], V[b][k], v);
                U[a][k] = v;
            }
        double kv, scale;
        if (KIND == GABO_SPD_AI_GAUSSIAN) {
            kv = variance * exp(-beta * s2);
            scale = -beta * kv;
        } else {
            const double dist = sqrt(s2);
            kv = variance * exp(-beta * dist);
            scale = (dist > 0.0) ? -beta * kv / (2.0 * dist) : 0.0;     // not differentiable at d = 0 (as the kernel itself)
        }
        const double w1 = alpha[i] * scale, w2 = Bm[i * ldb + m] * scale;
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
            for (int c = r; c < D; ++c) {
                double g = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) g = fma(ck[k] * U[r][k], U[c][k], g);
                g *= (r == c) ? 2.0 : 2.0 * 1.41421356237309504880;
                const int slot = mandel_slot<D>(r, c);
                a1[slot] = fma(w1, g, a1[slot]);
                a2[slot] = fma(w2, g, a2[slot]);
            }
    }
#pragma unroll
    for (int s = 0; s < MP; ++s) { red[0][s][tid] = a1[s]; red[1][s][tid] = a2[s]; }
    __syncthreads();
    if (tid < MP) {
        double g1 = 0.0, g2 = 0.0;
        for (int t = 0; t < 128; ++t) { g1 += red[0][tid][t]; g2 += red[1][tid][t]; }
        g2 *= -2.0;
        gmean[m * MP + tid] = g1;
        gvar[m * MP + tid] = g2;
        if (dei != nullptr) {
            const double v = var[m];
            const double sd = sqrt(fmax(v, 1e-6));
            const double z = (fmin - mean[m]) / sd;
            const double cdf = 0.5 * erfc(-z * 0.70710678118654752440);
            const double phi = exp(-0.5 * z * z) * 0.39894228040143267794;
            dei[m * MP + tid] = -cdf * g1 + ((v > 1e-6) ? phi / (2.0 * sd) * g2 : 0.0);   // the variance floor has zero slope
        }
    }
}

template <int D>
int launch_spd_ai_post_grad(int kind, int64_t n, int64_t M, const double* W1, const double* S2, const double* alpha, const double* Bm,
                            int64_t ldb, double beta, double variance, const double* mean, const double* var, double fmin,
                            double* gmean, double* gvar, double* dei, cudaStream_t stream) {
    if (kind == GABO_SPD_AI_GAUSSIAN)
        spd_ai_post_grad_kernel<D, GABO_SPD_AI_GAUSSIAN><<<(unsigned)M, 128, 0, stream>>>(n, W1, S2, alpha, Bm, ldb, beta, variance, mean, var,
                                                                                         fmin, gmean, gvar, dei);
    else
        spd_ai_post_grad_kernel<D, GABO_SPD_AI_LAPLACE><<<(unsigned)M, 128, 0, stream>>>(n, W1, S2, alpha, Bm, ldb, beta, variance, mean, var,
                                                                                        fmin, gmean, gvar, dei);
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
                                 double* out_logdet, double* out_logm_mandel, double* out_Lt, int64_t ldt, int32_t* info,
                                 gabo_stream_t stream) {
    using namespace gabo;
    if (n < 0) return -2;
    if (d < 2 || d > 8) return -4;
    if (ldx < d * (d + 1) / 2) return -3;
    if (out_Lt != nullptr && ldt < n) return -10;
    if (n == 0) return 0;
    if (Xmandel == nullptr) return -1;
    if (info == nullptr) return -11;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int threads = 64;
    const unsigned grid = (unsigned)ceil_div64(n, threads);
#define GABO_PREP_CASE(DV)                                                                                     \
    case DV:                                                                                                   \
        spd_prep_kernel<DV><<<grid, threads, 0, s>>>(Xmandel, n, ldx, out_S, out_W, out_logdet, out_logm_mandel, out_Lt, ldt, info); \
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
                                 const double* S2, const double* L2t, int64_t ldt2, const double* logdet2, int64_t n2,
                                 int symmetric, double beta, double variance, double* K, int64_t ldk, int64_t row0,
                                 int64_t row1, int uplo, gabo_stream_t stream) {
    using namespace gabo;
    if (kind < GABO_SPD_AI_GAUSSIAN || kind > GABO_SPD_STEIN) return -1;
    if (d < 2 || d > 8) return -2;
    if (n1 < 0) return -6;
    if (n2 < 0) return -11;
    if (kind == GABO_SPD_STEIN) {
        if (S1 == nullptr) return -4;
        if (logdet1 == nullptr) return -5;
        if (S2 == nullptr) return -7;
        if (logdet2 == nullptr) return -10;
    } else {
        if (W1 == nullptr) return -3;
        if (L2t == nullptr) return -8;
        if (ldt2 < n2) return -9;
    }
    if (K == nullptr) return -15;
    if (ldk < n2) return -16;
    if (row0 < 0 || row0 > n1 || row0 % 32 != 0) return -17;
    if (row1 < row0 || row1 > n1) return -18;
    if (uplo < GABO_FULL || uplo > GABO_SYMM_MIRROR) return -19;
    if (uplo != GABO_FULL && !symmetric) return -19;
    if (uplo == GABO_SYMM_MIRROR && !(row0 == 0 && row1 == n1)) return -19;
    if (symmetric && n1 != n2) return -12;
    if (row1 == row0 || n2 == 0) return 0;
    SpdGramParams p;
    p.R = (kind == GABO_SPD_STEIN) ? S1 : W1;
    p.Rld = logdet1; p.Cm = S2; p.Lt = L2t; p.ldt = ldt2; p.Cld = logdet2; p.K = K; p.ldk = ldk; p.n1 = n1; p.n2 = n2;
    p.row0 = row0; p.row1 = row1; p.symmetric = symmetric ? 1 : 0; p.uplo = uplo; p.beta = beta; p.variance = variance;
    p.cyc_world = 0; p.cyc_rank = 0;
    for (int r = 0; r < GABO_MAX_PEERS; ++r) p.peer[r] = nullptr;
    return gabo::dispatch_spd_gram(kind, d, p, static_cast<cudaStream_t>(stream));
}

// Multi-GPU symmetric SPD Gram K(X, X), distributed by 512-row panels 1-D block-cyclic over `world` ranks (panel P on rank
// P % world, local slot P / world).  Every rank calls this with the same per-point data; rank `rank` evaluates its share of
// the lower 32 x 128 blocks (dealt round-robin, so the triangular work is balanced) and stores each block and its mirror
// image directly into the buffers of the ranks that own those rows (K_bases_host[r]: base of rank r's [local_rows, ldk]
// buffer as mapped in this process -- own memory or CUDA IPC peer mappings over NVLink).  No collective; the caller
// synchronises (stream + barrier, or the flag handshake of gaboflow_b200/dist.py) before the rows are read.
extern "C" int gabo_spd_gram_cyclic_f64(int kind, int d, const double* W1, const double* S1, const double* logdet1,
                                        const double* L1t, int64_t ldt, int64_t n, double beta, double variance,
                                        double* const* K_bases_host, int64_t ldk, int world, int rank, gabo_stream_t stream) {
    using namespace gabo;
    if (kind < GABO_SPD_AI_GAUSSIAN || kind > GABO_SPD_STEIN) return -1;
    if (d < 2 || d > 8) return -2;
    if (n < 0) return -8;
    if (kind == GABO_SPD_STEIN) {
        if (S1 == nullptr) return -4;
        if (logdet1 == nullptr) return -5;
    } else {
        if (W1 == nullptr) return -3;
        if (L1t == nullptr) return -6;
        if (ldt < n) return -7;
    }
    if (K_bases_host == nullptr) return -11;
    if (ldk < n) return -12;
    if (world < 1 || world > GABO_MAX_PEERS) return -13;
    if (rank < 0 || rank >= world) return -14;
    if (n == 0) return 0;
    SpdGramParams p;
    p.R = (kind == GABO_SPD_STEIN) ? S1 : W1;
    p.Rld = logdet1; p.Cm = S1; p.Lt = L1t; p.ldt = ldt; p.Cld = logdet1; p.K = K_bases_host[rank]; p.ldk = ldk;
    p.n1 = n; p.n2 = n; p.row0 = 0; p.row1 = n; p.symmetric = 1; p.uplo = GABO_SYMM_MIRROR; p.beta = beta; p.variance = variance;
    p.cyc_world = world; p.cyc_rank = rank;
    for (int r = 0; r < GABO_MAX_PEERS; ++r) {
        p.peer[r] = (r < world) ? K_bases_host[r] : nullptr;
        if (r < world && p.peer[r] == nullptr) return -11;
    }
    return gabo::dispatch_spd_gram(kind, d, p, static_cast<cudaStream_t>(stream));
}


// Euclidean gradients (Mandel coordinates) of the posterior mean / variance and of Expected Improvement with respect to each of M
// candidate SPD matrices, affine-invariant kernels.  W1 [n, d*d]: inverse Cholesky factors of the training matrices; S2 [M, d*d]:
// the candidates; alpha [n] = Ky^-1 (y - m); Bm [n, M] = Ky^-1 K(X, X*); mean / var [M] the posterior (needed for dei only).
// gmean, gvar, dei: [M, d(d+1)/2]; dei may be NULL.
extern "C" int gabo_spd_ai_posterior_grad_f64(int kind, int d, int64_t n, int64_t M, const double* W1, const double* S2,
                                              const double* alpha, const double* Bm, int64_t ldb, double beta, double variance,
                                              const double* mean, const double* var, double fmin, double* gmean, double* gvar,
                                              double* dei, gabo_stream_t stream) {
    using namespace gabo;
    if (kind != GABO_SPD_AI_GAUSSIAN && kind != GABO_SPD_AI_LAPLACE) return -1;
    if (d < 2 || d > 6) return -2;        // a thread holds W, M, V and U: d <= 6 (every S^d_++ of the reference examples is d = 2)
    if (n < 0) return -3;
    if (M < 0) return -4;
    if (M == 0) return 0;
    if (n > 0 && W1 == nullptr) return -5;
    if (S2 == nullptr) return -6;
    if (n > 0 && alpha == nullptr) return -7;
    if (n > 0 && Bm == nullptr) return -8;
    if (ldb < M) return -9;
    if (!(beta > 0.0)) return -10;
    if (!(variance > 0.0)) return -11;
    if (dei != nullptr && (mean == nullptr || var == nullptr)) return -12;
    if (gmean == nullptr) return -15;
    if (gvar == nullptr) return -16;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (d) {
        case 2: return launch_spd_ai_post_grad<2>(kind, n, M, W1, S2, alpha, Bm, ldb, beta, variance, mean, var, fmin, gmean, gvar, dei, s);
        case 3: return launch_spd_ai_post_grad<3>(kind, n, M, W1, S2, alpha, Bm, ldb, beta, variance, mean, var, fmin, gmean, gvar, dei, s);
        case 4: return launch_spd_ai_post_grad<4>(kind, n, M, W1, S2, alpha, Bm, ldb, beta, variance, mean, var, fmin, gmean, gvar, dei, s);
        case 5: return launch_spd_ai_post_grad<5>(kind, n, M, W1, S2, alpha, Bm, ldb, beta, variance, mean, var, fmin, gmean, gvar, dei, s);
        default: return launch_spd_ai_post_grad<6>(kind, n, M, W1, S2, alpha, Bm, ldb, beta, variance, mean, var, fmin, gmean, gvar, dei, s);
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
