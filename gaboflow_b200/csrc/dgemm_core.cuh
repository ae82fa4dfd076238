// FP64 tensor-pipe GEMM core for sm_100a:  C[m,n] -= op(A) * op(B),  accumulating on DMMA.8x8x4 (mma.sync m8n8k4).
//
// This is the trailing-update engine of the blocked Cholesky (SYRK-shaped when A == B), of the explicit-inverse TRSM
// panels, of the blocked triangular solves and of cov = K** - A^T A.  There is no tcgen05 FP64 kind, so the FP64
// "tensor pipe" on B200 is warp-level DMMA; measured peak 37.2 TFLOP/s (profiles/r01_fp64_probe.txt), shared with DFMA.
//
// CTA tile 128x64, 8 warps (4x2), 32x32 outputs per warp in registers (64 accumulator registers), two CTAs per SM so
// that one CTA's prologue / C read-modify-write epilogue / barrier skew is covered by the other CTA's DMMA stream
// (one 16-warp CTA per SM left the pipe idle 24% of the time: profiles/r01_gemm_v3_summary.txt).  BK = 32 per stage,
// 2-stage cp.async (LDGSTS) ring (110 KB per CTA).  Operand tiles sit in shared memory with a leading dimension = 4 (mod 16) doubles,
// which makes every fragment load of the m8n8k4 layout bank-conflict free for both operand orientations.
#pragma once
#include "gabo_common.cuh"

namespace gabo {

constexpr int GTM = 128;         // CTA tile rows
constexpr int GTN = 64;          // CTA tile columns
constexpr int GT = GTM;          // row granularity used by callers
constexpr int GBK = 32;          // k per stage
constexpr int GSTAGES = 2;
constexpr int GTHREADS = 256;
constexpr int GWN = GTN / 32;    // warps along n (2); warps along m = 4
constexpr int GEMM_GROUP_M = 16; // tile rows per rasterisation group
constexpr int LDK_MAJ = GBK + 4;     // 36: K-major tile [rows][36]
constexpr int LDM_MAJ = GTM + 4;     // 132: M-major A tile [32][132]
constexpr int LDN_MAJ = GTN + 4;     // 68:  N-major B tile [32][68]
constexpr int TILE_A_DOUBLES = (GTM * LDK_MAJ > GBK * LDM_MAJ) ? GTM * LDK_MAJ : GBK * LDM_MAJ;   // 2560
constexpr int TILE_B_DOUBLES = (GTN * LDK_MAJ > GBK * LDN_MAJ) ? GTN * LDK_MAJ : GBK * LDN_MAJ;   // 1280
constexpr int STAGE_DOUBLES = TILE_A_DOUBLES + TILE_B_DOUBLES;
constexpr size_t GEMM_SMEM_BYTES = (size_t)GSTAGES * STAGE_DOUBLES * sizeof(double);               // 92,160

struct GemmParams {
    const double* A; int64_t lda;
    const double* B; int64_t ldb;
    double* C; int64_t ldc;
    int64_t m, n, k;
    int lower;       // 1: only entries with global col <= global row are produced (tiles above the diagonal are skipped)
    // global coordinates for the mask: gcol = col_base + col; grow = row_base + row, or, for rows stored block-cyclically
    // (cyc_nb > 0: local row block b = row / cyc_nb is global block b * cyc_world + cyc_rank), the mapped global row
    int64_t row_base, col_base;
    int cyc_nb, cyc_world, cyc_rank;
    int vecA, vecB, vecC;   // 16-byte access allowed (even leading dimension, 16 B aligned base)
    int tiles_m, tiles_n;   // tile grid; CTAs are rasterised in groups of GEMM_GROUP_M tile rows (see gemm_kernel)
    double alpha;    // C <- C + alpha * op(A) op(B)   (alpha = -1 for the updates; +1 with beta0 for products)
    int beta0;       // 1: C <- alpha * op(A) op(B)  (C not read)
};

__device__ __forceinline__ void cp_async_bytes(double* dst, const double* src, int bytes16, bool vec) {
    // bytes16: number of valid source bytes in this 16-byte chunk (0, 8 or 16); the rest is zero filled
    const uint32_t d = smem_u32(dst);
    if (vec) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(bytes16) : "memory");
    } else {
        const int b0 = bytes16 >= 8 ? 8 : 0, b1 = bytes16 >= 16 ? 8 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(b0) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d + 8), "l"(src + 1), "r"(b1) : "memory");
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// Loads one operand tile (ROWS x GBK) for k-block k0 into shared memory.
//  KMAJ: global element (r, kk) at G[(r0 + r) * ld + k0 + kk]      -> smem [r][kk], ld 20
// !KMAJ: global element (kk, r) at G[(k0 + kk) * ld + r0 + r]      -> smem [kk][r], ld ROWS + 4
template <bool KMAJ, int ROWS>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ G, int64_t ld, int64_t r0, int64_t rmax,
                                          int64_t k0, int64_t kmax, bool vec, int tid) {
    constexpr int CHUNKS = ROWS * GBK / 2;   // 16-byte chunks
#pragma unroll
    for (int it = 0; it < CHUNKS / GTHREADS; ++it) {
        const int c = tid + it * GTHREADS;
        if (KMAJ) {
            const int r = c / (GBK / 2), kc = (c % (GBK / 2)) * 2;
            int64_t gr = r0 + r;
            if (gr >= rmax) gr = rmax - 1;                 // clamp: duplicated rows are masked at the store
            const int64_t gk = k0 + kc;
            int bytes = 0;
            if (gk + 1 < kmax) bytes = 16; else if (gk < kmax) bytes = 8;
            const double* src = G + gr * ld + (gk < kmax ? gk : 0);
            cp_async_bytes(s + r * LDK_MAJ + kc, src, bytes, vec);
        } else {
            const int kk = c / (ROWS / 2), rc = (c % (ROWS / 2)) * 2;
            const int64_t gk = k0 + kk;
            const int64_t gr = r0 + rc;
            int bytes = 0;
            if (gk < kmax) { if (gr + 1 < rmax) bytes = 16; else if (gr < rmax) bytes = 8; }
            const double* src = G + (gk < kmax ? gk : 0) * ld + (gr < rmax ? gr : 0);
            cp_async_bytes(s + kk * (ROWS + 4) + rc, src, bytes, vec);
        }
    }
}

// acc += A_tile * B_tile^T over the whole k range, for the 128x64 tile at (row0, col0).
template <bool A_KMAJ, bool B_KMAJ>
__device__ __forceinline__ void gemm_tile_mainloop(const GemmParams& p, int64_t row0, int64_t col0, double* smem,
                                                   double (&acc)[4][4][2]) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    const int wm = warp / GWN, wn = warp % GWN;
    const int64_t nkt = (p.k + GBK - 1) / GBK;

    auto issue = [&](int64_t kt) {
        if (kt < nkt) {
            double* st = smem + (size_t)(kt % GSTAGES) * STAGE_DOUBLES;
            load_tile<A_KMAJ, GTM>(st, p.A, p.lda, row0, p.m, kt * GBK, p.k, p.vecA != 0, tid);
            load_tile<B_KMAJ, GTN>(st + TILE_A_DOUBLES, p.B, p.ldb, col0, p.n, kt * GBK, p.k, p.vecB != 0, tid);
        }
        cp_async_commit();
    };

#pragma unroll
    for (int s = 0; s < GSTAGES - 1; ++s) issue(s);

    for (int64_t kt = 0; kt < nkt; ++kt) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();                    // stage kt landed for everyone; stage (kt-1) is free for refill
        issue(kt + GSTAGES - 1);
        const double* a_s = smem + (size_t)(kt % GSTAGES) * STAGE_DOUBLES;
        const double* b_s = a_s + TILE_A_DOUBLES;
#pragma unroll
        for (int ks = 0; ks < GBK / 4; ++ks) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
                a[mi] = A_KMAJ ? a_s[(wm * 32 + mi * 8 + g) * LDK_MAJ + ks * 4 + q]
                               : a_s[(ks * 4 + q) * LDM_MAJ + wm * 32 + mi * 8 + g];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
                b[ni] = B_KMAJ ? b_s[(wn * 32 + ni * 8 + g) * LDK_MAJ + ks * 4 + q]
                               : b_s[(ks * 4 + q) * LDN_MAJ + wn * 32 + ni * 8 + g];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();
}

__device__ __forceinline__ int64_t gemm_global_row(const GemmParams& p, int64_t row) {
    if (p.cyc_nb > 0) {
        const int64_t b = row / p.cyc_nb;
        return p.row_base + (b * p.cyc_world + p.cyc_rank) * p.cyc_nb + (row - b * p.cyc_nb);
    }
    return p.row_base + row;
}

// C tile <- C tile + alpha*acc (or alpha*acc when beta0), honouring the bounds and the lower-triangle mask.
__device__ __forceinline__ void gemm_tile_epilogue(const GemmParams& p, int64_t row0, int64_t col0,
                                                   const double (&acc)[4][4][2]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int wm = warp / GWN, wn = warp % GWN;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const int64_t row = row0 + wm * 32 + mi * 8 + g;
        if (row >= p.m) continue;
        double* crow = p.C + row * p.ldc;
        int64_t lim = p.n;   // valid columns: col < lim
        if (p.lower) {
            const int64_t l2 = gemm_global_row(p, row) + 1 - p.col_base;
            lim = l2 < p.n ? l2 : p.n;
        }
        if (p.vecC && !p.beta0) {
            // issue the four 16-byte loads of this row first, then update and store
            double2 c[4];
            bool full[4];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int64_t col = col0 + wn * 32 + ni * 8 + 2 * q;
                full[ni] = col + 1 < lim;
                if (full[ni]) c[ni] = *reinterpret_cast<const double2*>(crow + col);
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int64_t col = col0 + wn * 32 + ni * 8 + 2 * q;
                if (full[ni]) {
                    c[ni].x = fma(p.alpha, acc[mi][ni][0], c[ni].x);
                    c[ni].y = fma(p.alpha, acc[mi][ni][1], c[ni].y);
                    *reinterpret_cast<double2*>(crow + col) = c[ni];
                } else if (col < lim) {
                    crow[col] = fma(p.alpha, acc[mi][ni][0], crow[col]);
                }
            }
        } else {
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int64_t col = col0 + wn * 32 + ni * 8 + 2 * q;
                if (col < lim) crow[col] = fma(p.alpha, acc[mi][ni][0], p.beta0 ? 0.0 : crow[col]);
                if (col + 1 < lim) crow[col + 1] = fma(p.alpha, acc[mi][ni][1], p.beta0 ? 0.0 : crow[col + 1]);
            }
        }
    }
}

template <bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__(GTHREADS, 2) gemm_kernel(const GemmParams p) {
    extern __shared__ __align__(16) double gemm_smem[];
    // Grouped rasterisation: consecutive CTAs walk DOWN a group of GEMM_GROUP_M tile rows before moving to the next tile
    // column, so the CTAs in flight share a few B tiles and GEMM_GROUP_M A tiles through L2.  With the plain row-major
    // order the k = 512 panel of the N = 32768 update (132 MB > L2) was re-streamed from HBM for every tile row:
    // 15.5 GB of DRAM reads against 4.3 GB algorithmic (profiles/r01_gemm_bench_shape_ncu.txt).
    const int64_t t = blockIdx.x;
    const int64_t per_group = (int64_t)GEMM_GROUP_M * p.tiles_n;
    const int64_t grp = t / per_group, rem = t - grp * per_group;
    const int64_t bj = rem / GEMM_GROUP_M, bi = grp * GEMM_GROUP_M + (rem - bj * GEMM_GROUP_M);
    if (bi >= p.tiles_m) return;
    const int64_t row0 = bi * GTM, col0 = bj * GTN;
    if (p.lower) {   // tile strictly above the diagonal (global coordinates of its last row / first column)
        const int64_t last = (row0 + GTM - 1 < p.m) ? row0 + GTM - 1 : p.m - 1;
        if (p.col_base + col0 > gemm_global_row(p, last)) return;
    }
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    gemm_tile_mainloop<A_KMAJ, B_KMAJ>(p, row0, col0, gemm_smem, acc);
    gemm_tile_epilogue(p, row0, col0, acc);
}

inline bool aligned16(const void* ptr, int64_t ld) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld % 2) == 0; }

// C[m,n] (+)= alpha * op(A) op(B).  a_kmaj: A given as [m,k] row-major (else [k,m]); b_kmaj: B given as [n,k] row-major
// (else [k,n]).
struct GemmMask {   // lower-triangle mask in global coordinates (all zero: local coordinates, C(0,0) on the diagonal)
    int64_t row_base = 0, col_base = 0;
    int cyc_nb = 0, cyc_world = 1, cyc_rank = 0;
};
int launch_gemm(bool a_kmaj, bool b_kmaj, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                const double* B, int64_t ldb, bool beta0, double* C, int64_t ldc, bool lower, cudaStream_t stream,
                const GemmMask& mask = GemmMask());

}  // namespace gabo
