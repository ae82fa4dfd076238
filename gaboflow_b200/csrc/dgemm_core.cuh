// FP64 tensor-pipe GEMM core for sm_100a:  C[m,n] -= op(A) * op(B),  accumulating on DMMA.8x8x4 (mma.sync m8n8k4).
//
// This is the trailing-update engine of the blocked Cholesky (SYRK-shaped when A == B), of the explicit-inverse TRSM
// panels, of the blocked triangular solves and of cov = K** - A^T A.  There is no tcgen05 FP64 kind, so the FP64
// "tensor pipe" on B200 is warp-level DMMA; measured peak 37.2 TFLOP/s (profiles/r01_fp64_probe.txt), shared with DFMA.
//
// CTA tile 128x128, 16 warps (4x4), 32x32 outputs per warp in registers (64 accumulator registers), BK = 16 per
// stage, 4-stage cp.async (LDGSTS) ring.  Operand tiles sit in shared memory with a leading dimension = 4 (mod 16)
// doubles, which makes every fragment load of the m8n8k4 layout bank-conflict free for both the K-major and the
// M/N-major operand orientations.
#pragma once
#include "gabo_common.cuh"

namespace gabo {

constexpr int GT = 128;          // CTA tile side
constexpr int GBK = 16;          // k per stage
constexpr int GSTAGES = 4;
constexpr int GTHREADS = 512;
constexpr int LDK_MAJ = GBK + 4;   // 20: tile stored [128][20]  (K-major operand: row = m or n index, k contiguous)
constexpr int LDN_MAJ = GT + 4;    // 132: tile stored [16][132] (M/N-major operand: row = k index, m or n contiguous)
constexpr int TILE_DOUBLES = (GT * LDK_MAJ > GBK * LDN_MAJ) ? GT * LDK_MAJ : GBK * LDN_MAJ;   // 2560
constexpr size_t GEMM_SMEM_BYTES = (size_t)GSTAGES * 2 * TILE_DOUBLES * sizeof(double);      // 163,840

struct GemmParams {
    const double* A; int64_t lda;
    const double* B; int64_t ldb;
    double* C; int64_t ldc;
    int64_t m, n, k;
    int lower;       // 1: only entries with col <= row are produced (tiles above the diagonal are skipped)
    int vecA, vecB, vecC;   // 16-byte access allowed (even leading dimension, 16 B aligned base)
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

// Loads one operand tile for k-block kt into shared memory.
//  KMAJ: global element (r, kk) at G[(r0 + r) * ld + k0 + kk]      -> smem [r][kk], ld 20
// !KMAJ: global element (kk, r) at G[(k0 + kk) * ld + r0 + r]      -> smem [kk][r], ld 132
template <bool KMAJ>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ G, int64_t ld, int64_t r0, int64_t rmax,
                                          int64_t k0, int64_t kmax, bool vec, int tid) {
#pragma unroll
    for (int it = 0; it < (GT * GBK / 2) / GTHREADS; ++it) {   // 1024 16-byte chunks / 512 threads
        const int c = tid + it * GTHREADS;
        if (KMAJ) {
            const int r = c >> 3, kc = (c & 7) * 2;
            int64_t gr = r0 + r;
            if (gr >= rmax) gr = rmax - 1;                 // clamp: duplicated rows are masked at the store
            const int64_t gk = k0 + kc;
            int bytes = 0;
            if (gk + 1 < kmax) bytes = 16; else if (gk < kmax) bytes = 8;
            const double* src = G + gr * ld + (gk < kmax ? gk : 0);
            cp_async_bytes(s + r * LDK_MAJ + kc, src, bytes, vec);
        } else {
            const int kk = c >> 6, rc = (c & 63) * 2;
            const int64_t gk = k0 + kk;
            const int64_t gr = r0 + rc;
            int bytes = 0;
            if (gk < kmax) { if (gr + 1 < rmax) bytes = 16; else if (gr < rmax) bytes = 8; }
            const double* src = G + (gk < kmax ? gk : 0) * ld + (gr < rmax ? gr : 0);
            cp_async_bytes(s + kk * LDN_MAJ + rc, src, bytes, vec);
        }
    }
}

// acc[mi][ni][0..1] += A_tile * B_tile^T over the whole k range, for the 128x128 tile at (row0, col0).
template <bool A_KMAJ, bool B_KMAJ>
__device__ __forceinline__ void gemm_tile_mainloop(const GemmParams& p, int64_t row0, int64_t col0, double* smem,
                                                   double (&acc)[4][4][2]) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int64_t nkt = (p.k + GBK - 1) / GBK;

    auto sA = [&](int s) { return smem + (size_t)s * 2 * TILE_DOUBLES; };
    auto sB = [&](int s) { return smem + (size_t)s * 2 * TILE_DOUBLES + TILE_DOUBLES; };
    auto issue = [&](int64_t kt) {
        if (kt < nkt) {
            const int s = (int)(kt % GSTAGES);
            load_tile<A_KMAJ>(sA(s), p.A, p.lda, row0, p.m, kt * GBK, p.k, p.vecA != 0, tid);
            load_tile<B_KMAJ>(sB(s), p.B, p.ldb, col0, p.n, kt * GBK, p.k, p.vecB != 0, tid);
        }
        cp_async_commit();
    };

#pragma unroll
    for (int s = 0; s < GSTAGES - 1; ++s) issue(s);

    for (int64_t kt = 0; kt < nkt; ++kt) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();                    // stage kt landed for everyone; stage (kt-1) is free for refill
        issue(kt + GSTAGES - 1);
        const int s = (int)(kt % GSTAGES);
        const double* a_s = sA(s);
        const double* b_s = sB(s);
#pragma unroll
        for (int ks = 0; ks < GBK / 4; ++ks) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
                a[mi] = A_KMAJ ? a_s[(wm * 32 + mi * 8 + g) * LDK_MAJ + ks * 4 + q]
                               : a_s[(ks * 4 + q) * LDN_MAJ + wm * 32 + mi * 8 + g];
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
    __syncthreads();
}

// C tile <- C tile + alpha*acc (or alpha*acc when beta0), honouring the bounds and the lower-triangle mask.
__device__ __forceinline__ void gemm_tile_epilogue(const GemmParams& p, int64_t row0, int64_t col0,
                                                   const double (&acc)[4][4][2]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const int64_t row = row0 + wm * 32 + mi * 8 + g;
        if (row >= p.m) continue;
        double* crow = p.C + row * p.ldc;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int64_t col = col0 + wn * 32 + ni * 8 + 2 * q;
            const int64_t lim = p.lower ? (row + 1 < p.n ? row + 1 : p.n) : p.n;   // valid columns: col < lim
            if (col + 1 < lim && p.vecC) {
                double2 c = make_double2(0.0, 0.0);
                if (!p.beta0) c = *reinterpret_cast<const double2*>(crow + col);
                c.x = fma(p.alpha, acc[mi][ni][0], c.x);
                c.y = fma(p.alpha, acc[mi][ni][1], c.y);
                *reinterpret_cast<double2*>(crow + col) = c;
            } else {
                if (col < lim) crow[col] = fma(p.alpha, acc[mi][ni][0], p.beta0 ? 0.0 : crow[col]);
                if (col + 1 < lim) crow[col + 1] = fma(p.alpha, acc[mi][ni][1], p.beta0 ? 0.0 : crow[col + 1]);
            }
        }
    }
}

template <bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__(GTHREADS, 1) gemm_kernel(const GemmParams p) {
    extern __shared__ __align__(16) double gemm_smem[];
    const int64_t bi = blockIdx.y, bj = blockIdx.x;
    const int64_t row0 = bi * GT, col0 = bj * GT;
    if (p.lower && col0 > row0 + GT - 1) return;   // tile strictly above the diagonal
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
int launch_gemm(bool a_kmaj, bool b_kmaj, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                const double* B, int64_t ldb, bool beta0, double* C, int64_t ldc, bool lower, cudaStream_t stream);

}  // namespace gabo
