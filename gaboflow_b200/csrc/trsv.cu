// Triangular solve with ONE right-hand side as a single persistent kernel: L x = b or L^T x = b for the row-major lower
// Cholesky factor (tf.matrix_triangular_solve on y - m in GPflow 0.5's GPR.build_likelihood, and on every posterior mean).
//
// Round 1 ran this as a chain of 2 small launches per 128-block (512 launches, 3.9 ms at N = 32768: 17 % of the HBM time of
// reading L once).  Here every 128-row block of x is owned by one resident CTA; the CTA streams its block row of L (forward)
// or block column (transposed) from HBM, consuming each earlier block x_b as soon as the CTA that owns it has published it
// (a monotone progress counter in global memory, ld.acquire / st.release, no grid-wide barrier, no host round trip), then
// applies the pre-inverted 128 x 128 diagonal block and publishes its own x_r.  The early part of the solve is bound by HBM
// (all CTAs read one block column of L per step), the late part by the dependency chain (one 128 x 128 product + one inverse
// application + one L2 round trip per block).
#include "gabo_common.cuh"

namespace gabo {
namespace {

constexpr int TB = 128;          // block size (= NB_IN of linalg.cu: the diagonal-block inverses are 128 x 128)
constexpr int TRSV_THREADS = 256;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// sums over the 32 lanes of 16 per-lane values; on return lane l holds the total of value (l >> 1)
__device__ __forceinline__ double warp_reduce16(double (&a)[16], int lane) {
    double b8[8], b4[4], b2[2];
    {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double send = up ? a[i] : a[i + 8], keep = up ? a[i + 8] : a[i];
            b8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double send = up ? b8[i] : b8[i + 4], keep = up ? b8[i + 4] : b8[i];
            b4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double send = up ? b4[i] : b4[i + 2], keep = up ? b4[i + 2] : b4[i];
            b2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    const bool up = lane & 2;
    const double send = up ? b2[0] : b2[1], keep = up ? b2[1] : b2[0];
    double v = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// wait until at least `need` blocks are complete (lane 0 polls, the warp follows)
__device__ __forceinline__ void wait_done(const unsigned* done, unsigned need, unsigned& seen, int lane) {
    if (seen >= need) return;
    unsigned v = 0;
    if (lane == 0) {
        v = ld_acquire_u32(done);
        while (v < need) {
            __nanosleep(20);
            v = ld_acquire_u32(done);
        }
    }
    seen = __shfl_sync(0xffffffffu, v, 0);
}

template <bool TRANS>
__global__ void __launch_bounds__(TRSV_THREADS, 2) trsv_persistent_kernel(const double* __restrict__ L, int64_t ldl, int64_t n,
                                                                          const double* __restrict__ Winv, double* x,
                                                                          int64_t ldb, unsigned* done) {
    __shared__ double sy[TB];
    __shared__ double spart[8][TB];     // TRANS: per-warp partial sums over rows
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t nblk = (n + TB - 1) / TB;
    unsigned seen = 0;
    for (int64_t ord = blockIdx.x; ord < nblk; ord += gridDim.x) {
        // the ord-th block to complete: forward r = ord, transposed r = nblk - 1 - ord
        const int64_t r = TRANS ? (nblk - 1 - ord) : ord;
        const int64_t r0 = r * TB;
        const double* W = Winv + r * (int64_t)TB * TB;
        // the diagonal-block inverse is needed at the very end of the dependency chain: pull it into L2 now
        for (int e = tid; e < TB * TB / 16; e += TRSV_THREADS) prefetch_l2(W + (int64_t)e * 16);
        if (!TRANS) {
            // y_r = b_r - sum_{b < r} L[r, b] x_b : warp w owns rows 16 w .. 16 w + 15, lane l columns l + 32 k of every block
            double acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.0;
            const double* Lr = L + (r0 + 16 * warp) * ldl + lane;
            for (int64_t b = 0; b < r; ++b) {
                if (b + 1 < r) {   // next block of this row strip towards L2 while this one is consumed
                    const double* nx = L + (r0 + 16 * warp + (lane >> 1)) * ldl + (b + 1) * TB + (lane & 1) * 64;
                    if (r0 + 16 * warp + (lane >> 1) < n) { prefetch_l2(nx); prefetch_l2(nx + 16); prefetch_l2(nx + 32); prefetch_l2(nx + 48); }
                }
                wait_done(done, (unsigned)(b + 1), seen, lane);
                double xb[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) xb[k] = __ldcg(x + (b * TB + lane + 32 * k) * ldb);
                const double* Lb = Lr + b * TB;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (r0 + 16 * warp + i < n) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[i] = fma(__ldcs(Lb + i * ldl + 32 * k), xb[k], acc[i]);
                    }
                }
            }
            const double s = warp_reduce16(acc, lane);
            const int64_t row = r0 + 16 * warp + (lane >> 1);
            if ((lane & 1) == 0) sy[16 * warp + (lane >> 1)] = (row < n) ? x[row * ldb] - s : 0.0;
            __syncthreads();
            // x_r = Winv_r y_r (Winv lower triangular, identity-padded past n)
            double a2[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const double* wr = W + (16 * warp + i) * TB + lane;
                a2[i] = 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) a2[i] = fma(__ldcg(wr + 32 * k), sy[lane + 32 * k], a2[i]);
            }
            const double xr = warp_reduce16(a2, lane);
            if ((lane & 1) == 0 && row < n) x[row * ldb] = xr;
        } else {
            // y_r = b_r - sum_{b > r} L[b, r]^T x_b : warp w owns rows 16 w .. of every block, lane l columns r0 + l + 32 k
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int64_t b = nblk - 1; b > r; --b) {
                wait_done(done, (unsigned)(nblk - b), seen, lane);
                const int64_t row0 = b * TB + 16 * warp;
                double xb[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) xb[i] = (row0 + i < n) ? __ldcg(x + (row0 + i) * ldb) : 0.0;
                const double* Lb = L + row0 * ldl + r0 + lane;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (row0 + i < n) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[k] = fma(__ldcs(Lb + i * ldl + 32 * k), xb[i], acc[k]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) spart[warp][lane + 32 * k] = acc[k];
            __syncthreads();
            if (tid < TB) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += spart[w][tid];
                sy[tid] = (r0 + tid < n) ? x[(r0 + tid) * ldb] - s : 0.0;
            }
            __syncthreads();
            // x_r = Winv_r^T y_r : x_r[c] = sum_{row >= c} Winv[row][c] y[row]
            double a2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int row = 16 * warp + i;
                const double yv = sy[row];
#pragma unroll
                for (int k = 0; k < 4; ++k) a2[k] = fma(__ldcg(W + row * TB + lane + 32 * k), yv, a2[k]);
            }
            __syncthreads();     // everyone has read spart / sy of the first reduction
#pragma unroll
            for (int k = 0; k < 4; ++k) spart[warp][lane + 32 * k] = a2[k];
            __syncthreads();
            if (tid < TB && r0 + tid < n) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += spart[w][tid];
                x[(r0 + tid) * ldb] = s;
            }
        }
        __syncthreads();         // all rows of x_r are written
        if (tid == 0) {
            __threadfence();
            st_release_u32(done, (unsigned)(ord + 1));
        }
        // sy / spart are rewritten only after the next block's loop, which ends in __syncthreads-separated phases
    }
}

}  // namespace

size_t trsv_persistent_workspace_bytes() { return 256; }

// x (stride ldb, in place) <- L^-1 x or L^-T x.  Winv: inverses of the 128 x 128 diagonal blocks of L.  `scratch`: 256 B of
// device memory (the progress counter).  Cooperative launch: every CTA must be resident, they wait on each other.
int trsv_persistent(bool trans, int64_t n, const double* L, int64_t ldl, const double* Winv, double* x, int64_t ldb,
                    void* scratch, cudaStream_t stream) {
    if (n <= 0) return 0;
    unsigned* done = static_cast<unsigned*>(scratch);
    GABO_CUDA_TRY(cudaMemsetAsync(done, 0, sizeof(unsigned), stream));
    static int max_ctas = 0;     // resident CTAs of this kernel on the current device (2 per SM by launch bounds)
    if (max_ctas == 0) {
        int per_sm = 0;
        GABO_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trsv_persistent_kernel<false>, TRSV_THREADS, 0));
        int per_sm_t = 0;
        GABO_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_t, trsv_persistent_kernel<true>, TRSV_THREADS, 0));
        if (per_sm_t < per_sm) per_sm = per_sm_t;
        if (per_sm < 1) return GABO_ERR_CUDA - (int)cudaErrorLaunchOutOfResources;
        max_ctas = per_sm * sm_count();
    }
    const int64_t nblk = (n + TB - 1) / TB;
    const unsigned grid = (unsigned)(nblk < max_ctas ? nblk : max_ctas);
    void* args[] = {(void*)&L, (void*)&ldl, (void*)&n, (void*)&Winv, (void*)&x, (void*)&ldb, (void*)&done};
    const void* fn = trans ? (const void*)trsv_persistent_kernel<true> : (const void*)trsv_persistent_kernel<false>;
    GABO_CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(TRSV_THREADS), args, 0, stream));
    count_launch();
    return 0;
}

}  // namespace gabo
