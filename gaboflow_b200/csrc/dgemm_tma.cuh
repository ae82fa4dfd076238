// TMA-fed variant of the FP64 DMMA GEMM core for the hot "NT" case (both operands K-major): C[m,n] (+)= alpha A B^T.
//
// Same tiling as dgemm_core.cuh (128x64 CTA tile, 8 consumer warps x 32x32, 2 CTAs/SM, grouped rasterisation, lower
// mask in global coordinates), but the operand pipeline is Blackwell/Hopper style instead of cp.async + CTA barriers:
//   * a dedicated producer warp issues 2-D TMA tensor loads (cp.async.bulk.tensor, SASS UTMALDG) of boxes {4 k, rows}:
//     a box lands as a dense [rows][4] block, so the m8n8k4 fragment of k-step ks reads 8 rows x 32 B = 256 contiguous
//     bytes of box ks -- bank-conflict free with NO padding and no swizzle; out-of-range rows / k are zero-filled by TMA;
//   * full/empty mbarriers per stage: consumer warps never rendezvous with each other, they only wait for the bytes of
//     their stage (complete_tx) and release it with one arrive per warp;
//   * no LDGSTS address arithmetic in the consumer warps (ncu on the cp.async version: 85 % DMMA-active, the rest
//     wait / barrier stalls).
// Falls back (return code GABO_TMA_NA) when the operands are not 16-byte aligned with even leading dimensions.
#pragma once
#include <cuda.h>
#include "dgemm_core.cuh"

namespace gabo {

constexpr int GABO_TMA_NA = -999;
constexpr int TBK = 16;                 // k per stage
constexpr int TSTAGES = 3;
constexpr int TBOX = TBK / 4;           // boxes of 4 k per operand and stage
constexpr int T_CONSUMERS = 8;          // consumer warps
constexpr int T_THREADS = (T_CONSUMERS + 1) * 32;
constexpr int T_A_DOUBLES = GTM * TBK;  // 2048
constexpr int T_B_DOUBLES = GTN * TBK;  // 1024
constexpr int T_STAGE_DOUBLES = T_A_DOUBLES + T_B_DOUBLES;
constexpr size_t T_SMEM_BYTES = (size_t)TSTAGES * T_STAGE_DOUBLES * sizeof(double) + 128;   // 73,856: small enough that potf2 (133 KB) fits beside one resident CTA

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(T_THREADS, 2)
gemm_nt_tma_kernel(const GemmParams p, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB) {
    extern __shared__ __align__(128) unsigned char tma_smem_raw[];
    double* smem = reinterpret_cast<double*>(tma_smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tma_smem_raw + (size_t)TSTAGES * T_STAGE_DOUBLES * sizeof(double));
    uint64_t* empty_bar = full_bar + TSTAGES;

    const int64_t t = blockIdx.x;
    const int64_t per_group = (int64_t)GEMM_GROUP_M * p.tiles_n;
    const int64_t grp = t / per_group, rem = t - grp * per_group;
    const int64_t bj = rem / GEMM_GROUP_M, bi = grp * GEMM_GROUP_M + (rem - bj * GEMM_GROUP_M);
    if (bi >= p.tiles_m) return;
    const int64_t row0 = bi * GTM, col0 = bj * GTN;
    if (p.lower) {
        const int64_t last = (row0 + GTM - 1 < p.m) ? row0 + GTM - 1 : p.m - 1;
        if (p.col_base + col0 > gemm_global_row(p, last)) return;
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TSTAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], T_CONSUMERS); }
        fence_mbar_init();
    }
    __syncthreads();
    const int nkt = (int)((p.k + TBK - 1) / TBK);

    if (warp == T_CONSUMERS) {
        // ===== producer warp: one lane issues the TMA loads =====
        if (lane == 0) {
            for (int kt = 0; kt < nkt; ++kt) {
                const int s = kt % TSTAGES;
                mbar_wait(&empty_bar[s], (uint32_t)(((kt / TSTAGES) & 1) ^ 1));   // slot free (passes at once on the first lap)
                double* sA = smem + (size_t)s * T_STAGE_DOUBLES;
                double* sB = sA + T_A_DOUBLES;
                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(T_STAGE_DOUBLES * sizeof(double)));
#pragma unroll
                for (int bx = 0; bx < TBOX; ++bx) {
                    tma_load_2d(sA + bx * GTM * 4, &mapA, kt * TBK + bx * 4, (int)row0, &full_bar[s]);
                    tma_load_2d(sB + bx * GTN * 4, &mapB, kt * TBK + bx * 4, (int)col0, &full_bar[s]);
                }
            }
        }
        return;
    }

    // ===== consumer warps =====
    const int g = lane >> 2, q = lane & 3;
    const int wm = warp / GWN, wn = warp % GWN;
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    for (int kt = 0; kt < nkt; ++kt) {
        const int s = kt % TSTAGES;
        mbar_wait(&full_bar[s], (uint32_t)((kt / TSTAGES) & 1));
        const double* a_s = smem + (size_t)s * T_STAGE_DOUBLES + (wm * 32 + g) * 4 + q;
        const double* b_s = smem + (size_t)s * T_STAGE_DOUBLES + T_A_DOUBLES + (wn * 32 + g) * 4 + q;
#pragma unroll
        for (int ks = 0; ks < TBOX; ++ks) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = a_s[ks * GTM * 4 + mi * 32];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = b_s[ks * GTN * 4 + ni * 32];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);   // this warp is done reading stage s
    }
    gemm_tile_epilogue(p, row0, col0, acc);
}

// Host side: tensor maps + launch.  Returns GABO_TMA_NA when TMA cannot be used for these operands.
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

inline bool make_kmajor_map(CUtensorMap* map, const double* base, int64_t rows, int64_t k, int64_t ld, int box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (enc == nullptr) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(double)};
    const cuuint32_t box[2] = {4u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int launch_gemm_nt_tma(const GemmParams& p, dim3 grid, cudaStream_t stream) {
    if (!p.vecA || !p.vecB) return GABO_TMA_NA;                       // 16-byte aligned base, even leading dimension
    if (p.k < 1 || p.m > 2000000000LL || p.n > 2000000000LL) return GABO_TMA_NA;
    CUtensorMap mapA, mapB;
    if (!make_kmajor_map(&mapA, p.A, p.m, p.k, p.lda, GTM)) return GABO_TMA_NA;
    if (!make_kmajor_map(&mapB, p.B, p.n, p.k, p.ldb, GTN)) return GABO_TMA_NA;
    GABO_CUDA_TRY(cudaFuncSetAttribute(gemm_nt_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM_BYTES));
    gemm_nt_tma_kernel<<<grid, T_THREADS, T_SMEM_BYTES, stream>>>(p, mapA, mapB);
    GABO_LAUNCH_CHECK();
    return 0;
}

}  // namespace gabo
