// Shared device helpers for the sm_100a kernels: DMMA.8x8x4 wrapper, mbarrier + TMA bulk-copy PTX,
// status-code plumbing.  Compiled only for compute_100a.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gabo_b200.h"

#define GABO_CUDA_TRY(expr)                                                    \
    do {                                                                       \
        cudaError_t e_ = (expr);                                               \
        if (e_ != cudaSuccess) return GABO_ERR_CUDA - (int)e_;                 \
    } while (0)

// every kernel launch of the library passes through here: error check + diagnostic launch counter
#define GABO_LAUNCH_CHECK()                                \
    do {                                                   \
        ::gabo::count_launch();                            \
        GABO_CUDA_TRY(cudaPeekAtLastError());              \
    } while (0)

namespace gabo {

constexpr int kNumSMsB200 = 148;

void count_launch();   // bumps the process-wide diagnostic counter read by gabo_launch_count() (gabo_api.cu)
int sm_count();  // cached cudaDevAttrMultiProcessorCount of the current device (gabo_api.cu)

// D(8x8) += A(8x4, row) * B(4x8, col), FP64 tensor pipe.  Fragment ownership (lane = 4*g + q):
//   a = A[g][q], b = B[q][g] (i.e. element k=q of column g), c0/c1 = C[g][2q], C[g][2q+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared (1-D, contiguous), completion on an mbarrier.  SASS: UBLKCP.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void st_cs_v2(double* p, double a, double b) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};\n" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_cs(double* p, double a) {
    asm volatile("st.global.cs.f64 [%0], %1;\n" ::"l"(p), "d"(a) : "memory");
}

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace gabo
