// Diagnostic entry point: evaluates the lean device math of gabo_math.cuh (the acos / exp / log that the Gram epilogues
// inline) element-wise, through exactly the batched code paths the kernels use, so that tests can measure their error in
// ulps against mpmath (tests/test_gpu_gram.py::test_fast_math_accuracy) and pin the exp underflow behaviour.
#include "gabo_common.cuh"
#include "gabo_math.cuh"

namespace gabo {
namespace {

__global__ void math_probe_kernel(int which, const double* __restrict__ x, int64_t n, double scale, double* __restrict__ out) {
    __shared__ double tab[32];
    if (threadIdx.x < 32) tab[threadIdx.x] = kExp2Tab[threadIdx.x] * scale;
    __syncthreads();
    // whole warps walk the array in batches of 4 so that the warp votes inside the batched helpers are uniform
    const int64_t nb = (n + 3) / 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nb_pad = ((nb + 31) / 32) * 32;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb_pad; b += stride) {
        double in[4], res[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = b * 4 + k;
            in[k] = (i < n) ? x[i] : (which == 2 ? 1.0 : 0.0);
        }
        if (which == 0) {
            acos_clamped_fast_n<4>(in, res);
        } else if (which == 1) {
            exp_neg_scaled_fast_n<4>(in, res, tab, scale, exp_slow_threshold_hi(scale));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) res[k] = log_pos_normal(in[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = b * 4 + k;
            if (i < n) out[i] = res[k];
        }
    }
}

}  // namespace
}  // namespace gabo

extern "C" int gabo_math_probe_f64(int which, const double* x, int64_t n, double scale, double* out, gabo_stream_t stream) {
    if (which < 0 || which > 2) return -1;
    if (n < 0) return -3;
    if (n == 0) return 0;
    if (x == nullptr) return -2;
    if (out == nullptr) return -5;
    const int threads = 128;
    int64_t blocks = gabo::ceil_div64(gabo::ceil_div64(n, 4), threads);
    if (blocks > 4096) blocks = 4096;
    gabo::math_probe_kernel<<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(which, x, n, scale, out);
    GABO_LAUNCH_CHECK();
    return 0;
}
