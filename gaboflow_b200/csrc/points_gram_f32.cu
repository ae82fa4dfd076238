// FP32 mode of the point-set Gram kernel (north_star: Gram entries within 1e-5 relative in FP32 mode): entry point and the
// SIMT fallback.  Same contract as gabo_points_gram_f64 (reference sphere_utils_tf.py:11-60 + kernels_sphere_tf.py:59-88,
// 179-204 and the pair stage of kernels_spd_tf.py:510-533,618-641), with float inputs and outputs.  Point dimensions up to 56
// (every configuration of BASELINE.json) run on the tensor cores (points_gram_f32_tc.cu); larger ones, or GABO_F32_SIMT=1, on
// the FFMA kernel below (128x128 tile, 8x8 outputs per thread, libm acosf / expf).
#include <cstdlib>
#include "gabo_common.cuh"

namespace gabo {
// tensor-core path (points_gram_f32_tc.cu): tcgen05.mma kind::tf32, three-term split, accumulators in TMEM
bool points_gram_f32_tc_supported(int dim);
int points_gram_f32_tc(int kind, const float* X, int64_t n1, int64_t ldx, const float* X2, int64_t n2, int64_t ldx2, int dim,
                       float beta, float variance, float* K, int64_t ldk, int64_t row0, int64_t row1, int uplo, bool symmetric,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream);
namespace {

constexpr int FT = 128;     // tile side
constexpr int FBK = 16;     // k per shared-memory step
constexpr int FLD = FT + 4; // [k][132] floats: 16-byte aligned rows, conflict-free float4 reads

struct GramParamsF32 {
    const float* X; int64_t ldx;
    const float* X2; int64_t ldx2;
    float* K; int64_t ldk;
    int64_t n1, n2, row0, row1;
    int dim, kind, uplo, symmetric, vec_ok;
    float beta, variance;
};

__device__ __forceinline__ float gram_entry_f32(int kind, float acc, float na, float nb, float beta, float variance) {
    if (kind == GABO_SQDIST_GAUSSIAN) {
        float d2 = fmaxf(fmaf(-2.0f, acc, na + nb), 0.0f);
        return variance * expf(-beta * d2);
    }
    const float t = fminf(fmaxf(acc, -1.0f), 1.0f);   // sphere_utils_tf.py:59
    const float d = acosf(t);
    if (kind == GABO_SPHERE_GAUSSIAN) return variance * expf(-beta * d * d);
    return variance * expf(-beta * d);
}

__global__ void __launch_bounds__(256) points_gram_f32_kernel(const GramParamsF32 p) {
    __shared__ __align__(16) float sA[FBK][FLD];
    __shared__ __align__(16) float sB[FBK][FLD];
    __shared__ float sNa[FT], sNb[FT];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t i0 = p.row0 + (int64_t)blockIdx.y * FT;
    const int64_t j0 = (int64_t)blockIdx.x * FT;
    if (p.uplo != GABO_FULL && j0 > i0 + FT - 1) return;
    const bool mirror = (p.uplo == GABO_SYMM_MIRROR) && (j0 != i0);

    float acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.0f;
    float na_part = 0.0f, nb_part = 0.0f;   // thread tid < 128 accumulates |x_r|^2 of row r = tid of each operand

    for (int k0 = 0; k0 < p.dim; k0 += FBK) {
        // 128 rows x 16 k per operand; thread loads 8 elements of each: row = e / 16, k = e % 16 (k fastest: coalesced)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int e = tid + it * 256;
            const int r = e >> 4, kk = e & 15;
            const int k = k0 + kk;
            const int64_t gi = i0 + r, gj = j0 + r;
            sA[kk][r] = (gi < p.row1 && k < p.dim) ? p.X[gi * p.ldx + k] : 0.0f;
            sB[kk][r] = (gj < p.n2 && k < p.dim) ? p.X2[gj * p.ldx2 + k] : 0.0f;
        }
        __syncthreads();
        if (p.kind == GABO_SQDIST_GAUSSIAN && tid < FT) {
#pragma unroll
            for (int kk = 0; kk < FBK; ++kk) {
                na_part = fmaf(sA[kk][tid], sA[kk][tid], na_part);
                nb_part = fmaf(sB[kk][tid], sB[kk][tid], nb_part);
            }
        }
#pragma unroll
        for (int kk = 0; kk < FBK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&sA[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&sA[kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sB[kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int x = 0; x < 8; ++x)
#pragma unroll
                for (int y = 0; y < 8; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
        }
        __syncthreads();
    }
    if (p.kind == GABO_SQDIST_GAUSSIAN) {
        if (tid < FT) { sNa[tid] = na_part; sNb[tid] = nb_part; }
        __syncthreads();
    }
    const bool diag_tile = p.symmetric && (j0 == i0);
#pragma unroll
    for (int x = 0; x < 8; ++x) {
        const int rl = (x < 4) ? ty * 4 + x : 64 + ty * 4 + (x - 4);
        const int64_t gi = i0 + rl;
        if (gi >= p.row1) continue;
        float* krow = p.K + (gi - p.row0) * p.ldk;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int cl = h * 64 + tx * 4;
            const int64_t gj = j0 + cl;
            float v[4];
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const float na = (p.kind == GABO_SQDIST_GAUSSIAN) ? sNa[rl] : 0.0f;
                const float nb = (p.kind == GABO_SQDIST_GAUSSIAN) ? sNb[cl + y] : 0.0f;
                v[y] = gram_entry_f32(p.kind, acc[x][h * 4 + y], na, nb, p.beta, p.variance);
                if (diag_tile && gj + y == gi) v[y] = p.variance;
            }
            if (gj + 3 < p.n2 && p.vec_ok) {
                __stcs(reinterpret_cast<float4*>(krow + gj), make_float4(v[0], v[1], v[2], v[3]));
            } else {
#pragma unroll
                for (int y = 0; y < 4; ++y)
                    if (gj + y < p.n2) krow[gj + y] = v[y];
            }
            if (mirror) {
#pragma unroll
                for (int y = 0; y < 4; ++y)
                    if (gj + y < p.n2) p.K[(gj + y) * p.ldk + gi] = v[y];
            }
        }
    }
}

}  // namespace
}  // namespace gabo

extern "C" int gabo_points_gram_f32(int kind, const float* X, int64_t n1, int64_t ldx, const float* X2, int64_t n2,
                                    int64_t ldx2, int dim, float beta, float variance, float* K, int64_t ldk,
                                    int64_t row0, int64_t row1, int uplo, void* workspace, size_t workspace_bytes,
                                    gabo_stream_t stream) {
    using namespace gabo;
    if (kind < GABO_SPHERE_GAUSSIAN || kind > GABO_SQDIST_GAUSSIAN) return -1;
    if (n1 < 0) return -3;
    if (n1 == 0) return 0;
    if (X == nullptr) return -2;
    if (ldx < dim) return -4;
    const bool symmetric = (X2 == nullptr);
    if (symmetric) { X2 = X; n2 = n1; ldx2 = ldx; }
    if (n2 < 0) return -6;
    if (ldx2 < dim) return -7;
    if (dim < 1) return -8;
    if (K == nullptr) return -11;
    if (ldk < n2) return -12;
    if (row0 < 0 || row0 % FT != 0 || row0 > n1) return -13;
    if (row1 < row0 || row1 > n1) return -14;
    if (uplo < GABO_FULL || uplo > GABO_SYMM_MIRROR) return -15;
    if (uplo != GABO_FULL && !symmetric) return -15;
    if (uplo == GABO_SYMM_MIRROR && !(row0 == 0 && row1 == n1)) return -15;
    if (row1 == row0 || n2 == 0) return 0;
    static const bool force_simt = (getenv("GABO_F32_SIMT") != nullptr);
    if (points_gram_f32_tc_supported(dim) && !force_simt)
        return points_gram_f32_tc(kind, X, n1, ldx, symmetric ? nullptr : X2, n2, ldx2, dim, beta, variance, K, ldk, row0, row1, uplo,
                                  symmetric, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
    GramParamsF32 p;
    p.X = X; p.ldx = ldx; p.X2 = X2; p.ldx2 = ldx2; p.K = K; p.ldk = ldk; p.n1 = n1; p.n2 = n2; p.row0 = row0; p.row1 = row1;
    p.dim = dim; p.kind = kind; p.uplo = uplo; p.symmetric = symmetric ? 1 : 0;
    p.vec_ok = ((ldk % 4) == 0 && (reinterpret_cast<uintptr_t>(K) & 15) == 0) ? 1 : 0;
    p.beta = beta; p.variance = variance;
    dim3 grid((unsigned)ceil_div64(n2, FT), (unsigned)ceil_div64(row1 - row0, FT));
    points_gram_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    GABO_LAUNCH_CHECK();
    return 0;
}
