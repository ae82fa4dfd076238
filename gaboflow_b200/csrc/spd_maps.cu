// Batched exponential / logarithm maps of the SPD manifold with the affine-invariant metric (sm_100a).
//
// Replaces  expmap_tf (reference SPD_utils_tf.py:142-167):  X = S expm(S^-1 U)   per matrix: solve + complex64 expm + matmul
//           logmap_tf (SPD_utils_tf.py:170-195):            U = S logm(S^-1 X)   per matrix: solve + complex64 logm + matmul
// and their NumPy / Mandel twins (SPD_utils.py:104-177).  With S = L L^T and W = L^-1 both maps are a symmetric eigenproblem:
//     S f(S^-1 Y) = L f(W Y W^T) L^T = (L V) f(Lambda) (L V)^T,     W Y W^T = V Lambda V^T,
// f = exp for Y = U (symmetric), f = log for Y = X (SPD), so everything stays real FP64 (the reference's complex64 round trip
// is only ~1e-7 accurate) and the result is symmetric by construction.  One thread per point, all d x d work in registers
// (Cholesky of the base, congruence, cyclic Jacobi with vectors).  Inputs / outputs are either [n, d*d] row-major matrices
// or Mandel vectors [n, m] (SPD_utils_tf.py:11-92 layout); the base point is one matrix for all points (base_stride == 0, what
// the reference supports) or one per point.
#include "spd_small.cuh"

namespace gabo {
namespace {

template <int D>
__device__ __forceinline__ void load_sym(const double* __restrict__ p, bool mandel, double (&A)[D][D]) {
    const double inv_sqrt2 = 0.70710678118654752440;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = r; c < D; ++c) {
            double v;
            if (mandel) {
                v = p[mandel_slot<D>(r, c)];
                if (c != r) v *= inv_sqrt2;
            } else {
                v = 0.5 * (p[r * D + c] + p[c * D + r]);
            }
            A[r][c] = v;
            A[c][r] = v;
        }
}

template <int D, bool IS_LOG>
__global__ void spd_map_kernel(const double* __restrict__ Y, int64_t ldy, const double* __restrict__ base, int64_t base_stride,
                               int64_t n, int mandel, double* __restrict__ out, int64_t ldo, int32_t* __restrict__ info) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double sqrt2 = 1.41421356237309504880;
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    double S[D][D], L[D][D], W[D][D];
    load_sym<D>(base + i * base_stride, mandel != 0, S);
    int bad = chol_lower<D>(S, L);
    tri_inverse_lower<D>(L, W);
    double A[D][D];
    load_sym<D>(Y + i * ldy, mandel != 0, A);
    // A <- W A W^T (W lower triangular)
    double T[D][D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k <= r; ++k) v = fma(W[r][k], A[k][c], v);
            T[r][c] = v;
        }
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = r; c < D; ++c) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k <= c; ++k) v = fma(T[r][k], W[c][k], v);
            A[r][c] = v;
            A[c][r] = v;
        }
    double V[D][D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) V[r][c] = (r == c) ? 1.0 : 0.0;
    jacobi_eig<D, true>(A, V);
    double f[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        if (IS_LOG) {
            if (!(A[k][k] > 0.0) && bad == 0) bad = D + 1 + k;     // X not positive definite: NaN, as tf.log would give
            f[k] = log(A[k][k]);
        } else {
            f[k] = exp(A[k][k]);
        }
    }
    // B = L V;  out = B diag(f) B^T
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k <= r; ++k) v = fma(L[r][k], V[k][c], v);
            T[r][c] = v;
        }
    double* o = out + i * ldo;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = r; c < D; ++c) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) v = fma(T[r][k] * f[k], T[c][k], v);
            if (bad) v = nanv;
            if (mandel) {
                o[mandel_slot<D>(r, c)] = (c != r) ? v * sqrt2 : v;
            } else {
                o[r * D + c] = v;
                o[c * D + r] = v;
            }
        }
    if (info) info[i] = bad;
}

template <int D>
__global__ void sym_to_mandel_kernel(const double* __restrict__ S, int64_t lds, int64_t n, double* __restrict__ out, int64_t ldo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double sqrt2 = 1.41421356237309504880;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = r; c < D; ++c) {
            const double v = S[i * lds + r * D + c];
            out[i * ldo + mandel_slot<D>(r, c)] = (c != r) ? v * sqrt2 : v;     // SPD_utils_tf.py:87-89
        }
}

template <bool IS_LOG>
int launch_spd_map(int d, const double* Y, int64_t ldy, const double* base, int64_t base_stride, int64_t n, int mandel,
                   double* out, int64_t ldo, int32_t* info, cudaStream_t s) {
    const int threads = 64;
    const unsigned grid = (unsigned)ceil_div64(n, threads);
#define GABO_MAP_CASE(DV)                                                                                       \
    case DV:                                                                                                    \
        spd_map_kernel<DV, IS_LOG><<<grid, threads, 0, s>>>(Y, ldy, base, base_stride, n, mandel, out, ldo, info); \
        break;
    switch (d) {
        GABO_MAP_CASE(2)
        GABO_MAP_CASE(3)
        GABO_MAP_CASE(4)
        GABO_MAP_CASE(5)
        GABO_MAP_CASE(6)
        GABO_MAP_CASE(7)
        GABO_MAP_CASE(8)
        default:
            return -1;
    }
#undef GABO_MAP_CASE
    GABO_LAUNCH_CHECK();
    return 0;
}

int check_map_args(int d, const double* Y, int64_t ldy, const double* base, int64_t base_stride, int64_t n, int mandel,
                   double* out, int64_t ldo) {
    if (d < 2 || d > 8) return -1;
    const int64_t w = mandel ? (int64_t)d * (d + 1) / 2 : (int64_t)d * d;
    if (n < 0) return -6;
    if (n == 0) return 0;
    if (Y == nullptr) return -2;
    if (ldy < w) return -3;
    if (base == nullptr) return -4;
    if (base_stride != 0 && base_stride < w) return -5;
    if (out == nullptr) return -8;
    if (ldo < w) return -9;
    return 1;
}

}  // namespace
}  // namespace gabo

extern "C" int gabo_spd_expmap_f64(int d, const double* U, int64_t ldu, const double* base, int64_t base_stride, int64_t n,
                                   int mandel, double* out, int64_t ldo, int32_t* info, gabo_stream_t stream) {
    const int c = gabo::check_map_args(d, U, ldu, base, base_stride, n, mandel, out, ldo);
    if (c <= 0) return c;
    return gabo::launch_spd_map<false>(d, U, ldu, base, base_stride, n, mandel, out, ldo, info, static_cast<cudaStream_t>(stream));
}

extern "C" int gabo_spd_logmap_f64(int d, const double* X, int64_t ldx, const double* base, int64_t base_stride, int64_t n,
                                   int mandel, double* out, int64_t ldo, int32_t* info, gabo_stream_t stream) {
    const int c = gabo::check_map_args(d, X, ldx, base, base_stride, n, mandel, out, ldo);
    if (c <= 0) return c;
    return gabo::launch_spd_map<true>(d, X, ldx, base, base_stride, n, mandel, out, ldo, info, static_cast<cudaStream_t>(stream));
}

extern "C" int gabo_sym_to_mandel_f64(int d, const double* S, int64_t lds, int64_t n, double* out, int64_t ldo, gabo_stream_t stream) {
    if (d < 2 || d > 8) return -1;
    if (n < 0) return -4;
    if (n == 0) return 0;
    if (S == nullptr) return -2;
    if (lds < (int64_t)d * d) return -3;
    if (out == nullptr) return -5;
    if (ldo < (int64_t)d * (d + 1) / 2) return -6;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)gabo::ceil_div64(n, 128);
    switch (d) {
        case 2: gabo::sym_to_mandel_kernel<2><<<grid, 128, 0, s>>>(S, lds, n, out, ldo); break;
        case 3: gabo::sym_to_mandel_kernel<3><<<grid, 128, 0, s>>>(S, lds, n, out, ldo); break;
        case 4: gabo::sym_to_mandel_kernel<4><<<grid, 128, 0, s>>>(S, lds, n, out, ldo); break;
        case 5: gabo::sym_to_mandel_kernel<5><<<grid, 128, 0, s>>>(S, lds, n, out, ldo); break;
        case 6: gabo::sym_to_mandel_kernel<6><<<grid, 128, 0, s>>>(S, lds, n, out, ldo); break;
        case 7: gabo::sym_to_mandel_kernel<7><<<grid, 128, 0, s>>>(S, lds, n, out, ldo); break;
        default: gabo::sym_to_mandel_kernel<8><<<grid, 128, 0, s>>>(S, lds, n, out, ldo); break;
    }
    GABO_LAUNCH_CHECK();
    return 0;
}
