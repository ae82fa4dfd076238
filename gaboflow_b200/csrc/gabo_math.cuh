// Lean FP64 acos / exp for the Gram epilogues (sm_100a).  The FP64 pipe is the binding resource of the sphere Gram
// (profiles/r01_fp64_probe.txt: DFMA and DMMA share one 64 FMA/clk/SM pipe), and CUDA's libm acos+exp cost ~55
// pipe slots per entry and ~150 instructions of straight-line code per call; these versions cost ~30 slots and are
// small enough that the epilogue loop stays resident in the instruction cache.
//   acos : |t| <= 0.5 -> pi/2 - (t + t z P(z)), z = t^2;  else s = sqrt((1-|t|)/2), r = s + s z P(z), 2r or pi - 2r.
//          P = degree-11 near-minimax on [0, 0.25] (tools/gen_math_coeffs.py); |error| < 3e-16 absolute.
//   exp  : x <= 0; x = (32 m + j) ln2/32 + r, |r| <= ln2/64; exp = 2^m * T[j] * poly6(r); relative error < 3e-16.
// Both are validated against CUDA libm / mpmath in tests (test_gpu_gram.py::test_fast_math_accuracy).
#pragma once
#include <cuda_runtime.h>

namespace gabo {

// Coefficients live in constant memory so that each DFMA takes its constant as a c[bank][offset] operand; as literals
// they cost two extra UMOV per FMA (ncu: 37 % of the issued instructions of the first version were UMOV/IMAD moves).
static __constant__ double kAsinC[12] = {
    0x1.555555555554fp-3, 0x1.3333333336da5p-4, 0x1.6db6db684b6a1p-5, 0x1.f1c71f95269afp-6,
    0x1.6e8b2b3b10be4p-6, 0x1.1c593c7b1d958p-6, 0x1.c87265d47ef49p-7, 0x1.8522ddffa6208p-7,
    0x1.ff5fc4d14c735p-8, 0x1.06b9d26d10838p-6, -0x1.603991d6060e0p-7, 0x1.cd864394d2ff2p-6};
static __constant__ double kExpC[12] = {
    6755399441055744.0,        // [0] 1.5 * 2^52
    0x1.71547652b82fep+5,      // [1] 32 / ln2
    -0x1.62e42fefa0000p-6,     // [2] -ln2/32 hi (36 bits)
    -0x1.cf79abc9e3b3ap-45,    // [3] -ln2/32 lo
    1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0,   // [4..9] Taylor coefficients of exp on |r| <= ln2/64
    0x1.921fb54442d18p+0,      // [10] pi/2
    0x1.921fb54442d18p+1};     // [11] pi

__device__ __forceinline__ double asin_poly(double z) {
    double p = kAsinC[11];
#pragma unroll
    for (int i = 10; i >= 0; --i) p = fma(p, z, kAsinC[i]);
    return p * z;
}

// acos for t already clamped to [-1, 1].  The |t| > 0.5 path (needs a square root) is entered warp-uniformly so the
// common case carries no divergence bookkeeping.
__device__ __forceinline__ double acos_fast(double t) {
    const double a = fabs(t);
    const bool big = a > 0.5;
    const double z = big ? fma(-0.5, a, 0.5) : t * t;
    const double w = asin_poly(z);
    double r = (kExpC[10] - t) - t * w;                     // pi/2 - asin(t), valid for |t| <= 0.5
    if (__any_sync(0xffffffffu, big)) {
        const double s = sqrt(z);
        const double h = fma(s, w, s);                      // asin(sqrt((1-|t|)/2))
        const double rb = (t > 0.0) ? 2.0 * h : fma(-2.0, h, kExpC[11]);
        r = big ? rb : r;
    }
    return r;
}

// the 32-entry table 2^(j/32); kernels copy it to shared memory once (divergent constant-bank reads would serialise)
static __constant__ double kExp2Tab[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};

// exp(x) for x <= 0, `tab` = shared-memory copy of kExp2Tab.  Falls back to libm below -700 (denormal range) or NaN,
// warp-uniformly.
__device__ __forceinline__ double exp_neg_fast(double x, const double* __restrict__ tab) {
    const double tn = fma(x, kExpC[1], kExpC[0]);             // x * 32/ln2, rounded to integer in the low bits
    const int n = __double2loint(tn);
    const double nf = tn - kExpC[0];
    double r = fma(nf, kExpC[2], x);
    r = fma(nf, kExpC[3], r);
    double p = kExpC[4];
#pragma unroll
    for (int i = 5; i <= 9; ++i) p = fma(p, r, kExpC[i]);
    p = fma(p, r, kExpC[9]);
    const double v = tab[n & 31] * p;
    const int m = n >> 5;                                     // arithmetic shift: floor division
    double res = __hiloint2double(__double2hiint(v) + (m << 20), __double2loint(v));
    const bool slow = !(x > -700.0);
    if (__any_sync(0xffffffffu, slow)) {
        if (slow) res = exp(x);
    }
    return res;
}

// ---- batched forms: N independent entries, straight-line (the compiler interleaves the N dependency chains), with ONE
// warp vote per batch for the rare slow paths (a vote per entry is a scheduling barrier and serialises the chains).
// Range tests are done on the high 32 bits with INTEGER compares: an FP64 compare (DSETP) or min/max occupies the same
// FP64 pipe as the DFMAs and the DMMAs, and that pipe is what bounds the Gram kernel.
__device__ __forceinline__ int hi_abs(double x) { return __double2hiint(x) & 0x7fffffff; }

// acos(clamp(t, -1, 1)) for N values (the clamp of sphere_utils_tf.py:59 is part of the big-|t| slow path).
template <int N>
__device__ __forceinline__ void acos_clamped_fast_n(const double (&t)[N], double (&out)[N]) {
    double w[N];
    bool anybig = false;
#pragma unroll
    for (int i = 0; i < N; ++i) anybig = anybig || (hi_abs(t[i]) > 0x3fe00000);     // |t| > 0.5 (or NaN)
#pragma unroll
    for (int i = 0; i < N; ++i) w[i] = kAsinC[11];
#pragma unroll
    for (int c = 10; c >= 0; --c)
#pragma unroll
        for (int i = 0; i < N; ++i) w[i] = fma(w[i], t[i] * t[i], kAsinC[c]);
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = (kExpC[10] - t[i]) - t[i] * (w[i] * (t[i] * t[i]));   // pi/2 - asin(t), |t| <= 0.5
    if (__any_sync(0xffffffffu, anybig)) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (hi_abs(t[i]) > 0x3fe00000) {      // per-lane slow path: NO warp-level primitive may appear in here
                const double tc = fmin(fmax(t[i], -1.0), 1.0);                 // sphere_utils_tf.py:59
                const double z = fma(-0.5, fabs(tc), 0.5);
                const double s = sqrt(z);
                const double h = fma(s, asin_poly(z), s);                       // asin(sqrt((1-|t|)/2))
                const double rb = (tc > 0.0) ? 2.0 * h : fma(-2.0, h, kExpC[11]);
                out[i] = (t[i] == t[i]) ? rb : t[i];
            }
        }
    }
}

// out-of-line libm fallback of the batched exp (N inlined copies of exp() would sit in the middle of the hot loops' code)
// scale * e^x as (scale * e^(x/2)) * e^(x/2): e^(x/2) is still a normal number where e^x is subnormal, so a result in the
// subnormal range is rounded once, from full-precision factors (scale * exp(x) would multiply an already truncated subnormal).
static __device__ __noinline__ double exp_scaled_slow(double x, double scale) {
    const double r = exp(0.5 * x);
    return (scale * r) * r;
}

// exp(x) * scale for x <= 0; `tab` = shared-memory table 2^(j/32) ALREADY multiplied by `scale`.
// The power of two 2^m is applied by an integer add on the exponent field of v = tab[j] * poly, which is only valid while
// the result stays a NORMAL number.  v carries the exponent of `scale` (or one more), so the result is normal whenever
// m >= 2 - E with E the biased exponent of scale, i.e. |x| <= (E - 2) ln 2: entries beyond `slow_hi` -- the high word of
// min(700, (E - 2) ln 2), from exp_slow_threshold_hi(scale), computed once per thread -- take the libm slow path, as do
// NaN / inf.  (Round 1 tested |x| >= 700 only: with scale = 1e-6 and x = -699 the exponent field wrapped to -8.7e306.)
__device__ __forceinline__ int exp_slow_threshold_hi(double scale) {
    const int E = (__double2hiint(scale) >> 20) & 0x7ff;           // 0: zero / subnormal scale, 0x7ff: inf / NaN
    double lim = (double)(E - 2) * 0.693147180559945;
    if (E == 0x7ff || !(scale > 0.0)) lim = 0.0;
    lim = fmin(fmax(lim, 0.0), 700.0);
    return __double2hiint(lim);                                      // |x| >= lim  <=>  hi_abs(x) >= hi(lim) (conservatively)
}

template <int N>
__device__ __forceinline__ void exp_neg_scaled_fast_n(const double (&x)[N], double (&out)[N], const double* __restrict__ tab,
                                                      double scale, int slow_hi) {
    double r[N], p[N];
    int n[N];
    bool anyslow = false;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double tn = fma(x[i], kExpC[1], kExpC[0]);
        n[i] = __double2loint(tn);
        const double nf = tn - kExpC[0];
        r[i] = fma(nf, kExpC[3], fma(nf, kExpC[2], x[i]));
        p[i] = kExpC[4];
        anyslow = anyslow || (hi_abs(x[i]) >= slow_hi);
    }
#pragma unroll
    for (int c = 5; c <= 9; ++c)
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], kExpC[c]);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        p[i] = fma(p[i], r[i], kExpC[9]);
        const double v = tab[n[i] & 31] * p[i];
        out[i] = __hiloint2double(__double2hiint(v) + ((n[i] >> 5) << 20), __double2loint(v));
    }
    if (__any_sync(0xffffffffu, anyslow)) {
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (hi_abs(x[i]) >= slow_hi) out[i] = exp_scaled_slow(x[i], scale);
    }
}

template <int N>
__device__ __forceinline__ void acos_fast_n(const double (&t)[N], double (&out)[N]) { acos_clamped_fast_n<N>(t, out); }

template <int N>
__device__ __forceinline__ void exp_neg_fast_n(const double (&x)[N], double (&out)[N], const double* __restrict__ tab) {
    exp_neg_scaled_fast_n<N>(x, out, tab, 1.0, 0x4085e000);
}

// ---- lean natural logarithm for POSITIVE, NORMAL, finite x (the eigenvalues of an SPD congruence) ------------------------
//   x = 2^k m, m in [sqrt(1/2), sqrt(2));  s = (m-1)/(m+1);  log m = 2 atanh(s) = 2s + s z P(z), z = s^2 <= 0.0295,
//   P = the series 2/3 + 2/5 z + ... + 2/19 z^8 (next term < 1e-17);  the quotient is a hardware reciprocal seed, two
//   Newton steps and one residual correction.  ~30 instructions, no branches, constants as c[bank] operands; CUDA's log()
//   costs ~50 plus two UMOV per literal coefficient.  Error < 2e-16 relative (tests: test_fast_math_accuracy).
static __constant__ double kLogC[11] = {
    2.0 / 3.0, 2.0 / 5.0, 2.0 / 7.0, 2.0 / 9.0, 2.0 / 11.0, 2.0 / 13.0, 2.0 / 15.0, 2.0 / 17.0, 2.0 / 19.0,
    0x1.62e42fefa3800p-1,      // [9]  ln2 hi (41 bits: k * hi is exact for |k| < 2^11)
    0x1.ef35793c76730p-45};    // [10] ln2 lo

__device__ __forceinline__ double log_pos_normal(double x) {
    const int t = __double2hiint(x) + 0x00095f62;                       // 0x3ff00000 - 0x3fe6a09e
    const int k = (t >> 20) - 1023;
    const double m = __hiloint2double((t & 0x000fffff) + 0x3fe6a09e, __double2loint(x));
    const double f = m - 1.0, den = m + 1.0;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    r = fma(fma(-den, r, 1.0), r, r);
    r = fma(fma(-den, r, 1.0), r, r);
    double s = f * r;
    s = fma(fma(-s, den, f), r, s);
    const double z = s * s;
    double p = kLogC[8];
#pragma unroll
    for (int i = 7; i >= 0; --i) p = fma(p, z, kLogC[i]);
    const double kd = (double)k;
    const double tail = fma(kd, kLogC[10], (s * z) * p);
    return fma(kd, kLogC[9], s + s) + tail;
}

}  // namespace gabo
