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

__device__ __forceinline__ double asin_poly(double z) {
    double p = 0x1.cd864394d2ff2p-6;
    p = fma(p, z, -0x1.603991d6060e0p-7);
    p = fma(p, z, 0x1.06b9d26d10838p-6);
    p = fma(p, z, 0x1.ff5fc4d14c735p-8);
    p = fma(p, z, 0x1.8522ddffa6208p-7);
    p = fma(p, z, 0x1.c87265d47ef49p-7);
    p = fma(p, z, 0x1.1c593c7b1d958p-6);
    p = fma(p, z, 0x1.6e8b2b3b10be4p-6);
    p = fma(p, z, 0x1.f1c71f95269afp-6);
    p = fma(p, z, 0x1.6db6db684b6a1p-5);
    p = fma(p, z, 0x1.3333333336da5p-4);
    p = fma(p, z, 0x1.555555555554fp-3);
    return p * z;
}

// acos for t already clamped to [-1, 1]
__device__ __forceinline__ double acos_fast(double t) {
    const double a = fabs(t);
    if (a <= 0.5) {
        const double w = asin_poly(t * t);
        return (0x1.921fb54442d18p+0 - t) - t * w;          // pi/2 - asin(t)
    }
    const double z = fma(-0.5, a, 0.5);
    const double w = asin_poly(z);
    const double s = sqrt(z);
    const double r = fma(s, w, s);
    return (t > 0.0) ? 2.0 * r : fma(-2.0, r, 0x1.921fb54442d18p+1);
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

// exp(x) for x <= 0, `tab` = shared-memory copy of kExp2Tab.  Falls back to libm below -700 (denormal range).
__device__ __forceinline__ double exp_neg_fast(double x, const double* __restrict__ tab) {
    if (!(x > -700.0)) return exp(x);                         // also NaN
    const double MAGIC = 6755399441055744.0;                  // 1.5 * 2^52
    const double tn = fma(x, 0x1.71547652b82fep+5, MAGIC);    // x * 32/ln2, rounded to integer in the low bits
    const int n = __double2loint(tn);
    const double nf = tn - MAGIC;
    double r = fma(nf, -0x1.62e42fefa0000p-6, x);             // ln2/32 split hi (36 bits) + lo
    r = fma(nf, -0x1.cf79abc9e3b3ap-45, r);
    double p = 1.0 / 720.0;
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double v = tab[n & 31] * p;
    const int m = n >> 5;                                     // arithmetic shift: floor division
    return __hiloint2double(__double2hiint(v) + (m << 20), __double2loint(v));
}

}  // namespace gabo
