// Throughput of the lean batched acos + exp of the Gram epilogues (csrc/gabo_math.cuh) on register data, no memory traffic:
// the ceiling of any sphere Gram kernel's epilogue.  Reports ps per element and FP64-pipe slots per element equivalent.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gaboflow_b200/csrc -o tools/epi_math_probe tools/epi_math_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "gabo_math.cuh"
using namespace gabo;

template <int NB, int WHAT>
__global__ void k(double* out, int iters, double t0, double nbeta) {
    __shared__ double exptab[32];
    if (threadIdx.x < 32) exptab[threadIdx.x] = kExp2Tab[threadIdx.x];
    __syncthreads();
    double t[NB], acc = 0.0;
#pragma unroll
    for (int i = 0; i < NB; ++i) t[i] = t0 + 1e-3 * i + 1e-6 * threadIdx.x;
    const int slow_hi = exp_slow_threshold_hi(1.0);
    for (int it = 0; it < iters; ++it) {
        double d[NB], x[NB], v[NB];
        if (WHAT & 1) acos_clamped_fast_n<NB>(t, d);
        else {
#pragma unroll
            for (int i = 0; i < NB; ++i) d[i] = t[i];
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) x[i] = (nbeta * d[i]) * d[i];
        if (WHAT & 2) exp_neg_scaled_fast_n<NB>(x, v, exptab, 1.0, slow_hi);
        else {
#pragma unroll
            for (int i = 0; i < NB; ++i) v[i] = x[i];
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) { acc += v[i]; t[i] = fma(t[i], 0.999, 1e-5); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int NB, int WHAT>
void run(const char* name, int threads, double* out) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NB, WHAT><<<sms, threads>>>(out, iters, 0.1, -2.0);
    cudaEventRecord(e0);
    k<NB, WHAT><<<sms, threads>>>(out, iters, 0.1, -2.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)sms * threads * iters * NB;
    const double ps = ms * 1e9 / n;
    printf("%-22s batch %d threads/SM %4d: %.3f ms  %.2f ps/elem = %.1f FP64-pipe slots/elem (64 lanes x 1.965 GHz x %d SMs)\n", name, NB, threads, ms, ps,
           ps * 1e-12 * 64 * 1.965e9 * sms, sms);
}

int main() {
    double* out; cudaMalloc(&out, 148 * 1024 * 8);
    for (int threads : {128, 256, 512, 1024}) {
        run<8, 3>("acos+exp", threads, out);
        run<4, 3>("acos+exp", threads, out);
        run<8, 1>("acos", threads, out);
        run<8, 2>("exp", threads, out);
        run<8, 0>("neither (loop only)", threads, out);
    }
    return 0;
}
