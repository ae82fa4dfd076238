// FP64 pipe probe for B200 (sm_100a): how fast are DFMA and DMMA.8x8x4, and do they overlap?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu
// The answers pick the inner loops of the Gram tile kernel and of the Cholesky trailing update.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void k_dmma(double* out, int iters, double a, double b) {
    double c0[CHAINS], c1[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { c0[i] = threadIdx.x * 1e-9 + i; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA and DFMA interleaved: NM mma chains + NF fma chains per thread.
template <int NM, int NF>
__global__ void k_mixed(double* out, int iters, double a, double b) {
    double c0[NM + 1], c1[NM + 1], f[NF + 1];
#pragma unroll
    for (int i = 0; i < NM; ++i) { c0[i] = threadIdx.x * 1e-9 + i; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NM; ++i) dmma884(c0[i], c1[i], a, b);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; ++i) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// transcendental cost: acos + exp per element (the Gram epilogue)
__global__ void k_trans(double* out, int iters, double t0, double beta) {
    double t = t0 + threadIdx.x * 1e-4, s = 0;
    for (int it = 0; it < iters; ++it) {
        double d = acos(t);
        s += exp(-beta * d * d);
        t = t * 0.999 + 1e-5;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_acos(double* out, int iters, double t0) {
    double t = t0 + threadIdx.x * 1e-4, s = 0;
    for (int it = 0; it < iters; ++it) { s += acos(t); t = t * 0.999 + 1e-5; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_exp(double* out, int iters, double t0) {
    double t = t0 + threadIdx.x * 1e-4, s = 0;
    for (int it = 0; it < iters; ++it) { s += exp(-t); t = t * 0.999 + 1e-5; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s  SMs %d  cc %d.%d\n", p.name, sms, p.major, p.minor);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int bps : {1, 2}) {
            if (threads * bps > 2048) continue;
            int grid = sms * bps;
            float ms = time_ms([&] { k_dfma<8><<<grid, threads>>>(out, iters, 0.999, 1e-3); });
            double fl = 2.0 * 8 * iters * (double)threads * grid;
            printf("DFMA  chains=8 threads=%4d cta/sm=%d : %8.3f ms  %7.2f TFLOP/s\n", threads, bps, ms, fl / ms * 1e-9);
            ms = time_ms([&] { k_dmma<8><<<grid, threads>>>(out, iters, 0.999, 1e-3); });
            fl = 2.0 * 256 * 8 * iters * (double)(threads / 32) * grid;
            printf("DMMA  chains=8 threads=%4d cta/sm=%d : %8.3f ms  %7.2f TFLOP/s\n", threads, bps, ms, fl / ms * 1e-9);
        }
    }
    {
        int threads = 512, grid = sms * 2;
        float ms = time_ms([&] { k_dmma<16><<<grid, threads>>>(out, iters, 0.999, 1e-3); });
        double fl = 2.0 * 256 * 16 * iters * (double)(threads / 32) * grid;
        printf("DMMA  chains=16 threads=512 cta/sm=2 : %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_dmma<2><<<grid, threads>>>(out, iters, 0.999, 1e-3); });
        fl = 2.0 * 256 * 2 * iters * (double)(threads / 32) * grid;
        printf("DMMA  chains=2  threads=512 cta/sm=2 : %8.3f ms  %7.2f TFLOP/s (latency probe)\n", ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_dmma<1><<<sms, 32>>>(out, iters, 0.999, 1e-3); });
        printf("DMMA  dependent chain, 1 warp/SM: %.1f ns per DMMA\n", ms * 1e6 / iters);
        ms = time_ms([&] { k_dfma<1><<<sms, 32>>>(out, iters, 0.999, 1e-3); });
        printf("DFMA  dependent chain, 1 warp/SM: %.1f ns per DFMA\n", ms * 1e6 / iters);
        // mixed
        float m1 = time_ms([&] { k_mixed<8, 0><<<grid, threads>>>(out, iters, 0.999, 1e-3); });
        float m2 = time_ms([&] { k_mixed<0, 8><<<grid, threads>>>(out, iters, 0.999, 1e-3); });
        float m3 = time_ms([&] { k_mixed<8, 8><<<grid, threads>>>(out, iters, 0.999, 1e-3); });
        printf("mixed: 8 DMMA only %.3f ms | 8 DFMA only %.3f ms | both %.3f ms (sum %.3f => %s)\n", m1, m2, m3, m1 + m2,
               m3 < 0.8f * (m1 + m2) ? "OVERLAP" : "SHARED PIPE");
        const int it2 = 2000;
        double n = (double)it2 * threads * grid;
        ms = time_ms([&] { k_trans<<<grid, threads>>>(out, it2, 0.1, 2.0); });
        printf("acos+exp: %.3f ms  %.3e elem/s  (%.1f ps/elem)\n", ms, n / ms * 1e3, ms * 1e9 / n);
        ms = time_ms([&] { k_acos<<<grid, threads>>>(out, it2, 0.1); });
        printf("acos    : %.3f ms  %.3e elem/s\n", ms, n / ms * 1e3);
        ms = time_ms([&] { k_exp<<<grid, threads>>>(out, it2, 0.1); });
        printf("exp     : %.3f ms  %.3e elem/s\n", ms, n / ms * 1e3);
    }
    CK(cudaFree(out));
    return 0;
}
