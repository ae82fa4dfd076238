// Microbenchmark: does an FP64 instruction (half-rate pipe: 16 lanes per SM sub-partition) block the issue port for two cycles?
// Runs (a) DFMA only, (b) FFMA only, (c) DFMA + FFMA interleaved 1:1, (d) DFMA + 2 FFMA, (e) DFMA + IADD-class, each with
// NCH independent chains per thread, 4 and 8 warps per scheduler.  If issue is not blocked, (c) takes as long as (a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_probe tools/issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NCH = 8;
constexpr int ITERS = 4096;

template <int MODE>
__global__ void probe(double* outd, float* outf, double a, float b) {
    double x[NCH];
    float y[2 * NCH];
    int z[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { x[i] = a + i + threadIdx.x; y[i] = b + i; y[i + NCH] = b - i; z[i] = threadIdx.x + i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 4) x[i] = fma(x[i], a, a);
            if (MODE == 1 || MODE == 2 || MODE == 3) y[i] = fmaf(y[i], b, b);
            if (MODE == 3) y[i + NCH] = fmaf(y[i + NCH], b, b);
            if (MODE == 4 || MODE == 5) z[i] = (z[i] ^ it) + 0x9e37;
            if (MODE == 6) z[i] = z[i] * 0x80 + it;
        }
        // modes 7 / 8: the shape of the int8 Gram epilogue -- 88 DFMA and 176 integer instructions per iteration, either as two
        // bursts (7: what a warp's instruction stream looks like when the phases follow each other) or finely interleaved (8)
        if (MODE == 7) {
#pragma unroll
            for (int r = 0; r < 11; ++r)
#pragma unroll
                for (int i = 0; i < NCH; ++i) x[i] = fma(x[i], a, a);
            asm volatile("" ::: "memory");
#pragma unroll
            for (int r = 0; r < 11; ++r)
#pragma unroll
                for (int i = 0; i < NCH; ++i) z[i] = ((z[i] ^ it) + 0x9e37) * 0x80 + r;
            asm volatile("" ::: "memory");
        }
        if (MODE == 8) {
#pragma unroll
            for (int r = 0; r < 11; ++r)
#pragma unroll
                for (int i = 0; i < NCH; ++i) { x[i] = fma(x[i], a, a); z[i] = ((z[i] ^ it) + 0x9e37) * 0x80 + r; }
        }
    }
    double s = 0; float t = 0; int u = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) { s += x[i]; t += y[i] + y[i + NCH]; u += z[i]; }
    outd[blockIdx.x * blockDim.x + threadIdx.x] = s + u;
    outf[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE>
void run(const char* name, int threads, double* od, float* of) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    probe<MODE><<<sms, threads>>>(od, of, 1.0000001, 1.0001f);
    cudaEventRecord(e0);
    probe<MODE><<<sms, threads>>>(od, of, 1.0000001, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk * 1e3;
    double warps_per_smsp = threads / 32 / 4.0;
    double slots = (double)ITERS * NCH * warps_per_smsp;   // loop-body groups per scheduler
    printf("%-28s threads/SM %4d  %.3f ms  cycles per (chain-step x warp) per scheduler: %.2f\n", name, threads, ms, cycles / slots);
}

int main() {
    double* od; float* of;
    cudaMalloc(&od, 148 * 1024 * 8); cudaMalloc(&of, 148 * 1024 * 4);
    for (int threads : {128, 512, 1024}) {
        run<0>("DFMA", threads, od, of);
        run<1>("FFMA", threads, od, of);
        run<2>("DFMA+FFMA", threads, od, of);
        run<3>("DFMA+2FFMA", threads, od, of);
        run<4>("DFMA+2INT(LOP,IADD)", threads, od, of);
        run<5>("2INT(LOP,IADD)", threads, od, of);
        run<6>("IMAD", threads, od, of);
        run<7>("88 DFMA then 88x(LOP,IADD,IMAD)", threads, od, of);
        run<8>("88x(DFMA,LOP,IADD,IMAD) mixed", threads, od, of);
    }
    return 0;
}
