// Microbenchmark: dependent-issue latency of DFMA / DADD / DMUL / FFMA / FFMA2 / MUFU.RSQ on sm_100a (one warp per scheduler,
// one dependency chain), and the same with 2 and 4 independent chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/latency_probe tools/latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 8192;
template <int MODE, int NCH>
__global__ void probe(double* out, double a, float b) {
    double x[NCH]; float y[NCH]; unsigned long long z[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { x[i] = a + i; y[i] = b + i; z[i] = ((unsigned long long)__float_as_uint(b + i) << 32) | __float_as_uint(b); }
    const unsigned long long bb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (MODE == 0) x[i] = fma(x[i], a, a);
            if (MODE == 1) x[i] = x[i] + a;
            if (MODE == 2) x[i] = x[i] * a;
            if (MODE == 3) y[i] = fmaf(y[i], b, b);
            if (MODE == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(z[i]) : "l"(bb));
            if (MODE == 5) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(y[i]));
            if (MODE == 6) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x[i]));
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += x[i] + y[i] + (double)z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[100000] = (double)(t1 - t0) / ITERS / NCH;
}
template <int MODE, int NCH>
void run(const char* name, double* o) {
    probe<MODE, NCH><<<148, 128>>>(o, 1.0000001, 1.0001f);
    cudaDeviceSynchronize();
    double c; cudaMemcpy(&c, o + 100000, 8, cudaMemcpyDeviceToHost);
    printf("%-10s chains %d: %.2f cycles per instruction per warp (=> %.2f cycles between dependent issues)\n", name, NCH, c, c * NCH);
}
int main() {
    double* o; cudaMalloc(&o, 8 * 200000);
    run<0, 1>("DFMA", o); run<0, 2>("DFMA", o); run<0, 4>("DFMA", o); run<0, 8>("DFMA", o);
    run<1, 1>("DADD", o); run<2, 1>("DMUL", o);
    run<3, 1>("FFMA", o); run<3, 4>("FFMA", o);
    run<4, 1>("FFMA2", o); run<4, 2>("FFMA2", o); run<4, 4>("FFMA2", o);
    run<5, 1>("MUFU.RSQ", o); run<5, 2>("MUFU.RSQ", o); run<5, 4>("MUFU.RSQ", o);
    run<6, 1>("RCP64H", o);
    return 0;
}
