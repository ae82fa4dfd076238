// Phase timing of the diagonal-block kernel's building block (warp-level 16x16 Cholesky + inverse), with clock64.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/potf2_probe tools/potf2_probe.cu
#include <cstdio>
#define GABO_POTF2_PROFILE 1
#include "../gaboflow_b200/csrc/linalg.cu"
namespace gabo { void count_launch() {} }
using namespace gabo;
__global__ void probe(const double* in, double* out, long long* cyc) {
    const int lane = threadIdx.x & 31;
    double a[BS], w[BS];
    for (int k = 0; k < BS; ++k) a[k] = (k <= (lane & 15)) ? in[(lane & 15) * BS + k] : 0.0;
    __syncwarp();
    long long t0 = clock64();
    int bad = warp_chol16_inv(a, w, lane);
    long long t1 = clock64();
    double s = bad;
    for (int k = 0; k < BS; ++k) s += a[k] + w[k];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double h[256];
    for (int i = 0; i < 16; ++i) for (int j = 0; j < 16; ++j) h[i * 16 + j] = (i == j) ? 20.0 : 1.0 / (1 + i + j);
    double *din, *dout; long long* dc;
    cudaMalloc(&din, sizeof(h)); cudaMalloc(&dout, 32 * 8); cudaMalloc(&dc, 8);
    cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 3; ++rep) {
        probe<<<1, 32>>>(din, dout, dc);
        long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        printf("warp_chol16_inv: %lld cycles (rep %d)\n", c, rep);
    }
    // whole kernel
    double *A, *W; int* info;
    cudaMalloc(&A, 128 * 128 * 8); cudaMalloc(&W, 128 * 128 * 8); cudaMalloc(&info, 4);
    double* hA = new double[128 * 128];
    for (int i = 0; i < 128; ++i) for (int j = 0; j < 128; ++j) hA[i * 128 + j] = (i == j) ? 200.0 : 1.0 / (1 + i + j);
    cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTF2V2_SMEM);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(A, hA, 128 * 128 * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        potf2_inv_kernel<<<1, 256, POTF2V2_SMEM>>>(A, 128, 128, 0.0, W, info, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("potf2_inv_kernel: %.1f us (%s)\n", ms * 1e3, cudaGetErrorString(cudaGetLastError()));
        long long st[8]; cudaMemcpyFromSymbol(st, g_potf2_stamps, sizeof(st));
        printf("  cycles: load %lld  factor %lld  store %lld  inverse %lld  winv-store %lld\n", st[1]-st[0], st[2]-st[1], st[3]-st[2], st[4]-st[3], st[5]-st[4]);
        long long ph[4]; cudaMemcpyFromSymbol(ph, g_potf2_phase, sizeof(ph));
        printf("  factor phases: chol16 %lld  panel %lld  trailing %lld\n", ph[0], ph[1], ph[2]);
    }
    return 0;
}
