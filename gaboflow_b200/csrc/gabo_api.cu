// Small C-ABI entry points that are not kernels of their own: version, device query, status text, Kdiag fill.
#include <atomic>
#include <cstring>
#include <cuda.h>
#include "gabo_common.cuh"

namespace gabo {

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMsB200;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return kNumSMsB200;
    return n;
}

namespace {
__global__ void fill_kernel(double* out, int64_t n, double v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}
}  // namespace
}  // namespace gabo

extern "C" int gabo_abi_version(void) { return GABO_ABI_VERSION; }

extern "C" long long gabo_launch_count(void) { return gabo::g_launches.load(std::memory_order_relaxed); }

extern "C" int gabo_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    GABO_CUDA_TRY(cudaGetDevice(&dev));
    if (sm_count) GABO_CUDA_TRY(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major) GABO_CUDA_TRY(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) GABO_CUDA_TRY(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    return 0;
}

extern "C" const char* gabo_status_string(int status) {
    if (status == 0) return "ok";
    if (status <= GABO_ERR_CUDA) return cudaGetErrorString((cudaError_t)(GABO_ERR_CUDA - status));
    if (status < 0) return "invalid argument (index = -status, 1-based)";
    return "numerical failure (LAPACK-style info > 0)";
}

// Kdiag = fill([n], variance): kernels_sphere_tf.py:105,221; kernels_spd_tf.py:265,417,550,658.
extern "C" int gabo_kdiag_const_f64(int64_t n, double variance, double* out, gabo_stream_t stream) {
    if (n < 0) return -1;
    if (n == 0) return 0;
    if (out == nullptr) return -3;
    gabo::fill_kernel<<<(unsigned)gabo::ceil_div64(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, variance);
    GABO_LAUNCH_CHECK();
    return 0;
}

// Lets kernels on the current device dereference memory of `peer_device` (needed before gabo_points_gram_cyclic_f64 when the
// peer buffers were opened through CUDA IPC under the exporter's device).  "Already enabled" is not an error.
extern "C" int gabo_enable_peer_access(int peer_device) {
    int dev = 0, can = 0;
    GABO_CUDA_TRY(cudaGetDevice(&dev));
    if (peer_device == dev) return 0;
    GABO_CUDA_TRY(cudaDeviceCanAccessPeer(&can, dev, peer_device));
    if (!can) return -1;
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return 0; }
    if (e != cudaSuccess) return GABO_ERR_CUDA - (int)e;
    return 0;
}

// ---- peer-shared buffers (multi-GPU mirrored Gram): plain cudaMalloc so that the allocation base can be exported with
// CUDA IPC, and opened by the other ranks ON THEIR OWN DEVICE with lazy peer access, which is what makes kernel stores
// from that device into this memory legal.
extern "C" int gabo_shared_alloc(size_t bytes, void** out) {
    if (out == nullptr) return -2;
    *out = nullptr;
    if (bytes == 0) return 0;
    GABO_CUDA_TRY(cudaMalloc(out, bytes));
    return 0;
}
extern "C" int gabo_shared_free(void* ptr) {
    if (ptr != nullptr) GABO_CUDA_TRY(cudaFree(ptr));
    return 0;
}
extern "C" int gabo_ipc_export(void* ptr, unsigned char* handle64) {
    if (ptr == nullptr) return -1;
    if (handle64 == nullptr) return -2;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    GABO_CUDA_TRY(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle64, &h, 64);
    return 0;
}
extern "C" int gabo_ipc_open(const unsigned char* handle64, void** out) {
    if (handle64 == nullptr) return -1;
    if (out == nullptr) return -2;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    GABO_CUDA_TRY(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int gabo_ipc_close(void* ptr) {
    if (ptr != nullptr) GABO_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return 0;
}

// ---- stream-ordered flags between ranks (distributed factorisation over peer memory) ------------------------------------
// signal: a one-warp kernel stores `value` into up to GABO_MAX_PEERS 32-bit flags (peers' memory mapped with CUDA IPC)
// after a system-scope fence, i.e. after everything earlier in the stream is visible to the peers.
// wait:   cuStreamWaitValue32(>=) on a flag in THIS device's memory: the stream stalls without occupying an SM, so a rank
// waiting for its peers never takes resources from the kernels that are running.
namespace {
struct FlagList { uint32_t* f[GABO_MAX_PEERS]; int n; };
__global__ void flag_signal_kernel(const FlagList fl, uint32_t value) {
    if ((int)threadIdx.x < fl.n) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(fl.f[threadIdx.x]), "r"(value) : "memory");
    }
}
typedef CUresult (*PFN_waitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
PFN_waitValue32 get_wait_value32() {
    static PFN_waitValue32 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_waitValue32>(ptr);
        else
            (void)cudaGetLastError();
    }
    return fn;
}
}  // namespace

extern "C" int gabo_flag_signal(uint32_t* const* flags_host, int n, uint32_t value, gabo_stream_t stream_) {
    if (n < 0 || n > GABO_MAX_PEERS) return -2;
    if (n == 0) return 0;
    if (flags_host == nullptr) return -1;
    FlagList fl;
    fl.n = n;
    for (int i = 0; i < n; ++i) {
        if (flags_host[i] == nullptr) return -1;
        fl.f[i] = flags_host[i];
    }
    flag_signal_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream_)>>>(fl, value);
    GABO_LAUNCH_CHECK();
    return 0;
}

// Small payload + flag in one launch: copies `count` doubles from src into each of the n destination buffers (peers' memory)
// and then raises the n flags, after a system-scope fence.  The substitution of the distributed solve publishes each solved
// 512-block this way: one launch on the critical path instead of n copy-engine transfers + a signal kernel.
namespace {
struct PushList { double* dst[GABO_MAX_PEERS]; uint32_t* f[GABO_MAX_PEERS]; int n; };
__global__ void __launch_bounds__(256) push_signal_kernel(const double* __restrict__ src, int64_t count, const PushList pl, uint32_t value) {
    for (int p = 0; p < pl.n; ++p)
        for (int64_t i = threadIdx.x; i < count; i += blockDim.x) pl.dst[p][i] = src[i];
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < pl.n) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(pl.f[threadIdx.x]), "r"(value) : "memory");
    }
}
}  // namespace

extern "C" int gabo_push_signal_f64(const double* src, int64_t count, double* const* dst_host, uint32_t* const* flags_host, int n,
                                    uint32_t value, gabo_stream_t stream_) {
    if (count < 0) return -2;
    if (n < 0 || n > GABO_MAX_PEERS) return -5;
    if (n == 0) return 0;
    if (src == nullptr && count > 0) return -1;
    if (dst_host == nullptr) return -3;
    if (flags_host == nullptr) return -4;
    PushList pl;
    pl.n = n;
    for (int i = 0; i < n; ++i) {
        if (dst_host[i] == nullptr) return -3;
        if (flags_host[i] == nullptr) return -4;
        pl.dst[i] = dst_host[i];
        pl.f[i] = flags_host[i];
    }
    push_signal_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream_)>>>(src, count, pl, value);
    GABO_LAUNCH_CHECK();
    return 0;
}

extern "C" int gabo_flag_wait_geq(const uint32_t* flag, uint32_t value, gabo_stream_t stream_) {
    if (flag == nullptr) return -1;
    PFN_waitValue32 fn = get_wait_value32();
    if (fn == nullptr) return GABO_ERR_CUDA - (int)cudaErrorNotSupported;
    const CUresult r = fn(static_cast<CUstream>(stream_), reinterpret_cast<CUdeviceptr>(flag), value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) return GABO_ERR_CUDA - (int)cudaErrorNotSupported;
    return 0;
}

// Asynchronous copy between two device buffers of this process' address space (local or IPC-mapped peer memory): goes through
// the copy engines, no SM involved.
extern "C" int gabo_copy_async(void* dst, const void* src, size_t bytes, gabo_stream_t stream_) {
    if (bytes == 0) return 0;
    if (dst == nullptr) return -1;
    if (src == nullptr) return -2;
    GABO_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream_)));
    return 0;
}
