// FP32 mode of the point-set Gram kernel on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::tf32 with the
// accumulators in TMEM, operands staged by TMA bulk copies, and the acos / exp epilogue reading the tile back with tcgen05.ld.
//
// Same contract as gabo_points_gram_f64 (reference sphere_utils_tf.py:11-60 + kernels_sphere_tf.py:59-88,179-204 and the pair
// stage of kernels_spd_tf.py:510-533,618-641) with float inputs and outputs; north_star tolerance for this mode: 1e-5 relative.
//
// Arithmetic: a TF32 product keeps 11 significand bits, which is not enough for 1e-5 after acos and exp, so each inner
// product is the classical three-term split  x.y ~ hi(x).hi(y) + hi(x).lo(y) + lo(x).hi(y)  with hi = rna-rounded TF32 and
// lo = x - hi (exact in FP32): three accumulating MMAs per k-step, error ~2^-21 |x||y| (measured < 3e-7 on unit vectors).
//
// Data layout: the pack kernel writes every 128-row tile of X as  [plane hi|lo][k-chunk][row 0..127][4 floats]  -- exactly
// the K-major, non-swizzled UMMA "core matrix" layout (8 rows x 16 B contiguous, next 8 rows +128 B, next 4 k's +2048 B) -- so
// a whole operand tile is ONE contiguous TMA bulk copy and the shared-memory descriptors are plain base + k * 4096.
//
// Kernel: persistent CTAs (one per SM), work item = up to 8 consecutive column tiles of one 128-row block (the A tile is
// loaded once per item, the B tiles stream through a 2-deep ring).  Warp 0 lane 0: TMA producer.  Warp 1 lane 0: issues the
// 3 x K/8 tcgen05.mma (M = N = 128, K = 8) per tile into one of two TMEM accumulator stages and commits to mbarriers.
// Warps 2..9: epilogue -- tcgen05.ld 32 columns of the thread's row, clamp / acos / exp in FP32, then the direct tile goes
// through a per-warp padded shared-memory transpose (128 B coalesced row stores) and the mirror image K[col][row] is stored
// straight from the registers (lane = row, so a register across the warp IS 32 consecutive floats of a mirrored row).
#include "gabo_common.cuh"

namespace gabo {
namespace {

constexpr int TT = 128;                 // tile side (UMMA M = N = 128)
constexpr int TC_THREADS = 320;         // warp 0 producer, warp 1 MMA, warps 2..9 epilogue
constexpr int SEG = 8;                  // column tiles per work item
constexpr int NBST = 2;                 // B-tile ring depth
constexpr int NACC = 2;                 // TMEM accumulator stages (128 columns each)
constexpr int CHUNK_BYTES = TT * 16;    // one k-chunk (4 floats) of all 128 rows

struct TcParams {
    const float* Ap;       // packed X  tiles
    const float* Bp;       // packed X2 tiles
    const float* normA;    // |x|^2 (sqdist kind) or nullptr
    const float* normB;
    float* K;
    int64_t ldk, row0, row1, n2;
    int kc;                // k-chunks per plane (K padded to a multiple of 8: kc = Kpad / 4, even)
    int kind, uplo, symmetric;
    float beta, variance;
    int64_t tile_r0, tiles_m, tiles_n;
    int64_t nitems;
};

// ---- packing: X [n, dim] -> tiles [n_pad/128][2][kc][128][4], hi / lo TF32 planes, zero padded ------------------------------
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__global__ void pack_tf32_kernel(const float* __restrict__ X, int64_t n, int64_t ldx, int dim, int kc, float* __restrict__ P,
                                 int64_t n_pad, float* __restrict__ norms) {
    // one warp per row
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (row >= n_pad) return;
    const int64_t tile = row / TT;
    const int r = (int)(row - tile * TT);
    float* base = P + tile * (int64_t)(2 * kc * TT * 4);
    float s = 0.0f;
    for (int k = lane; k < kc * 4; k += 32) {
        float v = 0.0f;
        if (row < n && k < dim) v = X[row * ldx + k];
        s = fmaf(v, v, s);
        const float hi = tf32_rna(v);
        const float lo = tf32_rna(v - hi);
        const int64_t off = (int64_t)(k >> 2) * (TT * 4) + r * 4 + (k & 3);
        base[off] = hi;
        base[(int64_t)kc * TT * 4 + off] = lo;
    }
    if (norms != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) norms[row] = s;
    }
}

// ---- tcgen05 / mbarrier PTX ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, M = N = 128, K = 8 (TF32)
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t zero = 0;    // disable-output-lane mask: none
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(zero)
        : "memory");
}
// K-major, no swizzle: 8 rows x 16 B contiguous core matrices; SBO (next 8 rows) = 128 B, LBO (next 16 B of K) = 2048 B
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
    d |= (uint64_t)(CHUNK_BYTES >> 4) << 16;     // leading byte offset
    d |= (uint64_t)(128 >> 4) << 32;             // stride byte offset
    d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
    return d;
}
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TT >> 3) << 17) | ((uint32_t)(TT >> 4) << 24);

__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
        "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- work items: (row tile bi, first column tile, number of column tiles <= SEG) ------------------------------------------------
struct ItemWalk {
    int64_t bi, seg;     // current row tile (local) and segment within it
};
__device__ __forceinline__ int64_t row_tiles(const TcParams& p, int64_t bi) {
    return (p.uplo == GABO_FULL) ? p.tiles_n : (p.tile_r0 + bi + 1);
}
__device__ __forceinline__ void walk_advance(const TcParams& p, ItemWalk& w, int64_t adv) {
    while (w.bi < p.tiles_m) {
        const int64_t left = (row_tiles(p, w.bi) + SEG - 1) / SEG - w.seg;
        if (adv < left) { w.seg += adv; return; }
        adv -= left;
        ++w.bi;
        w.seg = 0;
    }
}

// ---- FP32 epilogue math ---------------------------------------------------------------------------------------------------------
// asin(s) = s + s z P(z), z = s^2, P = degree-4 minimax on [0, 0.25] (|error| < 6e-8)
__device__ __forceinline__ float asin_poly_f32(float z) {
    float p = 0.0421632f;
    p = fmaf(p, z, 0.0241818f);
    p = fmaf(p, z, 0.0454700f);
    p = fmaf(p, z, 0.0749530f);
    return fmaf(p, z, 0.1666675f);
}
// acos on [-1, 1], branch-free: |t| <= 0.5: pi/2 - asin(t); else 2 asin(sqrt((1-|t|)/2)) (or pi minus it)
__device__ __forceinline__ float acos_f32(float t) {
    const float a = fabsf(t);
    const bool big = a > 0.5f;
    const float z = big ? fmaf(-0.5f, a, 0.5f) : t * t;
    float s;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(z));
    s = big ? s : a;
    const float r = fmaf(s * z, asin_poly_f32(z), s);                         // asin(s), s >= 0
    const float small_res = (t >= 0.0f) ? 1.57079632679f - r : 1.57079632679f + r;
    const float big_res = (t >= 0.0f) ? 2.0f * r : fmaf(-2.0f, r, 3.14159265359f);
    return big ? big_res : small_res;
}
__device__ __forceinline__ float ex2_f32(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// v[i] <- variance * exp(-beta * g(v[i])) for the 32 inner products (or, SQDIST, a.b values) of one thread.
// nb2 = -beta * log2(e).  Sphere kinds: when every |t| of the warp is <= 0.5 (always, for points in high dimension) the
// acos needs neither the clamp nor the square-root branch: 14 instructions per entry.
template <int KIND>
__device__ __forceinline__ void epilogue_math(float (&v)[32], float nb2, float variance, float na, const float* __restrict__ nb) {
    if (KIND == GABO_SQDIST_GAUSSIAN) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float d2 = fmaxf(fmaf(-2.0f, v[i], na + nb[i]), 0.0f);       // ||a-b||^2 = |a|^2 + |b|^2 - 2 a.b
            v[i] = variance * ex2_f32(nb2 * d2);
        }
        return;
    }
    bool anybig = false;
#pragma unroll
    for (int i = 0; i < 32; ++i) anybig = anybig || !(fabsf(v[i]) <= 0.5f);    // also true for NaN
    if (!__any_sync(0xffffffffu, anybig)) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float t = v[i], z = t * t;
            const float d = 1.57079632679f - fmaf(t * z, asin_poly_f32(z), t);   // sphere_utils_tf.py:60
            const float arg = (KIND == GABO_SPHERE_GAUSSIAN) ? (nb2 * d) * d : nb2 * d;   // kernels_sphere_tf.py:83-86, 199-202
            v[i] = variance * ex2_f32(arg);                                      // kernels_sphere_tf.py:88, 204
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float d = acos_f32(fminf(fmaxf(v[i], -1.0f), 1.0f));           // sphere_utils_tf.py:59-60
            const float arg = (KIND == GABO_SPHERE_GAUSSIAN) ? (nb2 * d) * d : nb2 * d;
            v[i] = variance * ex2_f32(arg);
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(TC_THREADS, 1) points_gram_f32_tc_kernel(const TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t tile_bytes = (uint32_t)(2 * p.kc) * CHUNK_BYTES;      // hi + lo planes
    unsigned char* const sA = smem_raw;
    unsigned char* const sB = smem_raw + tile_bytes;                      // NBST tiles
    float* const sStage = reinterpret_cast<float*>(smem_raw + (size_t)(1 + NBST) * tile_bytes);   // 8 warps x 32 x 33 floats
    __shared__ __align__(8) uint64_t bar_a_full, bar_a_empty, bar_b_full[NBST], bar_b_empty[NBST], bar_acc_full[NACC], bar_acc_empty[NACC];
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&bar_a_full, 1);
        mbar_init(&bar_a_empty, 1);
        for (int s = 0; s < NBST; ++s) { mbar_init(&bar_b_full[s], 1); mbar_init(&bar_b_empty[s], 1); }
        for (int s = 0; s < NACC; ++s) { mbar_init(&bar_acc_full[s], 1); mbar_init(&bar_acc_empty[s], 8); }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_smem)), "n"(NACC * TT));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    const int64_t first = blockIdx.x, stride = gridDim.x;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            ItemWalk w{0, 0};
            walk_advance(p, w, first);
            uint32_t a_phase = 0;
            int64_t bcount = 0;      // B tiles issued so far (ring position / phase)
            for (int64_t it = first; it < p.nitems; it += stride) {
                const int64_t T = row_tiles(p, w.bi);
                const int64_t bj0 = w.seg * SEG;
                const int cnt = (int)((T - bj0) < SEG ? (T - bj0) : SEG);
                mbar_wait(&bar_a_empty, a_phase ^ 1);                     // MMAs of the previous item have read A
                mbar_arrive_expect_tx(&bar_a_full, tile_bytes);
                tma_bulk_g2s(sA, reinterpret_cast<const unsigned char*>(p.Ap) + (p.tile_r0 + w.bi) * (int64_t)tile_bytes, tile_bytes, &bar_a_full);
                a_phase ^= 1;
                for (int j = 0; j < cnt; ++j, ++bcount) {
                    const int s = (int)(bcount % NBST);
                    const uint32_t ph = (uint32_t)((bcount / NBST) & 1);
                    mbar_wait(&bar_b_empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&bar_b_full[s], tile_bytes);
                    tma_bulk_g2s(sB + (size_t)s * tile_bytes, reinterpret_cast<const unsigned char*>(p.Bp) + (bj0 + j) * (int64_t)tile_bytes,
                                 tile_bytes, &bar_b_full[s]);
                }
                walk_advance(p, w, stride);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            ItemWalk w{0, 0};
            walk_advance(p, w, first);
            uint32_t a_phase = 0;
            int64_t bcount = 0;
            const uint32_t a_addr = smem_u32(sA);
            const uint32_t plane = (uint32_t)p.kc * CHUNK_BYTES;
            for (int64_t it = first; it < p.nitems; it += stride) {
                const int64_t T = row_tiles(p, w.bi);
                const int64_t bj0 = w.seg * SEG;
                const int cnt = (int)((T - bj0) < SEG ? (T - bj0) : SEG);
                mbar_wait(&bar_a_full, a_phase);
                a_phase ^= 1;
                for (int j = 0; j < cnt; ++j, ++bcount) {
                    const int s = (int)(bcount % NBST);
                    const uint32_t ph = (uint32_t)((bcount / NBST) & 1);
                    const int as = (int)(bcount % NACC);
                    const uint32_t aph = (uint32_t)((bcount / NACC) & 1);
                    mbar_wait(&bar_acc_empty[as], aph ^ 1);               // epilogue has drained this accumulator stage
                    mbar_wait(&bar_b_full[s], ph);
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(sB + (size_t)s * tile_bytes);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(as * TT);
                    uint32_t acc = 0;
                    // hi.hi + hi.lo + lo.hi, K = 8 (two k-chunks) per instruction
#pragma unroll 1
                    for (int term = 0; term < 3; ++term) {
                        const uint32_t ao = (term == 2) ? plane : 0u, bo = (term == 1) ? plane : 0u;
                        for (int k = 0; k < p.kc; k += 2) {
                            tc_mma_tf32(d_tmem, tc_smem_desc(a_addr + ao + (uint32_t)k * CHUNK_BYTES),
                                        tc_smem_desc(b_addr + bo + (uint32_t)k * CHUNK_BYTES), kIdescTf32, acc);
                            acc = 1;
                        }
                    }
                    tc_commit(&bar_b_empty[s]);                            // B slot free once these MMAs have read it
                    if (j == cnt - 1) tc_commit(&bar_a_empty);             // A free after the last tile of the item
                    tc_commit(&bar_acc_full[as]);                          // accumulator ready for the epilogue
                }
                walk_advance(p, w, stride);
            }
        }
    } else {
        // ===== epilogue warps 2..9 =====
        const int ew = warp - 2;
        const int q = warp & 3;                  // TMEM lane quarter this warp may access: lanes 32 q .. 32 q + 31
        const int h = ew >> 2;                   // column half: columns 64 h .. 64 h + 63
        float* stage = sStage + (size_t)ew * (32 * 33);
        const float nb2 = -p.beta * 1.4426950408889634f, variance = p.variance;
        const bool mirror_mode = (p.uplo == GABO_SYMM_MIRROR);
        const int64_t ldk = p.ldk;
        ItemWalk w{0, 0};
        walk_advance(p, w, first);
        int64_t bcount = 0;
        for (int64_t it = first; it < p.nitems; it += stride) {
            const int64_t T = row_tiles(p, w.bi);
            const int64_t bj0 = w.seg * SEG;
            const int cnt = (int)((T - bj0) < SEG ? (T - bj0) : SEG);
            const int64_t grow_tile = (p.tile_r0 + w.bi) * TT;
            const int64_t grow = grow_tile + q * 32 + lane;             // this thread's row
            const bool rows_in = grow_tile + TT <= p.row1;
            for (int j = 0; j < cnt; ++j, ++bcount) {
                const int as = (int)(bcount % NACC);
                const uint32_t aph = (uint32_t)((bcount / NACC) & 1);
                const int64_t gcol_tile = (bj0 + j) * TT;
                const bool diag_tile = p.symmetric && (gcol_tile == grow_tile);
                const bool mirror = mirror_mode && (gcol_tile <= grow_tile);
                // interior tile: every row and column exists and no entry is special -- stores without predicates
                const bool interior = rows_in && (gcol_tile + TT <= p.n2) && !diag_tile;
                mbar_wait(&bar_acc_full[as], aph);
                tc_fence_after();
                float na = 0.0f;
                if (KIND == GABO_SQDIST_GAUSSIAN) na = p.normA[grow];
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    const int c0 = h * 64 + cc * 32;
                    float v[32];
                    tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TT + c0), v);
                    const int64_t gcol0 = gcol_tile + c0;
                    epilogue_math<KIND>(v, nb2, variance, na, (KIND == GABO_SQDIST_GAUSSIAN) ? p.normB + gcol0 : nullptr);
                    if (diag_tile) {                                          // K(X,X) diagonal is exactly `variance`
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (gcol0 + i == grow) v[i] = variance;
                    }
                    // direct tile: padded transpose in shared memory, then 128 B coalesced row stores
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 32; ++i) stage[lane * 33 + i] = v[i];
                    __syncwarp();
                    const int64_t gc = gcol0 + lane;
                    float* kdir = p.K + (grow_tile + q * 32 - p.row0) * ldk + gc;
                    if (interior) {
#pragma unroll
                        for (int rr = 0; rr < 32; ++rr) __stcs(kdir + rr * ldk, stage[rr * 33 + lane]);
                        if (mirror) {   // register i across the warp = 32 consecutive floats of row (gcol0 + i)
                            float* kmir = p.K + gcol0 * ldk + grow;
#pragma unroll
                            for (int i = 0; i < 32; ++i) __stcs(kmir + i * ldk, v[i]);
                        }
                    } else {
#pragma unroll 4
                        for (int rr = 0; rr < 32; ++rr) {
                            const int64_t gr = grow_tile + q * 32 + rr;
                            const bool keep = !(mirror_mode && diag_tile) || gc <= gr;       // mirrored diagonal tile: lower part only
                            if (gr < p.row1 && gc < p.n2 && keep) __stcs(kdir + rr * ldk, stage[rr * 33 + lane]);
                        }
                        if (mirror && grow < p.row1) {
#pragma unroll 4
                            for (int i = 0; i < 32; ++i) {
                                const int64_t mr = gcol0 + i;
                                if (mr < grow && mr < p.n2) __stcs(p.K + mr * ldk + grow, v[i]);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_acc_empty[as]);
            }
            walk_advance(p, w, stride);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(NACC * TT));
    }
}

template <int KIND>
int launch_tc(const TcParams& p, int64_t grid, size_t smem, cudaStream_t stream) {
    static size_t attr_smem = 0;     // grows monotonically; a race only repeats the call
    if (smem > attr_smem) {
        GABO_CUDA_TRY(cudaFuncSetAttribute(points_gram_f32_tc_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    points_gram_f32_tc_kernel<KIND><<<(unsigned)grid, TC_THREADS, smem, stream>>>(p);
    return 0;
}

inline size_t align256f(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// Tensor-core FP32 path: dim <= 56 (three operand tiles of 2 * kc * 2 KB must fit in shared memory beside the epilogue staging).
bool points_gram_f32_tc_supported(int dim) { return dim >= 1 && dim <= 56; }

size_t points_gram_f32_tc_workspace_bytes(int64_t n1, int64_t n2, int dim) {
    const int kc = ((dim + 7) / 8) * 2;
    const int64_t n1p = ceil_div64(n1, TT) * TT, n2p = ceil_div64(n2, TT) * TT;
    return align256f((size_t)n1p * 2 * kc * 16) + align256f((size_t)n2p * 2 * kc * 16) + align256f((size_t)n1p * 4) +
           align256f((size_t)n2p * 4);
}

int points_gram_f32_tc(int kind, const float* X, int64_t n1, int64_t ldx, const float* X2, int64_t n2, int64_t ldx2, int dim,
                       float beta, float variance, float* K, int64_t ldk, int64_t row0, int64_t row1, int uplo, bool symmetric,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const int kc = ((dim + 7) / 8) * 2;
    const int64_t n1p = ceil_div64(n1, TT) * TT, n2p = ceil_div64(n2, TT) * TT;
    const size_t bytesA = align256f((size_t)n1p * 2 * kc * 16), bytesB = align256f((size_t)n2p * 2 * kc * 16);
    const size_t bytesNa = align256f((size_t)n1p * 4), bytesNb = align256f((size_t)n2p * 4);
    const size_t need = bytesA + bytesNa + (symmetric ? 0 : bytesB + bytesNb);
    if (workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 127) != 0) return -16;
    if (workspace_bytes < need) return -17;
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    float* Ap = reinterpret_cast<float*>(ws);
    float* nA = reinterpret_cast<float*>(ws + bytesA);
    float* Bp = Ap;
    float* nB = nA;
    const bool want_norm = (kind == GABO_SQDIST_GAUSSIAN);
    const int warps = 8;
    pack_tf32_kernel<<<(unsigned)ceil_div64(n1p, warps), warps * 32, 0, stream>>>(X, n1, ldx, dim, kc, Ap, n1p, want_norm ? nA : nullptr);
    GABO_LAUNCH_CHECK();
    if (!symmetric) {
        Bp = reinterpret_cast<float*>(ws + bytesA + bytesNa);
        nB = reinterpret_cast<float*>(ws + bytesA + bytesNa + bytesB);
        pack_tf32_kernel<<<(unsigned)ceil_div64(n2p, warps), warps * 32, 0, stream>>>(X2, n2, ldx2, dim, kc, Bp, n2p, want_norm ? nB : nullptr);
        GABO_LAUNCH_CHECK();
    }
    TcParams p;
    p.Ap = Ap; p.Bp = Bp; p.normA = want_norm ? nA : nullptr; p.normB = want_norm ? nB : nullptr;
    p.K = K; p.ldk = ldk; p.row0 = row0; p.row1 = row1; p.n2 = n2; p.kc = kc; p.kind = kind; p.uplo = uplo;
    p.symmetric = symmetric ? 1 : 0; p.beta = beta; p.variance = variance;
    p.tile_r0 = row0 / TT;
    p.tiles_m = ceil_div64(row1 - row0, TT);
    p.tiles_n = ceil_div64(n2, TT);
    int64_t nitems = 0;
    for (int64_t bi = 0; bi < p.tiles_m; ++bi) {
        const int64_t T = (uplo == GABO_FULL) ? p.tiles_n : (p.tile_r0 + bi + 1);
        nitems += (T + SEG - 1) / SEG;
    }
    p.nitems = nitems;
    if (nitems == 0) return 0;
    const size_t tile_bytes = (size_t)2 * kc * CHUNK_BYTES;
    const size_t smem = (1 + NBST) * tile_bytes + (size_t)8 * 32 * 33 * sizeof(float) + 1024;
    const int64_t grid = nitems < (int64_t)sm_count() ? nitems : (int64_t)sm_count();
    int rc;
    if (kind == GABO_SPHERE_GAUSSIAN) rc = launch_tc<GABO_SPHERE_GAUSSIAN>(p, grid, smem, stream);
    else if (kind == GABO_SPHERE_LAPLACE) rc = launch_tc<GABO_SPHERE_LAPLACE>(p, grid, smem, stream);
    else rc = launch_tc<GABO_SQDIST_GAUSSIAN>(p, grid, smem, stream);
    if (rc) return rc;
    GABO_LAUNCH_CHECK();
    return 0;
}

}  // namespace gabo
