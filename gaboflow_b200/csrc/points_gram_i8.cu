// FP64 sphere Gram with the inner products on the 5th-generation INTEGER tensor cores (sm_100a: tcgen05.mma kind::i8, S32
// accumulators in TMEM), error-free up to a bounded truncation -- the "Ozaki scheme" specialised to this kernel.
//
// Same contract as gabo_points_gram_f64 for the sphere kinds (reference sphere_utils_tf.py:11-60 + kernels_sphere_tf.py:59-88,
// 179-204), FP64 in and out, north_star tolerance 1e-12 relative on the Gram entries.
//
// Why: on B200 the FP64 tensor instruction (DMMA) and DFMA share ONE 64 FMA/clk/SM pipe (profiles/r01_fp64_probe.txt), so in
// points_gram_kernel_1c the 52-term inner product costs 52 of the ~82 FP64-pipe slots per entry; the acos/exp epilogue, the part
// that really needs FP64, is the smaller half.  The integer tensor pipe is a different unit and ~120x faster per MAC.
//
// Arithmetic.  Every coordinate is scaled by one power of two for the whole point set (2^-E with max |x| < 2^E; E = 0 for unit
// vectors) and rounded to a 56-bit fixed-point number v = rint(x 2^(55-E)), split into 8 signed base-128 digits
//     v = sum_{s=0..7} d_s 128^(7-s),   d_1..d_7 in [-64, 63], |d_0| <= 65           (exact: "slices", one int8 plane each),
// so that  x.y 2^(110-2E) = sum_{s,t} 128^(14-s-t) sum_k d_s[k] d'_t[k].  The 36 slice pairs with s + t <= 7 are accumulated,
// grouped by g = s + t, into 8 S32 accumulators D_g (|D_g| <= 8 * 64 * 65^2 < 2^22: exact); the pairs with s + t >= 8 are dropped:
// |error| <= K * 7 * 2^12 * 128^6 * 1.01 / 2^110 < 2^-53 per coordinate of the dot product, i.e. < 6e-15 absolute for K = 52
// (worst case, all digits extreme and aligned; ~1e-15 typical), on top of the 2^-56 representation error of each coordinate.
// The epilogue rebuilds the dot product from the D_g with integer shifts / adds (two 42-bit halves) and 3 FP64 operations, then
// runs the same lean FP64 clamp / acos / exp as the DMMA kernel (gabo_math.cuh).  dK/d(x.y) = K * 2 beta acos / sqrt(1 - t^2), so
// the relative error of a Gaussian Gram entry is <= ~6 beta |dot error| ~ 1e-14 beta.
//
// Data layout: the pack kernel writes every 128-row tile of X as [slice 0..7][k-chunk 0..3][row 0..127][16 int8] -- the K-major,
// non-swizzled UMMA core-matrix layout (8 rows x 16 B contiguous, next 8 rows +128 B, next 16 of K +2048 B) -- so an operand tile
// (64 KB, K padded to 64) is ONE contiguous TMA bulk copy and every shared-memory descriptor is base + constant.
//
// Kernel: persistent CTAs (one per SM, 576 threads).  Warp 0 lane 0: TMA producer (A tile per work item, B tiles through a
// 2-deep ring).  Warp 1 lane 0: MMA issuer -- per 128 x 32 output sub-tile, 8 slices x 2 k-steps = 16 tcgen05.mma (M = 128,
// N = 32 (8 - s): A_s against the STACK B_0 .. B_(7-s), K = 32) into one of two TMEM stages of 8 x 32 columns (the whole
// 512-column TMEM).  Warps 2..17: two sets of 8 epilogue warps, one set per TMEM stage; a warp owns 32 rows (its TMEM lane
// quarter) x 16 columns, in two passes of 8 columns: tcgen05.ld.16x256b of the 8 accumulators (the MMA-fragment layout: thread =
// rows r, r + 8 x columns c, c + 1), integer recombination, FP64 epilogue, then stores straight from the registers -- in that
// layout both the direct tile (4 lanes x 16 B = 64 B of a row) and its mirror image K[col][row] (8 lanes = 8 consecutive rows =
// 64 B of a mirrored row) go out in whole 64-byte segments with no transposition.
// Measurements, diagnostic knobs and the variants that were tried and dropped: DESIGN.md section 4.1b.
#include <cstdio>
#include <cstdlib>
#include "gabo_common.cuh"
#include "gabo_math.cuh"

namespace gabo {
namespace {

constexpr int OT = 128;                      // tile side (UMMA M)
constexpr int OKP = 64;                      // K after padding (dim <= 64)
constexpr int NSL = 8;                       // slices (base-128 digits)
constexpr int OCHUNK = OT * 16;              // one 16-byte k-chunk of all 128 rows
constexpr int OSLICE = (OKP / 16) * OCHUNK;  // 8 KB per slice
constexpr int OTILE = NSL * OSLICE;          // 64 KB per operand tile
constexpr int SUBN = 32;                     // UMMA N
constexpr int NSUB = OT / SUBN;              // sub-tiles per column tile
constexpr int ONACC = 2;                     // TMEM stages
constexpr int STAGE_COLS = NSL * SUBN;       // 256 columns per stage
constexpr int OSEG = 8;                      // column tiles per work item
constexpr int ONBST = 2;                     // B ring depth
constexpr int O_THREADS = 576;               // warp 0 producer, warp 1 MMA, warps 2..17 epilogue
constexpr int BCHUNK = NSL * SUBN * 16;      // B layout: one 16-byte k-chunk of the 8 stacked slices of a sub-tile (4 KB)
constexpr int BSUB = (OKP / 16) * BCHUNK;    // B layout: one sub-tile (16 KB)

struct OzParams {
    const int8_t* Ap;
    const int8_t* Bp;
    const int* expo;       // device: common binary exponents E_A, E_B (max |x| < 2^E)
    double* K;
    int64_t ldk, row0, row1, n2;
    int kind, uplo, symmetric;
    double beta, variance;
    int64_t tile_r0, tiles_m, tiles_n, nitems;
    int debug;                 // diagnostic knobs (GABO_I8_DEBUG): 1 = one MMA per accumulator only, 2 = no acos/exp
    int cyc_world, cyc_rank;   // > 0: local row tile bi is global row tile ((bi / 4) * world + rank) * 4 + bi % 4 (512-row panels)
};

// ---- packing -------------------------------------------------------------------------------------------------------------
// pass 1: E = smallest integer with max |x| < 2^E (0 for an all-zero set); atomicMax on an int initialised to INT_MIN
__global__ void oz_exponent_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int dim, int* __restrict__ expo) {
    int e = -2000;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n * dim; idx += (int64_t)gridDim.x * blockDim.x) {
        const double v = fabs(X[(idx / dim) * ldx + idx % dim]);
        if (v > 0.0 && v < 1.7976931348623157e308) { int ee; frexp(v, &ee); e = max(e, ee); }   // v = m 2^ee, m in [0.5, 1)
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) e = max(e, __shfl_xor_sync(0xffffffffu, e, o));
    if ((threadIdx.x & 31) == 0 && e > -2000) atomicMax(expo, e);
}
__global__ void oz_exponent_init_kernel(int* expo) { expo[0] = -1000; expo[1] = -1000; }
// all-zero (or all non-finite) set: E = 0; symmetric Gram: both operands share one exponent
__global__ void oz_exponent_fix_kernel(int* expo, int symmetric) {
    if (expo[0] == -1000) expo[0] = 0;
    if (symmetric) expo[1] = expo[0];
    else if (expo[1] == -1000) expo[1] = 0;
}

// pass 2: X [n, dim] -> two copies of every 128-row tile (64 KB each), one warp per row, a lane per coordinate:
//   A layout  [slice 0..7][k-chunk 0..3][row 0..127][16 B]              : operand A of MMA (s, kk) = slice s, chunks 2kk, 2kk+1
//   B layout  [sub-tile 0..3][k-chunk 0..3][slice 0..7][row 0..31][16 B]  : for one 32-column sub-tile, the slices are stacked
//             along N (row index t * 32 + j), so ONE MMA multiplies A_s with B_0 .. B_(7-s) at once (see the kernel)
__global__ void oz_pack_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int dim, const int* __restrict__ expo,
                               int8_t* __restrict__ PA, int8_t* __restrict__ PB, int64_t n_pad) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (row >= n_pad) return;
    const int E = expo[0];
    const int64_t tile = row / OT;
    const int r = (int)(row - tile * OT);
    int8_t* baseA = PA != nullptr ? PA + tile * (int64_t)OTILE : nullptr;
    int8_t* baseB = PB != nullptr ? PB + tile * (int64_t)OTILE : nullptr;
    for (int k = lane; k < OKP; k += 32) {
        double x = 0.0;
        if (row < n && k < dim) x = X[row * ldx + k];
        long long v = 0;
        if (x == x && fabs(x) < 1.7976931348623157e308) v = llrint(scalbn(x, 55 - E));   // |v| <= 2^55
        int8_t d[NSL];
#pragma unroll
        for (int s = NSL - 1; s >= 1; --s) {
            const long long dd = ((v + 64) & 127) - 64;     // balanced digit in [-64, 63]
            d[s] = (int8_t)dd;
            v = (v - dd) >> 7;                              // exact: v - dd is a multiple of 128
        }
        d[0] = (int8_t)v;                                   // |v| <= 65
        if (baseA != nullptr) {
            const int64_t off = (int64_t)(k >> 4) * OCHUNK + r * 16 + (k & 15);
#pragma unroll
            for (int s = 0; s < NSL; ++s) baseA[(int64_t)s * OSLICE + off] = d[s];
        }
        if (baseB != nullptr) {
            const int64_t off = (int64_t)(r >> 5) * BSUB + (int64_t)(k >> 4) * BCHUNK + (r & 31) * 16 + (k & 15);
#pragma unroll
            for (int s = 0; s < NSL; ++s) baseB[off + s * (SUBN * 16)] = d[s];
        }
    }
}

// ---- tcgen05 / mbarrier PTX -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void oz_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void oz_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void oz_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void oz_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, M = 128, N = 32, K = 32, signed 8-bit operands, S32 accumulators
__device__ __forceinline__ void oz_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t zero = 0;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(zero)
        : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 B; SBO (next 8 rows) = 128 B, LBO (next 16 B of K) = the chunk stride
__device__ __forceinline__ uint64_t oz_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(128 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// c_format S32 (2) | a_format S8 (1) | b_format S8 (1) | K-major A and B | N >> 3 | M >> 4
constexpr uint32_t kIdescI8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OT >> 4) << 24);      // | (N >> 3) << 17 per MMA

// 16 TMEM lanes x 8 columns: thread t gets rows t/4 and t/4 + 8 (of the 16), columns 2 (t%4) and 2 (t%4) + 1 -- the layout of an
// MMA accumulator fragment, in which BOTH the direct tile (4 lanes x 2 doubles = 64 B of a row) and its mirror image (8 lanes =
// 8 consecutive rows = 64 B of a mirrored row) store in whole 64-byte segments without any transposition.
__device__ __forceinline__ void oz_ld16x8(uint32_t taddr, int (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
}
// producer / MMA-issuer threads wait with a back-off: their try_wait loops otherwise take issue slots from the epilogue warps
__device__ __forceinline__ void oz_mbar_wait_idle(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(64);
    }
}
__device__ __forceinline__ void oz_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// ---- work items (as in points_gram_f32_tc.cu): (row tile, segment of up to OSEG column tiles) ----------------------------------
struct OzWalk { int64_t bi, seg; };
__device__ __forceinline__ int64_t oz_global_tile(const OzParams& p, int64_t bi) {
    return p.cyc_world > 0 ? ((bi >> 2) * p.cyc_world + p.cyc_rank) * 4 + (bi & 3) : p.tile_r0 + bi;
}
__device__ __forceinline__ int64_t oz_row_tiles(const OzParams& p, int64_t bi) {
    return (p.uplo == GABO_FULL) ? p.tiles_n : (oz_global_tile(p, bi) + 1);
}
__device__ __forceinline__ void oz_advance(const OzParams& p, OzWalk& w, int64_t adv) {
    while (w.bi < p.tiles_m) {
        const int64_t left = (oz_row_tiles(p, w.bi) + OSEG - 1) / OSEG - w.seg;
        if (adv < left) { w.seg += adv; return; }
        adv -= left;
        ++w.bi;
        w.seg = 0;
    }
}

template <int KIND>
__global__ void __launch_bounds__(O_THREADS, 1) points_gram_i8_kernel(const OzParams p) {
    extern __shared__ __align__(1024) unsigned char oz_smem[];
    unsigned char* const sA = oz_smem;
    unsigned char* const sB = oz_smem + OTILE;            // ONBST tiles
    __shared__ __align__(8) uint64_t bar_a_full, bar_a_empty, bar_b_full[ONBST], bar_b_empty[ONBST], bar_acc_full[ONACC], bar_acc_empty[ONACC];
    __shared__ uint32_t tmem_base_smem;
    __shared__ double exptab[32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&bar_a_full, 1);
        mbar_init(&bar_a_empty, 1);
        for (int s = 0; s < ONBST; ++s) { mbar_init(&bar_b_full[s], 1); mbar_init(&bar_b_empty[s], 1); }
        for (int s = 0; s < ONACC; ++s) { mbar_init(&bar_acc_full[s], 1); mbar_init(&bar_acc_empty[s], 8); }
        fence_mbar_init();
    }
    if (tid < 32) exptab[tid] = kExp2Tab[tid] * p.variance;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_smem)), "n"(ONACC * STAGE_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    oz_fence_before();
    __syncthreads();
    oz_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const int64_t first = blockIdx.x, stride = gridDim.x;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            OzWalk w{0, 0};
            oz_advance(p, w, first);
            uint32_t a_phase = 0;
            int64_t bcount = 0;
            for (int64_t it = first; it < p.nitems; it += stride) {
                const int64_t T = oz_row_tiles(p, w.bi);
                const int64_t bj0 = w.seg * OSEG;
                const int cnt = (int)((T - bj0) < OSEG ? (T - bj0) : OSEG);
                oz_mbar_wait_idle(&bar_a_empty, a_phase ^ 1);
                mbar_arrive_expect_tx(&bar_a_full, OTILE);
                tma_bulk_g2s(sA, p.Ap + oz_global_tile(p, w.bi) * (int64_t)OTILE, OTILE, &bar_a_full);
                a_phase ^= 1;
                for (int j = 0; j < cnt; ++j, ++bcount) {
                    const int s = (int)(bcount % ONBST);
                    const uint32_t ph = (uint32_t)((bcount / ONBST) & 1);
                    oz_mbar_wait_idle(&bar_b_empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&bar_b_full[s], OTILE);
                    tma_bulk_g2s(sB + (size_t)s * OTILE, p.Bp + (bj0 + j) * (int64_t)OTILE, OTILE, &bar_b_full[s]);
                }
                oz_advance(p, w, stride);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            OzWalk w{0, 0};
            oz_advance(p, w, first);
            uint32_t a_phase = 0;
            int64_t bcount = 0, scount = 0;
            long long dbg_cycles = 0, dbg_n = 0;      // B tiles / sub-tiles issued so far
            const uint32_t a_addr = smem_u32(sA);
            for (int64_t it = first; it < p.nitems; it += stride) {
                const int64_t T = oz_row_tiles(p, w.bi);
                const int64_t bj0 = w.seg * OSEG;
                const int cnt = (int)((T - bj0) < OSEG ? (T - bj0) : OSEG);
                oz_mbar_wait_idle(&bar_a_full, a_phase);
                a_phase ^= 1;
                for (int j = 0; j < cnt; ++j, ++bcount) {
                    const int s = (int)(bcount % ONBST);
                    const uint32_t ph = (uint32_t)((bcount / ONBST) & 1);
                    oz_mbar_wait_idle(&bar_b_full[s], ph);
                    const uint32_t b_addr = smem_u32(sB + (size_t)s * OTILE);
#pragma unroll 1
                    for (int sub = 0; sub < NSUB; ++sub, ++scount) {
                        const int as = (int)(scount & 1);
                        const uint32_t aph = (uint32_t)((scount >> 1) & 1);
                        oz_mbar_wait_idle(&bar_acc_empty[as], aph ^ 1);               // that set of epilogue warps has drained the stage
                        oz_fence_after();
                        // 16 MMAs per 128 x 32 sub-tile: A_s (k-step kk) times the STACK B_0 .. B_(7-s) of the sub-tile's columns
                        // (N = 32 (8 - s)), accumulated into the accumulators g = s .. 7 -- exactly the pairs s + t = g <= 7.
                        // The 72 narrow products of the first version (one per pair and k-step, each re-streaming its A slice) took
                        // 4.5x the tensor time of these 16 wide ones, which run at the MAC rate of the unit (~1150 cycles of MMA
                        // per sub-tile, 1711 from first issue to completion).
                        const uint32_t b_sub = b_addr + (uint32_t)(sub * BSUB);
                        const long long dbg_t0 = clock64();
#pragma unroll 1
                        for (int sa = 0; sa < ((p.debug & 1) ? 1 : NSL); ++sa) {
                            const uint32_t idesc = kIdescI8 | ((uint32_t)(((NSL - sa) * SUBN) >> 3) << 17);
                            const uint32_t d_tmem = tmem_base + (uint32_t)(as * STAGE_COLS + sa * SUBN);
#pragma unroll
                            for (int kk = 0; kk < OKP / 32; ++kk)
                                oz_mma_i8(d_tmem, oz_smem_desc(a_addr + (uint32_t)(sa * OSLICE + kk * 2 * OCHUNK), OCHUNK),
                                          oz_smem_desc(b_sub + (uint32_t)(kk * 2 * BCHUNK), BCHUNK), idesc, (sa | kk) != 0);
                        }
                        oz_commit(&bar_acc_full[as]);
                        if (p.debug & 16) {        // diagnostic: issue -> completion time of one sub-tile's MMAs (serialises them)
                            mbar_wait(&bar_acc_full[as], aph);
                            dbg_cycles += clock64() - dbg_t0;
                            ++dbg_n;
                        }
                    }
                    oz_commit(&bar_b_empty[s]);
                    if (j == cnt - 1) oz_commit(&bar_a_empty);
                }
                oz_advance(p, w, stride);
            }
            if ((p.debug & 16) && blockIdx.x == 0)
                printf("i8 gram: %lld sub-tiles, %.0f cycles from first MMA issue to completion per sub-tile\n", dbg_n, (double)dbg_cycles / (double)(dbg_n > 0 ? dbg_n : 1));
        }
    } else {
        // ===== epilogue: set 0 = warps 2..9 (TMEM stage 0), set 1 = warps 10..17 (stage 1) =====
        const int ew = (warp - 2) & 7;
        const int set = (warp - 2) >> 3;
        const int q = warp & 3;                  // TMEM lane quarter of this warp: lanes 32 q .. 32 q + 31
        const int h = ew >> 2;                   // column half of the 32-column sub-tile
        const double nbeta = -p.beta, variance = p.variance;
        const int exp_slow_hi = exp_slow_threshold_hi(variance);
        const bool mirror_mode = (p.uplo == GABO_SYMM_MIRROR);
        const int64_t ldk = p.ldk, ldk8 = 8 * p.ldk;
        const bool vec_ok = ((ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.K) & 15) == 0);   // 16-byte row-pair stores
        // dot = (H 2^28 + Lo) 2^(E_A + E_B - 110 + 49)
        const double c0 = scalbn(1.0, p.expo[0] + p.expo[1] - 61), c1 = scalbn(1.0, p.expo[0] + p.expo[1] - 33);
        const double kMagic = 6755399441055744.0;            // 2^52 + 2^51
        const double lo_bias = -kMagic * c0;
        OzWalk w{0, 0};
        oz_advance(p, w, first);
        int64_t scount = 0;
        for (int64_t it = first; it < p.nitems; it += stride) {
            const int64_t T = oz_row_tiles(p, w.bi);
            const int64_t bj0 = w.seg * OSEG;
            const int cnt = (int)((T - bj0) < OSEG ? (T - bj0) : OSEG);
            const int64_t grow_tile = oz_global_tile(p, w.bi) * OT;
            const int64_t lrow_tile = p.cyc_world > 0 ? w.bi * OT : grow_tile - p.row0;   // row of K (local buffer)
            for (int j = 0; j < cnt; ++j) {
                const int64_t gcol_tile = (bj0 + j) * OT;
                const bool diag_tile = p.symmetric && (gcol_tile == grow_tile);
                const bool mirror = mirror_mode && (gcol_tile < grow_tile);
                for (int sub = 0; sub < NSUB; ++sub, ++scount) {
                    if ((int)(scount & 1) != set) continue;
                    const uint32_t aph = (uint32_t)((scount >> 1) & 1);
                    mbar_wait(&bar_acc_full[set], aph);
                    oz_fence_after();
                    const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * STAGE_COLS + h * 16);
                    const bool rows_all = grow_tile + OT <= p.row1;
                    // fast store path: whole tile inside the matrix, no diagonal entries to overwrite (those tiles take the generic
                    // path below); everything that does not depend on the pass is computed once per sub-tile
                    const bool interior = rows_all && (gcol_tile + OT <= p.n2) && !diag_tile && vec_ok;
                    const int lr = lane >> 2, lc = (lane & 3) * 2;      // fragment coordinates: rows lr + 8 rg, columns lc, lc + 1
                    const int64_t gc_sub = gcol_tile + sub * SUBN + h * 16 + lc;      // this thread's first column of pass 0
                    const int64_t gr0 = grow_tile + q * 32 + lr;                      // ... and first row (then + 8 rg)
                    double* kdir_sub = p.K + (lrow_tile + q * 32 + lr) * ldk + gc_sub;
                    double* kmir_sub = p.K + gc_sub * ldk + gr0;
                    unsigned store_mode = (interior ? 1u : 0u) | (mirror ? 2u : 0u);
                    // opaque to the compiler, which otherwise re-derives all three (64-bit multiplies, ~50 instructions) in every pass
                    asm volatile("" : "+l"(kdir_sub), "+l"(kmir_sub), "+r"(store_mode));
#pragma unroll 1
                    for (int pass = 0; pass < 2; ++pass) {
                        // v[2 rg + e] = entry (row 32 q + 8 rg + lr, column gcol0 + lc + e)
                        const uint32_t ta = tbase + pass * 8, tb = ta + (16u << 16);
                        int d0[8], d1[8], d2[8], d3[8];
#define OZ_LD_GROUP(G, D)                                                          \
    do {                                                                           \
        int lo_[4], hi_[4];                                                        \
        oz_ld16x8(ta + (G) * SUBN, lo_);                                           \
        oz_ld16x8(tb + (G) * SUBN, hi_);                                           \
        D[0] = lo_[0]; D[1] = lo_[1]; D[2] = lo_[2]; D[3] = lo_[3];                \
        D[4] = hi_[0]; D[5] = hi_[1]; D[6] = hi_[2]; D[7] = hi_[3];                \
    } while (0)
                        OZ_LD_GROUP(0, d0); OZ_LD_GROUP(1, d1); OZ_LD_GROUP(2, d2); OZ_LD_GROUP(3, d3);
                        oz_ld_wait();
                        double v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int p01 = d0[i] * 128 + d1[i], p23 = d2[i] * 128 + d3[i];         // < 2^30
                            const long long H = (long long)p01 * 16384 + (long long)p23;            // < 2^44
                            v[i] = __hiloint2double((int)(H >> 32) + 0x43380000, (int)(H & 0xffffffffll)) - kMagic;   // exact
                        }
                        OZ_LD_GROUP(4, d0); OZ_LD_GROUP(5, d1); OZ_LD_GROUP(6, d2); OZ_LD_GROUP(7, d3);
#undef OZ_LD_GROUP
                        oz_ld_wait();
                        if (pass == 1) {            // every accumulator of this stage is in registers: hand the stage back
                            oz_fence_before();
                            __syncwarp();
                            if (lane == 0) oz_mbar_arrive(&bar_acc_empty[set]);
                        }
                        {
                            double dot[8], dd[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int p45 = d0[i] * 128 + d1[i], p67 = d2[i] * 128 + d3[i];
                                const long long L = (long long)p45 * 16384 + (long long)p67;
                                const double lraw = __hiloint2double((int)(L >> 32) + 0x43380000, (int)(L & 0xffffffffll));   // 2^52 + 2^51 + L
                                dot[i] = fma(v[i], c1, fma(lraw, c0, lo_bias));
                            }
                            if (p.debug & 2) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = dot[i];
                            } else {
                            acos_clamped_fast_n<8>(dot, dd);                                                   // sphere_utils_tf.py:59-60
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                dot[i] = (KIND == GABO_SPHERE_GAUSSIAN) ? (nbeta * dd[i]) * dd[i] : nbeta * dd[i];   // kernels_sphere_tf.py:83-86, 199-202
                            exp_neg_scaled_fast_n<8>(dot, v, exptab, variance, exp_slow_hi);                    // kernels_sphere_tf.py:88, 204
                            }
                        }
                        double* const kdir = kdir_sub + pass * 8;
                        if (store_mode & 1u) {
                            double* kd = kdir;
#pragma unroll
                            for (int rg = 0; rg < 4; ++rg, kd += ldk8) st_cs_v2(kd, v[2 * rg], v[2 * rg + 1]);
                            if (store_mode & 2u) {     // K[col][row]: the 8 lanes of equal lc write 8 consecutive rows = 64 B
                                double* const kmir = pass ? kmir_sub + ldk8 : kmir_sub;
                                double* const kmir1 = kmir + ldk;
#pragma unroll
                                for (int rg = 0; rg < 4; ++rg) {
                                    st_cs(kmir + 8 * rg, v[2 * rg]);
                                    st_cs(kmir1 + 8 * rg, v[2 * rg + 1]);
                                }
                            }
                        } else {
                            const int64_t gc = gc_sub + pass * 8;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int64_t gr = gr0 + 8 * (i >> 1), gcc = gc + (i & 1);
                                const double val = (diag_tile && gcc == gr) ? variance : v[i];   // K(X,X) diagonal is exactly `variance`
                                const bool keep = !(mirror_mode && diag_tile) || gcc <= gr;   // mirrored diagonal tile: lower part + diagonal
                                if (gr < p.row1 && gcc < p.n2 && keep) st_cs(kdir + (int64_t)(8 * (i >> 1)) * ldk + (i & 1), val);
                                // mirror image (for the diagonal tile: its upper part as the transpose of the lower one)
                                if (mirror_mode && gr < p.row1 && gcc < gr && (mirror || diag_tile)) st_cs(p.K + gcc * ldk + gr, val);
                            }
                        }
                    }
                }
            }
            oz_advance(p, w, stride);
        }
    }

    oz_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(ONACC * STAGE_COLS));
    }
}

template <int KIND>
int oz_launch(const OzParams& p, int64_t grid, size_t smem, cudaStream_t stream) {
    GABO_CUDA_TRY(cudaFuncSetAttribute(points_gram_i8_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    points_gram_i8_kernel<KIND><<<(unsigned)grid, O_THREADS, smem, stream>>>(p);
    return 0;
}

inline size_t oz_align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

bool points_gram_i8_supported(int kind, int dim) {
    return (kind == GABO_SPHERE_GAUSSIAN || kind == GABO_SPHERE_LAPLACE) && dim >= 1 && dim <= OKP;
}

size_t points_gram_i8_workspace_bytes(int64_t n1, int64_t n2) {
    const int64_t n1p = ceil_div64(n1, OT) * OT, n2p = ceil_div64(n2, OT) * OT;
    return oz_align256((size_t)n1p * OKP * NSL) + oz_align256((size_t)n2p * OKP * NSL) + 256;   // A layout of X + B layout of X2
}

// cyc_world > 0: block-cyclic row sharding (512-row panels of this rank, K = the rank's local [local_rows, ldk] buffer, FULL only)
int points_gram_i8(int kind, const double* X, int64_t n1, int64_t ldx, const double* X2, int64_t n2, int64_t ldx2, int dim, double beta,
                   double variance, double* K, int64_t ldk, int64_t row0, int64_t row1, int uplo, bool symmetric, int cyc_world,
                   int cyc_rank, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const int64_t n1p = ceil_div64(n1, OT) * OT, n2p = ceil_div64(n2, OT) * OT;
    const size_t bytesA = oz_align256((size_t)n1p * OKP * NSL), bytesB = oz_align256((size_t)n2p * OKP * NSL);
    const size_t need = bytesA + bytesB + 256;
    if (workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 127) != 0) return -16;
    if (workspace_bytes < need) return -17;
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    int8_t* Ap = reinterpret_cast<int8_t*>(ws);                  // X in the A layout
    int8_t* Bp = reinterpret_cast<int8_t*>(ws + bytesA);         // X2 (or X again) in the B layout
    int* expo = reinterpret_cast<int*>(ws + bytesA + bytesB);
    const int warps = 8;
    oz_exponent_init_kernel<<<1, 1, 0, stream>>>(expo);
    GABO_LAUNCH_CHECK();
    oz_exponent_kernel<<<(unsigned)(sm_count() * 4), 256, 0, stream>>>(X, n1, ldx, dim, expo);
    GABO_LAUNCH_CHECK();
    if (!symmetric) {
        oz_exponent_kernel<<<(unsigned)(sm_count() * 4), 256, 0, stream>>>(X2, n2, ldx2, dim, expo + 1);
        GABO_LAUNCH_CHECK();
    }
    oz_exponent_fix_kernel<<<1, 1, 0, stream>>>(expo, symmetric ? 1 : 0);
    GABO_LAUNCH_CHECK();
    oz_pack_kernel<<<(unsigned)ceil_div64(n1p, warps), warps * 32, 0, stream>>>(X, n1, ldx, dim, expo, Ap, symmetric ? Bp : nullptr, n1p);
    GABO_LAUNCH_CHECK();
    if (!symmetric) {
        oz_pack_kernel<<<(unsigned)ceil_div64(n2p, warps), warps * 32, 0, stream>>>(X2, n2, ldx2, dim, expo + 1, nullptr, Bp, n2p);
        GABO_LAUNCH_CHECK();
    }
    OzParams p;
    p.Ap = Ap; p.Bp = Bp; p.expo = expo; p.K = K; p.ldk = ldk; p.row0 = row0; p.row1 = row1; p.n2 = n2;
    p.kind = kind; p.uplo = uplo; p.symmetric = symmetric ? 1 : 0; p.beta = beta; p.variance = variance;
    p.cyc_world = cyc_world; p.cyc_rank = cyc_rank;
    { const char* e = getenv("GABO_I8_DEBUG"); p.debug = e ? atoi(e) : 0; }
    p.tile_r0 = row0 / OT;
    p.tiles_m = ceil_div64(row1 - row0, OT);
    p.tiles_n = ceil_div64(n2, OT);
    if (cyc_world > 0) {             // rows of this rank's 512-row panels
        const int64_t npanels = ceil_div64(n1, 512);
        const int64_t nslots = npanels > cyc_rank ? (npanels - 1 - cyc_rank) / cyc_world + 1 : 0;
        p.tile_r0 = 0;
        p.tiles_m = nslots * 4;
        p.row0 = 0; p.row1 = n1;
    }
    int64_t nitems = 0;
    for (int64_t bi = 0; bi < p.tiles_m; ++bi) {
        const int64_t gt = cyc_world > 0 ? ((bi >> 2) * cyc_world + cyc_rank) * 4 + (bi & 3) : p.tile_r0 + bi;
        const int64_t T = (uplo == GABO_FULL) ? p.tiles_n : (gt + 1);
        nitems += (T + OSEG - 1) / OSEG;
    }
    p.nitems = nitems;
    if (nitems == 0) return 0;
    const size_t smem = (size_t)(1 + ONBST) * OTILE + 1024;
    const int64_t grid = nitems < (int64_t)sm_count() ? nitems : (int64_t)sm_count();
    int rc;
    if (kind == GABO_SPHERE_GAUSSIAN) rc = oz_launch<GABO_SPHERE_GAUSSIAN>(p, grid, smem, stream);
    else rc = oz_launch<GABO_SPHERE_LAPLACE>(p, grid, smem, stream);
    if (rc) return rc;
    GABO_LAUNCH_CHECK();
    return 0;
}

}  // namespace gabo

extern "C" size_t gabo_points_gram_i8_workspace_bytes(int64_t n1, int64_t n2) {
    if (n1 < 0 || n2 < 0) return 0;
    return gabo::points_gram_i8_workspace_bytes(n1, n2);
}

extern "C" int gabo_points_gram_i8_f64(int kind, const double* X, int64_t n1, int64_t ldx, const double* X2, int64_t n2, int64_t ldx2,
                                       int dim, double beta, double variance, double* K, int64_t ldk, int64_t row0, int64_t row1,
                                       int uplo, int cyc_world, int cyc_rank, void* workspace, size_t workspace_bytes,
                                       gabo_stream_t stream) {
    if (!gabo::points_gram_i8_supported(kind, dim)) return (kind == GABO_SPHERE_GAUSSIAN || kind == GABO_SPHERE_LAPLACE) ? -8 : -1;
    if (n1 < 0) return -3;
    if (n1 == 0) return 0;
    if (X == nullptr) return -2;
    if (ldx < dim) return -4;
    const bool symmetric = (X2 == nullptr);
    if (symmetric) { n2 = n1; ldx2 = ldx; }
    if (n2 < 0) return -6;
    if (ldx2 < dim) return -7;
    if (K == nullptr) return -11;
    if (ldk < n2) return -12;
    if (uplo < GABO_FULL || uplo > GABO_SYMM_MIRROR) return -15;
    if (cyc_world < 0 || cyc_world > GABO_MAX_PEERS) return -16;
    if (cyc_world > 0) {
        if (cyc_rank < 0 || cyc_rank >= cyc_world) return -17;
        if (uplo != GABO_FULL || n1 % 512 != 0) return -15;
        row0 = 0; row1 = n1;
    }
    if (row0 < 0 || row0 % 128 != 0 || row0 > n1) return -13;
    if (row1 < row0 || row1 > n1) return -14;
    if (uplo != GABO_FULL && !symmetric) return -15;
    if (uplo == GABO_SYMM_MIRROR && !(row0 == 0 && row1 == n1)) return -15;
    if (row1 == row0 || n2 == 0) return 0;
    const int rc = gabo::points_gram_i8(kind, X, n1, ldx, X2, n2, ldx2, dim, beta, variance, K, ldk, row0, row1, uplo, symmetric,
                                        cyc_world, cyc_rank, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
    return (rc == -16) ? -18 : (rc == -17) ? -19 : rc;
}
