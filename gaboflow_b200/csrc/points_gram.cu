// Fused point-set Gram tile kernel (sm_100a): K = f(X . X2^T) for the sphere Gaussian / Laplace kernels and
// the squared-distance Gaussian kernel, one pass, no N1 x N2 x D intermediates.
//
// Replaces  sphere_distance_tf (reference sphere_utils_tf.py:11-60: tile/tile/batched-matmul/clamp/acos)
//           + SphereGaussianKernel.K / SphereLaplaceKernel.K (kernels_sphere_tf.py:59-88, 179-204)
//           + the pair stage of SpdFrobeniusGaussianKernel.K / SpdLogEuclideanGaussianKernel.K
//             (kernels_spd_tf.py:510-533, 618-641).
//
// Design (B200):
//   * points are re-packed once per call into panel layout  P[chunk][row][KLD]  (zero padded to 128 rows and KLD
//     columns), so that the 128-row operand tile of a chunk is ONE contiguous 128*KLD*8-byte block that a single
//     TMA bulk copy (cp.async.bulk, SASS UBLKCP) lands in shared memory; KLD = 4 (mod 16) makes every DMMA
//     fragment load bank-conflict free without swizzling;
//   * persistent CTAs (one per SM, 16 warps), static round-robin over 128x128 output tiles, two-stage shared
//     memory ring: the bulk copies of work item i+1 are in flight while item i is computed;
//   * inner products on the FP64 tensor pipe (mma.sync m8n8k4 -> DMMA.8x8x4), 32x32 outputs per warp held in
//     registers; the epilogue (clamp, acos, square, scale, exp, variance) runs on those registers and stores
//     with streaming 128-bit writes; the symmetric mode computes lower tiles only and mirror-stores them
//     (both stores fill whole 32-byte sectors: 64 B per quad row / per 8-lane column group).
#include <mutex>
#include <vector>
#include "gabo_common.cuh"
#include "gabo_math.cuh"

namespace gabo {
namespace {

constexpr int BT = 128;         // output tile side
constexpr int NTHREADS = 512;   // 16 warps, 4 x 4, 32x32 outputs each
constexpr int NSTAGE = 2;

struct StageEnt { int64_t off; int32_t first_slot; int32_t nrows; };   // staging block of (chunk, destination rank)

struct GramParams {
    const double* Ap;      // packed X   [nchunks][n1_pad][KLD]
    const double* Bp;      // packed X2  [nchunks][n2_pad][KLD]
    const double* normA;   // squared norms (sqdist kind) or nullptr
    const double* normB;
    double* K;
    int64_t ldk;
    int64_t n1_pad, n2_pad;
    int64_t row0, row1;    // global row range produced; K points at row row0
    int64_t n2;
    int nchunks;
    int kind;
    int uplo;
    int symmetric;
    int vec_ok;
    double beta, variance;
    int64_t tile_r0;       // row0 / BT
    int64_t tiles_m, tiles_n, ntiles;
    // block-cyclic multi-GPU mode (points_gram_kernel_1c<KLD, true>): row panels of CYC_PT tiles, panel P lives on rank
    // P % cyc_world in local slot P / cyc_world; peer[r] = base of rank r's local [local_rows, ldk] buffer (peer[cyc_rank]
    // is this GPU's own buffer, the others are NVLink peer mappings)
    int cyc_world, cyc_rank;
    double* peer[8];
    // hybrid mode: a panel pair (PI > PJ) owned by DIFFERENT ranks and selected by cyclic_recompute() is NOT mirrored over
    // NVLink; the owner of row panel PJ recomputes that upper block itself ("extra" tiles, enumerated after the
    // cyc_lower_tiles lower tiles; cyc_extra lists their (local slot, PI) pairs).  cyc_frac == 0: everything is mirrored.
    int cyc_frac;
    int64_t cyc_lower_tiles;
    const int32_t* cyc_extra;
    // staged mode (points_gram_kernel_1c<KLD, 2>): mirror images bound for OTHER ranks are written, transposed, into a local
    // staging buffer laid out so that everything a chunk of row tiles owes one peer is a single dense 2-D block; per-chunk
    // arrival counters raise a flag when the chunk is complete and copy-engine transfers (waiting on that flag on their own
    // streams) move the block over NVLink while the kernel computes on -- the SMs never issue a remote store.
    int cyc_nslots, cyc_split;      // local panel slots; the last cyc_split slots are chunked per row tile, the others per panel
    double* stage;
    const StageEnt* stage_tab;      // [chunk][8]
    uint32_t* prog_cnt;             // [chunks] arrival counters (zeroed before the launch)
    uint32_t* prog_flag;            // [chunks] completion flags (value = epoch)
    uint32_t epoch;
};

constexpr int CYC_PT = 4;   // tiles per 512-row panel

// Hybrid mode: is the upper block (PJ rows, PI columns), PI > PJ, recomputed by the owner of PJ instead of mirrored to it?
// Only pairs owned by different ranks qualify (a same-rank mirror never touches NVLink); a hash of the pair selects
// frac256 / 256 of them, uniformly over ranks and panels.
__host__ __device__ __forceinline__ bool cyclic_recompute(int64_t PI, int64_t PJ, int world, int frac256) {
    if (frac256 <= 0 || PI % world == PJ % world) return false;
    uint32_t h = (uint32_t)PI * 0x9E3779B1u ^ (uint32_t)PJ * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return (int)(h & 255u) < frac256;
}

// Staged mode: the recomputed pairs of (row panel PI, destination rank rJ) are the LOWEST slots of rJ below PI, so that the
// mirrored remainder is one contiguous row range on both sides of the copy: slots [first, count) with
// count = slots of rJ whose panel precedes PI, first = round(count * frac256 / 256).
__host__ __device__ __forceinline__ int64_t staged_slot_count(int64_t PI, int rJ, int world) {
    return PI / world + ((rJ < (int)(PI % world)) ? 1 : 0);
}
__host__ __device__ __forceinline__ int64_t staged_first_slot(int64_t PI, int rJ, int world, int frac256) {
    return (staged_slot_count(PI, rJ, world) * frac256 + 128) >> 8;
}
__host__ __device__ __forceinline__ bool staged_recompute(int64_t PI, int64_t PJ, int world, int frac256) {
    if (PI % world == PJ % world) return false;
    return PJ / world < staged_first_slot(PI, (int)(PJ % world), world, frac256);
}
template <int CYC>
__host__ __device__ __forceinline__ bool cyclic_recompute_m(int64_t PI, int64_t PJ, int world, int frac256) {
    return CYC == 2 ? staged_recompute(PI, PJ, world, frac256) : cyclic_recompute(PI, PJ, world, frac256);
}
// chunk of local row tile lt (staged mode) and the number of lower tiles it holds
__device__ __forceinline__ int staged_chunk(const GramParams& p, int64_t lt, int64_t gbi, uint32_t& total, int& col_off,
                                            int& width) {
    const int64_t s = lt / CYC_PT;
    const int rr = (int)(lt % CYC_PT);
    const int64_t whole = p.cyc_nslots - p.cyc_split;
    const int64_t g0 = gbi - rr;                      // first global row tile of the panel
    if (s < whole) { total = (uint32_t)(4 * g0 + 10); col_off = rr * BT; width = CYC_PT * BT; return (int)s; }
    total = (uint32_t)(gbi + 1); col_off = 0; width = BT;
    return (int)(whole + (s - whole) * CYC_PT + rr);
}

// u-th tile owned by this rank -> local row tile lt, global row tile gbi, column tile bj; extra = recomputed upper tile.
// Lower tiles before local slot s: C(s) = 8 W s (s-1) + (16 rank + 10) s   (rows of global panel s W + rank own 4P+1 .. 4P+4 tiles).
__device__ __forceinline__ void cyclic_tile_coords(const GramParams& p, int64_t u, int64_t& lt, int64_t& gbi, int64_t& bj,
                                                   bool& extra) {
    extra = u >= p.cyc_lower_tiles;
    if (extra) {
        const int64_t e = u - p.cyc_lower_tiles;
        const int64_t s = p.cyc_extra[2 * (e / 16)], PI = p.cyc_extra[2 * (e / 16) + 1];
        const int64_t rr = (e % 16) / 4, cc = e % 4;
        lt = s * CYC_PT + rr;
        gbi = (s * p.cyc_world + p.cyc_rank) * CYC_PT + rr;
        bj = PI * CYC_PT + cc;
        return;
    }
    const double W = (double)p.cyc_world, b = 16.0 * p.cyc_rank + 10.0 - 8.0 * W;
    int64_t s = (int64_t)floor((-b + sqrt(b * b + 32.0 * W * (double)u)) / (16.0 * W));
    if (s < 0) s = 0;
    auto C = [&](int64_t x) { return 8 * (int64_t)p.cyc_world * x * (x - 1) + (16 * (int64_t)p.cyc_rank + 10) * x; };
    while (C(s) > u) --s;
    while (C(s + 1) <= u) ++s;
    int64_t rem = u - C(s);
    const int64_t g0 = (s * p.cyc_world + p.cyc_rank) * CYC_PT;   // first global row tile of the panel
    int rr = 0;
    while (rem >= g0 + rr + 1) { rem -= g0 + rr + 1; ++rr; }
    lt = s * CYC_PT + rr;
    gbi = g0 + rr;
    bj = rem;
}

// ---- packing -----------------------------------------------------------------------------------------
template <int KLD, typename Tin>
__global__ void pack_points_kernel(const Tin* __restrict__ X, int64_t n, int64_t ldx, int dim, double* __restrict__ P,
                                   int64_t n_pad, int nchunks, double* __restrict__ norms) {
    // one warp per row: coalesced reads along the row, writes chunk by chunk
    int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x & 31;
    if (row >= n_pad) return;
    double s = 0.0;
    int width = nchunks * KLD;
    for (int k = lane; k < width; k += 32) {
        double v = 0.0;
        if (row < n && k < dim) v = (double)X[row * ldx + k];
        s = fma(v, v, s);
        int c = k / KLD, kk = k - c * KLD;
        P[((int64_t)c * n_pad + row) * KLD + kk] = v;
    }
    if (norms != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) norms[row] = s;
    }
}

// ---- epilogue math -----------------------------------------------------------------------------------
__device__ __forceinline__ double gram_entry(int kind, double acc, double na, double nb, double beta, double variance,
                                             const double* __restrict__ exptab) {
    double arg;
    if (kind == GABO_SQDIST_GAUSSIAN) {
        arg = fmax(fma(-2.0, acc, na + nb), 0.0);                 // ||a-b||^2 = |a|^2 + |b|^2 - 2 a.b
    } else {
        const double t = fmin(fmax(acc, -1.0), 1.0);              // sphere_utils_tf.py:59: clamp to [-1,1]
        const double d = acos_fast(t);                            // sphere_utils_tf.py:60
        arg = (kind == GABO_SPHERE_GAUSSIAN) ? d * d : d;         // kernels_sphere_tf.py:83-86 / :199-202
    }
    return variance * exp_neg_fast(-beta * arg, exptab);          // kernels_sphere_tf.py:88 / :204
}

__device__ __forceinline__ void tile_coords(const GramParams& p, int64_t t, int64_t& bi, int64_t& bj) {
    if (p.uplo == GABO_FULL) {
        bi = t / p.tiles_n;
        bj = t - bi * p.tiles_n;
    } else {
        // row tile bi (local) owns column tiles 0 .. tile_r0 + bi; tiles before it: bi*tile_r0 + bi(bi+1)/2
        double g = (double)p.tile_r0 + 0.5;
        int64_t b = (int64_t)floor(-g + sqrt(g * g + 2.0 * (double)t));
        if (b < 0) b = 0;
        while (b * p.tile_r0 + b * (b + 1) / 2 > t) --b;
        while ((b + 1) * p.tile_r0 + (b + 1) * (b + 2) / 2 <= t) ++b;
        bi = b;
        bj = t - (b * p.tile_r0 + b * (b + 1) / 2);
    }
}

template <int KLD>
__global__ void __launch_bounds__(NTHREADS, 1) points_gram_kernel(const GramParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr uint32_t TILE_BYTES = BT * KLD * sizeof(double);
    constexpr int TILE_DOUBLES = BT * KLD;
    double* const smem_d = reinterpret_cast<double*>(smem_raw);   // stage s: A tile at 2*s*TILE_DOUBLES, B tile right after
    __shared__ __align__(8) uint64_t full_bar[NSTAGE];
    __shared__ double exptab[32];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    if (tid < 32) exptab[tid] = kExp2Tab[tid];
    __syncthreads();

    // work items: (tile, chunk) pairs owned by this CTA, linearised
    const int64_t my_tiles = (p.ntiles > (int64_t)blockIdx.x) ? (p.ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int64_t nitems = my_tiles * p.nchunks;

    auto issue = [&](int64_t item) {
        int64_t tl = item / p.nchunks;
        int c = (int)(item - tl * p.nchunks);
        int64_t t = (int64_t)blockIdx.x + tl * gridDim.x;
        int64_t bi, bj;
        tile_coords(p, t, bi, bj);
        int s = (int)(item & 1);
        const double* srcA = p.Ap + ((int64_t)c * p.n1_pad + (p.tile_r0 + bi) * BT) * KLD;
        const double* srcB = p.Bp + ((int64_t)c * p.n2_pad + bj * BT) * KLD;
        mbar_arrive_expect_tx(&full_bar[s], 2 * TILE_BYTES);
        tma_bulk_g2s(smem_d + 2 * s * TILE_DOUBLES, srcA, TILE_BYTES, &full_bar[s]);
        tma_bulk_g2s(smem_d + (2 * s + 1) * TILE_DOUBLES, srcB, TILE_BYTES, &full_bar[s]);
    };

    if (tid == 0 && nitems > 0) issue(0);

    double acc[32];   // [mi][ni][e] flattened: the epilogue rotates it so that a rolled loop can index it statically
    for (int64_t item = 0; item < nitems; ++item) {
        const int64_t tl = item / p.nchunks;
        const int c = (int)(item - tl * p.nchunks);
        const int s = (int)(item & 1);
        // every warp is done reading stage s^1 (item-1): refill it with item+1 while this item computes
        __syncthreads();
        if (tid == 0 && item + 1 < nitems) issue(item + 1);
        if (c == 0) {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) acc[(mi * 4 + ni) * 2] = acc[(mi * 4 + ni) * 2 + 1] = 0.0;
        }
        mbar_wait(&full_bar[s], (uint32_t)((item >> 1) & 1));

        const double* a_base = smem_d + 2 * s * TILE_DOUBLES + (wm * 32 + g) * KLD + q;
        const double* b_base = smem_d + (2 * s + 1) * TILE_DOUBLES + (wn * 32 + g) * KLD + q;
#pragma unroll
        for (int ks = 0; ks < KLD / 4; ++ks) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = a_base[mi * 8 * KLD + ks * 4];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = b_base[ni * 8 * KLD + ks * 4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[(mi * 4 + ni) * 2], acc[(mi * 4 + ni) * 2 + 1], a[mi], b[ni]);
        }
        if (c != p.nchunks - 1) continue;

        // ---- epilogue for this tile ----
        const int64_t t = (int64_t)blockIdx.x + tl * gridDim.x;
        int64_t bi, bj;
        tile_coords(p, t, bi, bj);
        const int64_t grow_tile = (p.tile_r0 + bi) * BT;   // global row of the tile
        const int64_t gcol_tile = bj * BT;
        const bool mirror = (p.uplo == GABO_SYMM_MIRROR) && (gcol_tile != grow_tile);
        const bool diag_tile = p.symmetric && (gcol_tile == grow_tile);
        // Rolled loop (8 trips x 4 entries) so the epilogue code stays small enough for the instruction cache; the
        // accumulator array is rotated by 4 each trip, which keeps every register index static.
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
            const int mi = it >> 1, nb2 = (it & 1) * 2;
            const int64_t grow = grow_tile + wm * 32 + mi * 8 + g;
            const bool row_ok = grow < p.row1;
            double na = 0.0;
            if (p.kind == GABO_SQDIST_GAUSSIAN) na = p.normA[grow];
            double* krow = p.K + (grow - p.row0) * p.ldk;
            double v[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t gcol = gcol_tile + wn * 32 + (nb2 + h) * 8 + 2 * q;
                double nb0 = 0.0, nb1 = 0.0;
                if (p.kind == GABO_SQDIST_GAUSSIAN) { nb0 = p.normB[gcol]; nb1 = p.normB[gcol + 1]; }
                v[2 * h] = gram_entry(p.kind, acc[2 * h], na, nb0, p.beta, p.variance, exptab);
                v[2 * h + 1] = gram_entry(p.kind, acc[2 * h + 1], na, nb1, p.beta, p.variance, exptab);
                if (diag_tile) {   // K(X,X) diagonal is exactly `variance` (matches Kdiag, kernels_sphere_tf.py:105)
                    if (gcol == grow) v[2 * h] = p.variance;
                    if (gcol + 1 == grow) v[2 * h + 1] = p.variance;
                }
            }
#pragma unroll
            for (int k = 0; k < 28; ++k) acc[k] = acc[k + 4];
            if (row_ok) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t gcol = gcol_tile + wn * 32 + (nb2 + h) * 8 + 2 * q;
                    if (gcol + 1 < p.n2 && p.vec_ok) {
                        st_cs_v2(krow + gcol, v[2 * h], v[2 * h + 1]);
                    } else {
                        if (gcol < p.n2) st_cs(krow + gcol, v[2 * h]);
                        if (gcol + 1 < p.n2) st_cs(krow + gcol + 1, v[2 * h + 1]);
                    }
                    if (mirror) {   // K[col][row] = K[row][col]; 8 lanes of equal q write 64 contiguous bytes
                        st_cs(p.K + gcol * p.ldk + grow, v[2 * h]);
                        st_cs(p.K + (gcol + 1) * p.ldk + grow, v[2 * h + 1]);
                    }
                }
            }
        }
    }
}

// ---- single-chunk fast path (dim <= KLD, i.e. every configuration of BASELINE.json) -----------------------------------------
// Same tiling, ring and stores as points_gram_kernel, but each warp walks its 32x32 block one 8-row strip at a time in a
// ROLLED loop: 4 DMMA accumulators live at once instead of 16, so the epilogue needs no register rotation, indexes nothing
// dynamically and has 8 independent entries in flight; the B fragments are re-read from shared memory per strip
// (260 instead of 104 LDS per warp and tile, ~20 % of the shared-memory bandwidth).
template <int KLD, int CYC>
__global__ void __launch_bounds__(NTHREADS, 1) points_gram_kernel_1c(const GramParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr uint32_t TILE_BYTES = BT * KLD * sizeof(double);
    constexpr int TILE_DOUBLES = BT * KLD;
    double* const smem_d = reinterpret_cast<double*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[NSTAGE];
    __shared__ double exptab[32];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    // kernel parameters used in the hot loop, once into registers (they were re-read with LDC inside the loop)
    const double nbeta = -p.beta, variance = p.variance;
    const int kind = p.kind;
    const int64_t n2 = p.n2, ldk = p.ldk, row1 = p.row1;
    const bool vec_ok = p.vec_ok != 0;
    double* const Kbase = p.K;
    if (tid < 32) exptab[tid] = kExp2Tab[tid] * variance;   // variance folded into the exp table: one DMUL less per entry
    const int exp_slow_hi = exp_slow_threshold_hi(variance);   // beyond it variance * e^x leaves the normal range: libm path
    __syncthreads();

    const int64_t my_tiles = (p.ntiles > (int64_t)blockIdx.x) ? (p.ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    auto issue = [&](int64_t tl) {
        const int64_t t = (int64_t)blockIdx.x + tl * gridDim.x;
        int64_t bi, bj, gbi;
        bool extra = false;
        if (CYC) cyclic_tile_coords(p, t, bi, gbi, bj, extra);
        else { tile_coords(p, t, bi, bj); gbi = p.tile_r0 + bi; }
        const int s = (int)(tl & 1);
        mbar_arrive_expect_tx(&full_bar[s], 2 * TILE_BYTES);
        tma_bulk_g2s(smem_d + 2 * s * TILE_DOUBLES, p.Ap + gbi * BT * KLD, TILE_BYTES, &full_bar[s]);
        tma_bulk_g2s(smem_d + (2 * s + 1) * TILE_DOUBLES, p.Bp + bj * BT * KLD, TILE_BYTES, &full_bar[s]);
    };
    if (tid == 0 && my_tiles > 0) issue(0);

    int prev_chunk = -1;          // staged mode: chunk of the tile this CTA finished last (-1: none to announce)
    uint32_t prev_total = 0;
    auto announce = [&]() {       // thread 0, after a CTA barrier that follows the tile's last store
        __threadfence();
        const uint32_t old = atomicAdd(p.prog_cnt + prev_chunk, 1u);
        if (old + 1u == prev_total) {
            __threadfence();
            asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p.prog_flag + prev_chunk), "r"(p.epoch) : "memory");
        }
    };
    for (int64_t tl = 0; tl < my_tiles; ++tl) {
        const int s = (int)(tl & 1);
        __syncthreads();   // every warp is done with stage s^1 (tile tl-1): refill it while this tile computes
        if (CYC == 2 && tid == 0 && prev_chunk >= 0) announce();
        if (tid == 0 && tl + 1 < my_tiles) issue(tl + 1);
        const int64_t t = (int64_t)blockIdx.x + tl * gridDim.x;
        int64_t bi, bj, gbi;
        bool extra = false;
        if (CYC) cyclic_tile_coords(p, t, bi, gbi, bj, extra);
        else { tile_coords(p, t, bi, bj); gbi = p.tile_r0 + bi; }
        const int64_t grow_tile = gbi * BT, gcol_tile = bj * BT;
        bool mirror = (p.uplo == GABO_SYMM_MIRROR) && (gcol_tile != grow_tile);
        if (CYC) {   // hybrid: recomputed upper tiles have no mirror image to write, and their lower twins do not send one
            const int64_t PI = gbi / CYC_PT, PJ = bj / CYC_PT;
            if (extra || cyclic_recompute_m<CYC>(PI, PJ, p.cyc_world, p.cyc_frac)) mirror = false;
        }
        const bool diag_tile = p.symmetric && (gcol_tile == grow_tile);
        // where the tile and its mirror image are stored: rows of K are local rows of this rank (CYC: local row tile bi),
        // the mirror image belongs to the rank that owns global row panel bj / CYC_PT
        const int64_t lrow_tile = CYC ? bi * BT : grow_tile - p.row0;
        double* kmir_base = p.K;
        int64_t mir_ld = ldk;            // mirror image: entry (col jj, row) at kmir_base + jj * mir_ld + global row
        if (CYC) {
            const int64_t pj = bj / CYC_PT;
            const int rj = (int)(pj % p.cyc_world);
            kmir_base = p.peer[rj] + ((pj / p.cyc_world) * CYC_PT + (bj % CYC_PT)) * BT * p.ldk;
            if (CYC == 2) {
                prev_chunk = -1;
                if (!extra) {
                    int col_off, width;
                    const int c = staged_chunk(p, bi, gbi, prev_total, col_off, width);
                    prev_chunk = c;
                    const StageEnt e = p.stage_tab[c * 8 + rj];
                    // bound for another rank: slots below the staged range are stored remotely from here (kmir_base above),
                    // the others go into this chunk's staging block for that rank
                    if (mirror && rj != p.cyc_rank && pj / p.cyc_world >= e.first_slot) {
                        mir_ld = width;
                        kmir_base = p.stage + e.off + ((pj / p.cyc_world - e.first_slot) * (CYC_PT * BT) + (bj % CYC_PT) * BT) * mir_ld
                                    + col_off - grow_tile;
                    }
                }
            }
        } else {
            kmir_base = p.K + gcol_tile * p.ldk;
        }
        mbar_wait(&full_bar[s], (uint32_t)((tl >> 1) & 1));

        const double* b_base = smem_d + (2 * s + 1) * TILE_DOUBLES + (wn * 32 + g) * KLD + q;
        const int64_t gcol0 = gcol_tile + wn * 32 + 2 * q;
        double nb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) nb[j] = 0.0;
        if (p.kind == GABO_SQDIST_GAUSSIAN) {
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) { nb[2 * ni] = p.normB[gcol0 + ni * 8]; nb[2 * ni + 1] = p.normB[gcol0 + ni * 8 + 1]; }
        }
#pragma unroll 1
        for (int mi = 0; mi < 4; ++mi) {
            const double* a_ptr = smem_d + 2 * s * TILE_DOUBLES + (wm * 32 + mi * 8 + g) * KLD + q;
            double c[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) c[j] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KLD / 4; ++ks) {
                const double a = a_ptr[ks * 4];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(c[2 * ni], c[2 * ni + 1], a, b_base[ni * 8 * KLD + ks * 4]);
            }
            const int64_t grow = grow_tile + wm * 32 + mi * 8 + g;
            double na = 0.0;
            if (p.kind == GABO_SQDIST_GAUSSIAN) na = p.normA[grow];
            double v[8], arg[8];
            if (kind == GABO_SQDIST_GAUSSIAN) {
#pragma unroll
                for (int j = 0; j < 8; ++j) arg[j] = nbeta * fmax(fma(-2.0, c[j], na + nb[j]), 0.0);
            } else {
                double dd[8];
                acos_clamped_fast_n<8>(c, dd);                                         // sphere_utils_tf.py:59-60
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    arg[j] = (kind == GABO_SPHERE_GAUSSIAN) ? (nbeta * dd[j]) * dd[j] : nbeta * dd[j];   // kernels_sphere_tf.py:83-86, 199-202
            }
            exp_neg_scaled_fast_n<8>(arg, v, exptab, variance, exp_slow_hi);                       // kernels_sphere_tf.py:88, 204
            if (diag_tile) {   // K(X,X) diagonal is exactly `variance` (matches Kdiag, kernels_sphere_tf.py:105)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (gcol0 + (j >> 1) * 8 + (j & 1) == grow) v[j] = variance;
            }
            if (grow < row1) {
                double* krow = Kbase + (lrow_tile + wm * 32 + mi * 8 + g) * ldk;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const int64_t gcol = gcol0 + ni * 8;
                    if (gcol + 1 < n2 && vec_ok) {
                        st_cs_v2(krow + gcol, v[2 * ni], v[2 * ni + 1]);
                    } else {
                        if (gcol < n2) st_cs(krow + gcol, v[2 * ni]);
                        if (gcol + 1 < n2) st_cs(krow + gcol + 1, v[2 * ni + 1]);
                    }
                    if (mirror) {   // K[col][row] = K[row][col]; 8 lanes of equal q write 64 contiguous bytes (own HBM or a peer's over NVLink)
                        double* km = kmir_base + (int64_t)(wn * 32 + ni * 8 + 2 * q) * (CYC == 2 ? mir_ld : ldk) + grow;
                        st_cs(km, v[2 * ni]);
                        st_cs(km + (CYC == 2 ? mir_ld : ldk), v[2 * ni + 1]);
                    }
                }
            }
        }
    }
    if (CYC == 2) {
        __syncthreads();
        if (tid == 0 && prev_chunk >= 0) announce();
    }
}

template <int KLD>
int launch_points_gram(const GramParams& p, cudaStream_t stream) {
    constexpr size_t smem = (size_t)NSTAGE * 2 * BT * KLD * sizeof(double);
    GABO_CUDA_TRY(cudaFuncSetAttribute(points_gram_kernel<KLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GABO_CUDA_TRY(cudaFuncSetAttribute(points_gram_kernel_1c<KLD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GABO_CUDA_TRY(cudaFuncSetAttribute(points_gram_kernel_1c<KLD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GABO_CUDA_TRY(cudaFuncSetAttribute(points_gram_kernel_1c<KLD, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = p.ntiles < (int64_t)sm_count() ? p.ntiles : (int64_t)sm_count();
    if (grid <= 0) return 0;
    if (p.cyc_world > 0 && p.stage != nullptr) points_gram_kernel_1c<KLD, 2><<<(unsigned)grid, NTHREADS, smem, stream>>>(p);
    else if (p.cyc_world > 0) points_gram_kernel_1c<KLD, 1><<<(unsigned)grid, NTHREADS, smem, stream>>>(p);
    else if (p.nchunks == 1) points_gram_kernel_1c<KLD, 0><<<(unsigned)grid, NTHREADS, smem, stream>>>(p);
    else points_gram_kernel<KLD><<<(unsigned)grid, NTHREADS, smem, stream>>>(p);
    GABO_LAUNCH_CHECK();
    return 0;
}

inline int pick_kld(int dim) { return dim <= 4 ? 4 : (dim <= 20 ? 20 : 52); }

template <typename Tin>
int pack_dispatch(int kld, const Tin* X, int64_t n, int64_t ldx, int dim, double* P, int64_t n_pad, int nchunks,
                  double* norms, cudaStream_t stream) {
    const int warps = 8;
    unsigned grid = (unsigned)ceil_div64(n_pad, warps);
    if (kld == 4) pack_points_kernel<4, Tin><<<grid, warps * 32, 0, stream>>>(X, n, ldx, dim, P, n_pad, nchunks, norms);
    else if (kld == 20) pack_points_kernel<20, Tin><<<grid, warps * 32, 0, stream>>>(X, n, ldx, dim, P, n_pad, nchunks, norms);
    else pack_points_kernel<52, Tin><<<grid, warps * 32, 0, stream>>>(X, n, ldx, dim, P, n_pad, nchunks, norms);
    GABO_LAUNCH_CHECK();
    return 0;
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct PackPlan {
    int kld, nchunks;
    int64_t n1_pad, n2_pad;
    size_t bytesA, bytesB, bytesNorm1, bytesNorm2;
};

inline PackPlan make_plan(int64_t n1, int64_t n2, int dim) {
    PackPlan pl;
    pl.kld = pick_kld(dim);
    pl.nchunks = (dim + pl.kld - 1) / pl.kld;
    if (pl.nchunks < 1) pl.nchunks = 1;
    pl.n1_pad = ceil_div64(n1, BT) * BT;
    pl.n2_pad = ceil_div64(n2, BT) * BT;
    pl.bytesA = align256((size_t)pl.nchunks * pl.n1_pad * pl.kld * sizeof(double));
    pl.bytesB = align256((size_t)pl.nchunks * pl.n2_pad * pl.kld * sizeof(double));
    pl.bytesNorm1 = align256((size_t)pl.n1_pad * sizeof(double));
    pl.bytesNorm2 = align256((size_t)pl.n2_pad * sizeof(double));
    return pl;
}

}  // namespace

template <typename Tin>
int points_gram_impl(int kind, const Tin* X, int64_t n1, int64_t ldx, const Tin* X2, int64_t n2, int64_t ldx2, int dim,
                     double beta, double variance, double* K, int64_t ldk, int64_t row0, int64_t row1, int uplo,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (kind < GABO_SPHERE_GAUSSIAN || kind > GABO_SQDIST_GAUSSIAN) return -1;
    if (n1 < 0) return -3;
    if (n1 == 0) return 0;
    if (X == nullptr) return -2;
    if (ldx < dim) return -4;
    const bool symmetric = (X2 == nullptr);
    if (symmetric) { n2 = n1; ldx2 = ldx; }
    if (n2 < 0) return -6;
    if (ldx2 < dim) return -7;
    if (dim < 1) return -8;
    if (K == nullptr) return -11;
    if (ldk < n2) return -12;
    if (row0 < 0 || row0 % BT != 0 || row0 > n1) return -13;
    if (row1 < row0 || row1 > n1) return -14;
    if (uplo < GABO_FULL || uplo > GABO_SYMM_MIRROR) return -15;
    if (uplo != GABO_FULL && !symmetric) return -15;
    if (uplo == GABO_SYMM_MIRROR && !(row0 == 0 && row1 == n1)) return -15;
    if (row1 == row0 || n2 == 0) return 0;

    PackPlan pl = make_plan(n1, n2, dim);
    size_t need = pl.bytesA + pl.bytesNorm1 + (symmetric ? 0 : pl.bytesB + pl.bytesNorm2);
    if (workspace == nullptr) return -16;
    if (workspace_bytes < need) return -17;
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    if ((reinterpret_cast<uintptr_t>(ws) & 15) != 0) return -16;
    double* Ap = reinterpret_cast<double*>(ws);
    double* nA = reinterpret_cast<double*>(ws + pl.bytesA);
    double* Bp = Ap;
    double* nB = nA;
    const bool want_norm = (kind == GABO_SQDIST_GAUSSIAN);
    int rc = pack_dispatch<Tin>(pl.kld, X, n1, ldx, dim, Ap, pl.n1_pad, pl.nchunks, want_norm ? nA : nullptr, stream);
    if (rc) return rc;
    if (!symmetric) {
        Bp = reinterpret_cast<double*>(ws + pl.bytesA + pl.bytesNorm1);
        nB = reinterpret_cast<double*>(ws + pl.bytesA + pl.bytesNorm1 + pl.bytesB);
        rc = pack_dispatch<Tin>(pl.kld, X2, n2, ldx2, dim, Bp, pl.n2_pad, pl.nchunks, want_norm ? nB : nullptr, stream);
        if (rc) return rc;
    }

    GramParams p;
    p.Ap = Ap; p.Bp = Bp; p.normA = want_norm ? nA : nullptr; p.normB = want_norm ? nB : nullptr;
    p.K = K; p.ldk = ldk; p.n1_pad = pl.n1_pad; p.n2_pad = pl.n2_pad;
    p.row0 = row0; p.row1 = row1; p.n2 = n2; p.nchunks = pl.nchunks; p.kind = kind; p.uplo = uplo;
    p.symmetric = symmetric ? 1 : 0;
    p.vec_ok = ((ldk % 2) == 0 && (reinterpret_cast<uintptr_t>(K) & 15) == 0) ? 1 : 0;
    p.beta = beta; p.variance = variance;
    p.cyc_world = 0; p.cyc_rank = 0; p.cyc_frac = 0; p.cyc_lower_tiles = 0; p.cyc_extra = nullptr;
    p.cyc_nslots = 0; p.cyc_split = 0; p.stage = nullptr; p.stage_tab = nullptr; p.prog_cnt = nullptr; p.prog_flag = nullptr; p.epoch = 0;
    for (int r = 0; r < 8; ++r) p.peer[r] = nullptr;
    p.tile_r0 = row0 / BT;
    p.tiles_m = ceil_div64(row1 - row0, BT);
    p.tiles_n = ceil_div64(n2, BT);
    if (uplo == GABO_FULL) p.ntiles = p.tiles_m * p.tiles_n;
    else p.ntiles = p.tiles_m * p.tile_r0 + p.tiles_m * (p.tiles_m + 1) / 2;
    if (pl.kld == 4) return launch_points_gram<4>(p, stream);
    if (pl.kld == 20) return launch_points_gram<20>(p, stream);
    return launch_points_gram<52>(p, stream);
}

}  // namespace gabo

namespace gabo {
size_t points_gram_f32_tc_workspace_bytes(int64_t n1, int64_t n2, int dim);   // points_gram_f32_tc.cu
bool points_gram_f32_tc_supported(int dim);
}

extern "C" size_t gabo_points_gram_workspace_bytes(int64_t n1, int64_t n2, int dim) {
    if (n1 < 0 || n2 < 0 || dim < 1) return 0;
    gabo::PackPlan pl = gabo::make_plan(n1, n2, dim);
    size_t need = pl.bytesA + pl.bytesNorm1 + pl.bytesB + pl.bytesNorm2;
    if (gabo::points_gram_f32_tc_supported(dim)) {        // the FP32 tensor-core path packs hi / lo TF32 planes
        const size_t tc = gabo::points_gram_f32_tc_workspace_bytes(n1, n2, dim);
        if (tc > need) need = tc;
    }
    return need;
}

extern "C" int gabo_points_gram_f64(int kind, const double* X, int64_t n1, int64_t ldx, const double* X2, int64_t n2,
                                    int64_t ldx2, int dim, double beta, double variance, double* K, int64_t ldk,
                                    int64_t row0, int64_t row1, int uplo, void* workspace, size_t workspace_bytes,
                                    gabo_stream_t stream) {
    return gabo::points_gram_impl<double>(kind, X, n1, ldx, X2, n2, ldx2, dim, beta, variance, K, ldk, row0, row1, uplo,
                                          workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

// Multi-GPU symmetric Gram, 1-D block-cyclic over 512-row panels, symmetry exploited ACROSS GPUs: this rank computes the
// lower tiles of the row panels it owns, stores them into its own buffer and stores each off-diagonal tile's mirror image
// straight into the HBM of the rank that owns that row panel (peer mappings over NVLink / NVSwitch; no collective).
// K_bases_host: `world` device pointers (host array): base of every rank's [local_rows, ldk] buffer as mapped in THIS process.
// After the call + a stream sync + a barrier across ranks, every rank holds the full dense rows of its panels.
extern "C" int gabo_points_gram_cyclic_f64(int kind, const double* X, int64_t n, int64_t ldx, int dim, double beta,
                                           double variance, double* const* K_bases_host, int64_t ldk, int world, int rank,
                                           void* workspace, size_t workspace_bytes, gabo_stream_t stream_) {
    return gabo_points_gram_cyclic_hybrid_f64(kind, X, n, ldx, dim, beta, variance, K_bases_host, ldk, world, rank, 0,
                                              workspace, workspace_bytes, stream_);
}

extern "C" size_t gabo_points_gram_cyclic_workspace_bytes(int64_t n, int dim) {
    if (n <= 0 || dim <= 0) return 0;
    const int64_t npanels = (n + 511) / 512;
    return gabo_points_gram_workspace_bytes(n, n, dim) + (size_t)(npanels * npanels + 64) * sizeof(int32_t)
           + (size_t)(npanels + 16) * 8 * sizeof(gabo::StageEnt) + 256;      // recompute list + staging table
}

// Same, with part of the upper triangle RECOMPUTED instead of mirrored: recompute_frac256 / 256 of the panel pairs owned
// by different ranks (chosen by a hash of the pair, cyclic_recompute) are produced by the owner of the ROW panel itself,
// the others arrive as mirror stores from the owner of the column panel.  Balances the FP64 pipe against NVLink.
extern "C" int gabo_points_gram_cyclic_hybrid_f64(int kind, const double* X, int64_t n, int64_t ldx, int dim, double beta,
                                                  double variance, double* const* K_bases_host, int64_t ldk, int world,
                                                  int rank, int recompute_frac256, void* workspace, size_t workspace_bytes,
                                                  gabo_stream_t stream_) {
    using namespace gabo;
    if (recompute_frac256 < 0 || recompute_frac256 > 256) return -12;
    if (kind < GABO_SPHERE_GAUSSIAN || kind > GABO_SQDIST_GAUSSIAN) return -1;
    if (n < 0) return -3;
    if (n == 0) return 0;
    if (X == nullptr) return -2;
    if (ldx < dim) return -4;
    if (dim < 1 || dim > 52) return -5;               // single-chunk kernel only
    if (K_bases_host == nullptr) return -8;
    if (ldk < n) return -9;
    if (world < 1 || world > 8) return -10;
    if (rank < 0 || rank >= world) return -11;
    if (n % (BT * CYC_PT) != 0) return -3;            // whole 512-row panels only
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PackPlan pl = make_plan(n, n, dim);
    const size_t pack_bytes = align256(pl.bytesA + pl.bytesNorm1);
    const size_t need = recompute_frac256 > 0 ? gabo_points_gram_cyclic_workspace_bytes(n, dim) : pl.bytesA + pl.bytesNorm1;
    if (workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return -13;
    if (workspace_bytes < need) return -14;
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    double* Ap = reinterpret_cast<double*>(ws);
    double* nA = reinterpret_cast<double*>(ws + pl.bytesA);
    const bool want_norm = (kind == GABO_SQDIST_GAUSSIAN);
    int rc = pack_dispatch<double>(pl.kld, X, n, ldx, dim, Ap, pl.n1_pad, pl.nchunks, want_norm ? nA : nullptr, stream);
    if (rc) return rc;
    const int64_t npanels = n / (BT * CYC_PT);
    const int64_t nslots = (npanels > rank) ? (npanels - 1 - rank) / world + 1 : 0;
    GramParams p;
    p.Ap = Ap; p.Bp = Ap; p.normA = want_norm ? nA : nullptr; p.normB = p.normA;
    p.K = K_bases_host[rank]; p.ldk = ldk; p.n1_pad = pl.n1_pad; p.n2_pad = pl.n2_pad;
    p.row0 = 0; p.row1 = n; p.n2 = n; p.nchunks = 1; p.kind = kind; p.uplo = GABO_SYMM_MIRROR; p.symmetric = 1;
    p.vec_ok = ((ldk % 2) == 0 && (reinterpret_cast<uintptr_t>(p.K) & 15) == 0) ? 1 : 0;
    p.beta = beta; p.variance = variance;
    p.tile_r0 = 0; p.tiles_m = nslots * CYC_PT; p.tiles_n = n / BT;
    p.cyc_world = world; p.cyc_rank = rank;
    for (int r = 0; r < 8; ++r) p.peer[r] = (r < world) ? K_bases_host[r] : nullptr;
    p.cyc_lower_tiles = 8 * (int64_t)world * nslots * (nslots - 1) + (16 * (int64_t)rank + 10) * nslots;
    p.cyc_frac = recompute_frac256;
    p.cyc_extra = nullptr;
    int64_t extra_pairs = 0;
    if (recompute_frac256 > 0) {
        // (local slot, column panel) of every upper block this rank recomputes; small (< npanels^2 / 2 entries), rebuilt and
        // uploaded per call (a pageable-source cudaMemcpyAsync has consumed the host buffer when it returns)
        std::vector<int32_t> list;
        for (int64_t sl = 0; sl < nslots; ++sl) {
            const int64_t PJ = sl * world + rank;
            for (int64_t PI = PJ + 1; PI < npanels; ++PI)
                if (cyclic_recompute(PI, PJ, world, recompute_frac256)) { list.push_back((int32_t)sl); list.push_back((int32_t)PI); }
        }
        extra_pairs = (int64_t)list.size() / 2;
        if (pack_bytes + list.size() * sizeof(int32_t) > workspace_bytes) return -14;
        int32_t* dlist = reinterpret_cast<int32_t*>(ws + pack_bytes);
        if (!list.empty())
            GABO_CUDA_TRY(cudaMemcpyAsync(dlist, list.data(), list.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
        p.cyc_extra = dlist;
    }
    p.ntiles = p.cyc_lower_tiles + 16 * extra_pairs;
    p.cyc_nslots = (int)nslots; p.cyc_split = 0; p.stage = nullptr; p.stage_tab = nullptr; p.prog_cnt = nullptr; p.prog_flag = nullptr;
    p.epoch = 0;
    if (p.K == nullptr) return -8;
    if (pl.kld == 4) return launch_points_gram<4>(p, stream);
    if (pl.kld == 20) return launch_points_gram<20>(p, stream);
    return launch_points_gram<52>(p, stream);
}

// ---- staged mirror: the cross-GPU half of the symmetric Gram moves by copy engine, not by remote stores -------------------
namespace gabo {
namespace {
struct StagePlan {
    int64_t nslots, whole, nchunks;
    std::vector<StageEnt> tab;      // [nchunks][8]
    std::vector<int64_t> chunk_col0; // first global column of the chunk's block
    std::vector<int> chunk_width;
    size_t stage_doubles;
};
StagePlan make_stage_plan(int64_t n, int world, int rank, int frac256, int direct256, int split) {
    StagePlan sp;
    const int64_t npanels = n / (BT * CYC_PT);
    sp.nslots = (npanels > rank) ? (npanels - 1 - rank) / world + 1 : 0;
    if (split > sp.nslots) split = (int)sp.nslots;
    sp.whole = sp.nslots - split;
    sp.nchunks = sp.whole + (int64_t)split * CYC_PT;
    sp.tab.assign((size_t)sp.nchunks * 8, StageEnt{0, 0, 0});
    sp.chunk_col0.resize((size_t)sp.nchunks);
    sp.chunk_width.resize((size_t)sp.nchunks);
    int64_t off = 0;
    for (int64_t c = 0; c < sp.nchunks; ++c) {
        const int64_t s = c < sp.whole ? c : sp.whole + (c - sp.whole) / CYC_PT;
        const int rr0 = c < sp.whole ? 0 : (int)((c - sp.whole) % CYC_PT);
        const int width = c < sp.whole ? CYC_PT * BT : BT;
        const int64_t PI = s * world + rank;
        sp.chunk_col0[(size_t)c] = PI * CYC_PT * BT + rr0 * BT;
        sp.chunk_width[(size_t)c] = width;
        for (int r = 0; r < world; ++r) {
            if (r == rank) continue;
            // slots [0, a) recomputed by rank r, [a, first) stored remotely by the kernel, [first, count) staged + copy engine
            const int64_t count = staged_slot_count(PI, r, world), a = staged_first_slot(PI, r, world, frac256);
            int64_t first = a + ((count * direct256 + 128) >> 8);
            if (first > count) first = count;
            StageEnt e;
            e.off = off; e.first_slot = (int32_t)first; e.nrows = (int32_t)((count - first) * CYC_PT * BT);
            sp.tab[(size_t)c * 8 + r] = e;
            off += (int64_t)e.nrows * width;
        }
    }
    sp.stage_doubles = (size_t)off;
    return sp;
}

// per-device copy streams (one per destination rank) and their join events: created on first use, reused afterwards
struct CopyCtx { int dev; cudaStream_t st[GABO_MAX_PEERS]; cudaEvent_t ev[GABO_MAX_PEERS]; cudaEvent_t start; };
std::mutex g_copy_mutex;
std::vector<CopyCtx> g_copy_ctx;
int copy_ctx_get(CopyCtx* out) {
    int dev = 0;
    GABO_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_copy_mutex);
    for (auto& c : g_copy_ctx)
        if (c.dev == dev) { *out = c; return 0; }
    CopyCtx c;
    c.dev = dev;
    for (int i = 0; i < GABO_MAX_PEERS; ++i) {
        GABO_CUDA_TRY(cudaStreamCreateWithFlags(&c.st[i], cudaStreamNonBlocking));
        GABO_CUDA_TRY(cudaEventCreateWithFlags(&c.ev[i], cudaEventDisableTiming));
    }
    GABO_CUDA_TRY(cudaEventCreateWithFlags(&c.start, cudaEventDisableTiming));
    g_copy_ctx.push_back(c);
    *out = c;
    return 0;
}
}  // namespace
}  // namespace gabo

extern "C" size_t gabo_points_gram_cyclic_stage_bytes(int64_t n, int world, int rank, int recompute_frac256, int direct_frac256,
                                                      int split) {
    if (n <= 0 || world < 1 || world > 8 || rank < 0 || rank >= world || n % 512 != 0 || split < 0) return 0;
    if (recompute_frac256 < 0 || recompute_frac256 > 256 || direct_frac256 < 0 || direct_frac256 > 256) return 0;
    return gabo::make_stage_plan(n, world, rank, recompute_frac256, direct_frac256, split).stage_doubles * sizeof(double) + 256;
}

extern "C" int gabo_points_gram_cyclic_staged_f64(int kind, const double* X, int64_t n, int64_t ldx, int dim, double beta,
                                                  double variance, double* const* K_bases_host, int64_t ldk, int world,
                                                  int rank, int recompute_frac256, int direct_frac256, int split, double* stage,
                                                  size_t stage_bytes,
                                                  uint32_t* progress, int64_t progress_words, uint32_t epoch, void* workspace,
                                                  size_t workspace_bytes, gabo_stream_t stream_) {
    using namespace gabo;
    if (recompute_frac256 < 0 || recompute_frac256 > 256) return -12;
    if (direct_frac256 < 0 || direct_frac256 > 256) return -13;
    if (kind < GABO_SPHERE_GAUSSIAN || kind > GABO_SQDIST_GAUSSIAN) return -1;
    if (n < 0) return -3;
    if (n == 0) return 0;
    if (X == nullptr) return -2;
    if (ldx < dim) return -4;
    if (dim < 1 || dim > 52) return -5;
    if (K_bases_host == nullptr) return -8;
    if (ldk < n) return -9;
    if (world < 1 || world > 8) return -10;
    if (rank < 0 || rank >= world) return -11;
    if (n % (BT * CYC_PT) != 0) return -3;
    if (split < 0) return -14;
    for (int r = 0; r < world; ++r)
        if (K_bases_host[r] == nullptr) return -8;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    StagePlan sp = make_stage_plan(n, world, rank, recompute_frac256, direct_frac256, split);
    if (sp.stage_doubles > 0 && (stage == nullptr || stage_bytes < sp.stage_doubles * sizeof(double))) return -15;
    if (progress == nullptr || progress_words < 2 * sp.nchunks) return -17;
    PackPlan pl = make_plan(n, n, dim);
    const size_t pack_bytes = align256(pl.bytesA + pl.bytesNorm1);
    if (workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return -19;
    if (workspace_bytes < gabo_points_gram_cyclic_workspace_bytes(n, dim)) return -20;
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    double* Ap = reinterpret_cast<double*>(ws);
    double* nA = reinterpret_cast<double*>(ws + pl.bytesA);
    const bool want_norm = (kind == GABO_SQDIST_GAUSSIAN);
    int rc = pack_dispatch<double>(pl.kld, X, n, ldx, dim, Ap, pl.n1_pad, pl.nchunks, want_norm ? nA : nullptr, stream);
    if (rc) return rc;
    const int64_t npanels = n / (BT * CYC_PT);
    const int64_t nslots = sp.nslots;
    GramParams p;
    p.Ap = Ap; p.Bp = Ap; p.normA = want_norm ? nA : nullptr; p.normB = p.normA;
    p.K = K_bases_host[rank]; p.ldk = ldk; p.n1_pad = pl.n1_pad; p.n2_pad = pl.n2_pad;
    p.row0 = 0; p.row1 = n; p.n2 = n; p.nchunks = 1; p.kind = kind; p.uplo = GABO_SYMM_MIRROR; p.symmetric = 1;
    p.vec_ok = ((ldk % 2) == 0 && (reinterpret_cast<uintptr_t>(p.K) & 15) == 0) ? 1 : 0;
    p.beta = beta; p.variance = variance;
    p.tile_r0 = 0; p.tiles_m = nslots * CYC_PT; p.tiles_n = n / BT;
    p.cyc_world = world; p.cyc_rank = rank;
    for (int r = 0; r < 8; ++r) p.peer[r] = (r < world) ? K_bases_host[r] : nullptr;
    p.cyc_lower_tiles = 8 * (int64_t)world * nslots * (nslots - 1) + (16 * (int64_t)rank + 10) * nslots;
    p.cyc_frac = recompute_frac256;
    // upper blocks this rank recomputes: for every later panel PI of another rank, my lowest slots (staged_recompute)
    std::vector<int32_t> list;
    for (int64_t sl = 0; sl < nslots; ++sl) {
        const int64_t PJ = sl * world + rank;
        for (int64_t PI = PJ + 1; PI < npanels; ++PI)
            if (staged_recompute(PI, PJ, world, recompute_frac256)) { list.push_back((int32_t)sl); list.push_back((int32_t)PI); }
    }
    const int64_t extra_pairs = (int64_t)list.size() / 2;
    int32_t* dlist = reinterpret_cast<int32_t*>(ws + pack_bytes);
    const size_t list_bytes = align256(list.size() * sizeof(int32_t) + 16);
    StageEnt* dtab = reinterpret_cast<StageEnt*>(ws + pack_bytes + list_bytes);
    if (pack_bytes + list_bytes + sp.tab.size() * sizeof(StageEnt) > workspace_bytes) return -20;
    if (!list.empty())
        GABO_CUDA_TRY(cudaMemcpyAsync(dlist, list.data(), list.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    if (!sp.tab.empty())
        GABO_CUDA_TRY(cudaMemcpyAsync(dtab, sp.tab.data(), sp.tab.size() * sizeof(StageEnt), cudaMemcpyHostToDevice, stream));
    p.cyc_extra = dlist;
    p.ntiles = p.cyc_lower_tiles + 16 * extra_pairs;
    p.cyc_nslots = (int)nslots; p.cyc_split = (int)(nslots - sp.whole);
    p.stage = stage != nullptr ? stage : reinterpret_cast<double*>(ws);   // non-null selects the staged kernel (world == 1: unused)
    p.stage_tab = dtab; p.prog_cnt = progress; p.prog_flag = progress + sp.nchunks; p.epoch = epoch;
    if (sp.nchunks > 0) GABO_CUDA_TRY(cudaMemsetAsync(progress, 0, (size_t)sp.nchunks * sizeof(uint32_t), stream));
    // The copy streams (a per-device pool, not ordered with the caller's stream) must not look at the flag words before
    // everything the caller enqueued so far has run -- in particular the zero-initialisation of a freshly allocated `progress`
    // (recycled memory may hold a stale flag >= epoch: the copy would fire before the kernel produced the chunk; seen once in
    // the emulated-ranks test after other tests had freed such buffers).  Recorded BEFORE the kernel launch: the overlap stays.
    CopyCtx cc;
    if (world > 1) {
        if ((rc = copy_ctx_get(&cc))) return rc;
        GABO_CUDA_TRY(cudaEventRecord(cc.start, stream));
    }
    if (pl.kld == 4) rc = launch_points_gram<4>(p, stream);
    else if (pl.kld == 20) rc = launch_points_gram<20>(p, stream);
    else rc = launch_points_gram<52>(p, stream);
    if (rc) return rc;
    if (world == 1) return 0;
    // copy streams: per destination rank, chunk by chunk as the kernel completes them
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        bool any = false;
        GABO_CUDA_TRY(cudaStreamWaitEvent(cc.st[r], cc.start, 0));
        for (int64_t c = 0; c < sp.nchunks; ++c) {
            const StageEnt& e = sp.tab[(size_t)c * 8 + r];
            if (e.nrows <= 0) continue;
            if ((rc = gabo_flag_wait_geq(p.prog_flag + c, epoch, cc.st[r]))) return rc;
            const int width = sp.chunk_width[(size_t)c];
            double* dst = K_bases_host[r] + (int64_t)e.first_slot * CYC_PT * BT * ldk + sp.chunk_col0[(size_t)c];
            GABO_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)ldk * sizeof(double), stage + e.off, (size_t)width * sizeof(double),
                                            (size_t)width * sizeof(double), (size_t)e.nrows, cudaMemcpyDefault, cc.st[r]));
            any = true;
        }
        if (any) {
            GABO_CUDA_TRY(cudaEventRecord(cc.ev[r], cc.st[r]));
            GABO_CUDA_TRY(cudaStreamWaitEvent(stream, cc.ev[r], 0));
        }
    }
    return 0;
}
