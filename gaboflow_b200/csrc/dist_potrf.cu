// Host-side driver of the distributed right-looking Cholesky (1-D block-cyclic row panels, peer-memory panel exchange, one
// panel of look-ahead) as ONE C call per rank.  All arithmetic and all transfers are the kernels / copy-engine operations of
// linalg.cu and gabo_api.cu; this file only sequences them on the caller's main stream and a high-priority side stream.
// The same sequence exists in Python (gaboflow_b200/dist.py: potrf_block_cyclic_p2p_, kept as the readable specification and
// for GABO_DIST_PYLOOP=1): ~25 ctypes / tensor operations per panel there cost 0.2-0.3 ms of host time per panel, which is
// what the device needs for a whole panel step at 8 GPUs, so the tail of the factorisation ran at the speed of the host.
#include <cstdint>
#include <mutex>
#include <vector>
#include "gabo_common.cuh"

extern "C" int gabo_event_create(void** ev) {
    if (ev == nullptr) return -1;
    cudaEvent_t e;
    GABO_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    *ev = e;
    return 0;
}
extern "C" int gabo_event_destroy(void* ev) {
    if (ev != nullptr) GABO_CUDA_TRY(cudaEventDestroy(static_cast<cudaEvent_t>(ev)));
    return 0;
}

// In-place Cholesky of the leading n x n block of the symmetric matrix whose lower triangle is stored block-cyclically
// (global panel p of nb rows on rank p % world, local slot p / world; A_loc [local_rows, >= n]).
//   P_host  [world * 3]  block-column buffers: buffer b of rank r at P_host[3 r + b], each [n][nb] (own memory / CUDA IPC maps)
//   bc_host [world * 3]  diagonal-block buffers: factor [nb][nb] followed by the inverses of its 128-blocks [nb/128][128][128]
//   flag_host [world]    base of the 32-word flag area of every rank: [0] diagonal block arrived, [1 + r] rows of rank r arrived
//   base                 flag epoch: values base + kp + 1 are used, callers advance it by npan + 1 per factorisation
//   info_dev [npan]      per-panel LAPACK info of the diagonal blocks (device; only the owner's entries are written)
//   winv_store           NULL, or [local slots][nb/128][128][128]: the owner keeps the block inverses for the substitution
// Main stream: wait P(kp) flags -> U1(kp) -> U2(kp).  Side stream (released by U1): owner factors the diagonal block kp + 1, pushes
// it with copy engines, signals; every rank solves its rows of block column kp + 1 with the result stored into P[(kp+1) % 3] of
// EVERY rank from inside the solve kernel, then signals.
extern "C" int gabo_potrf_cyclic_p2p_f64(int64_t n, int nb, int world, int rank, double* A_loc, int64_t lda, double sigma2,
                                         double* const* P_host, double* const* bc_host, uint32_t* const* flag_host, uint32_t base,
                                         int32_t* info_dev, double* winv_store, void* ws_diag, size_t ws_diag_bytes, void* ws_trsm,
                                         size_t ws_trsm_bytes, gabo_stream_t main_, gabo_stream_t side_, void* ev_u1_, void* ev_side_) {
    if (n < 0) return -1;
    if (nb < 128 || nb > 512 || nb % 128 != 0) return -2;
    if (world < 1 || world > GABO_MAX_PEERS) return -3;
    if (rank < 0 || rank >= world) return -4;
    if (n == 0) return 0;
    if (A_loc == nullptr) return -5;
    if (lda < n) return -6;
    if (P_host == nullptr) return -8;
    if (bc_host == nullptr) return -9;
    if (flag_host == nullptr) return -10;
    if (info_dev == nullptr) return -12;
    if (ws_diag == nullptr) return -14;
    if (ws_trsm == nullptr) return -16;
    if (ev_u1_ == nullptr || ev_side_ == nullptr) return -20;
    cudaStream_t main_s = static_cast<cudaStream_t>(main_), side_s = static_cast<cudaStream_t>(side_);
    cudaEvent_t ev_u1 = static_cast<cudaEvent_t>(ev_u1_), ev_side = static_cast<cudaEvent_t>(ev_side_);
    const int64_t npan = (n + nb - 1) / nb;
    const int nwb = nb / 128;
    const size_t winv_elems = (size_t)nwb * 128 * 128;
    const size_t bc_bytes = ((size_t)nb * nb + winv_elems) * sizeof(double);
    int64_t rows_f = 0;
    for (int64_t p = rank; p < npan; p += world) rows_f += (n - p * nb < nb) ? (n - p * nb) : nb;
    auto slots_after = [&](int64_t kp) -> int64_t { return kp < rank ? 0 : (kp - rank) / world + 1; };
    uint32_t* f_bc[GABO_MAX_PEERS];      // flag [0] of the other ranks
    uint32_t* f_rows[GABO_MAX_PEERS];    // flag [1 + rank] of the other ranks
    int nother = 0;
    for (int r = 0; r < world; ++r)
        if (r != rank) { f_bc[nother] = flag_host[r]; f_rows[nother] = flag_host[r] + 1 + rank; ++nother; }
    GABO_CUDA_TRY(cudaMemsetAsync(info_dev, 0, sizeof(int32_t) * npan, main_s));
    int rc;

    auto diag_and_push = [&](int64_t kp, cudaStream_t s) -> int {
        const int b = (int)(kp % 3);
        const int64_t k0 = kp * nb, kb = (n - k0 < nb) ? (n - k0) : nb;
        double* Lkk = bc_host[3 * rank + b];
        double* Wv = Lkk + (size_t)nb * nb;
        const int64_t sl = kp / world;
        double* blk = A_loc + sl * nb * lda + k0;
        int r2 = gabo_potrf_diag_f64(kb, blk, lda, sigma2, info_dev + kp, Wv, ws_diag, ws_diag_bytes, s);
        if (r2) return r2;
        GABO_CUDA_TRY(cudaMemcpy2DAsync(Lkk, (size_t)nb * sizeof(double), blk, (size_t)lda * sizeof(double), (size_t)kb * sizeof(double),
                                        (size_t)kb, cudaMemcpyDeviceToDevice, s));
        if (winv_store != nullptr)
            GABO_CUDA_TRY(cudaMemcpyAsync(winv_store + sl * winv_elems, Wv, winv_elems * sizeof(double), cudaMemcpyDeviceToDevice, s));
        if (k0 + kb >= n) return 0;                    // last panel: nobody solves against it
        for (int r = 0; r < world; ++r)
            if (r != rank) GABO_CUDA_TRY(cudaMemcpyAsync(bc_host[3 * r + b], Lkk, bc_bytes, cudaMemcpyDefault, s));
        return gabo_flag_signal(f_bc, nother, base + (uint32_t)kp + 1u, s);
    };
    auto solve_and_scatter = [&](int64_t kp, cudaStream_t s) -> int {
        const int b = (int)(kp % 3);
        const int64_t k0 = kp * nb, kb = (n - k0 < nb) ? (n - k0) : nb, k1 = k0 + kb;
        if (k1 >= n) return 0;
        int r2;
        if (rank != (int)(kp % world)) {
            if ((r2 = gabo_flag_wait_geq(flag_host[rank], base + (uint32_t)kp + 1u, s))) return r2;
        }
        const double* Lkk = bc_host[3 * rank + b];
        const double* Wv = Lkk + (size_t)nb * nb;
        const int64_t s0 = slots_after(kp), lr0 = s0 * nb, m = rows_f - lr0;
        if (m > 0) {
            double* dst[GABO_MAX_PEERS];
            for (int r = 0; r < world; ++r) dst[r] = P_host[3 * r + b];
            if ((r2 = gabo_trsm_right_scatter_f64(m, kb, Lkk, nb, Wv, A_loc + lr0 * lda + k0, lda, dst, world, nb,
                                                  (s0 * world + rank) * nb - k1, nb, world, ws_trsm, ws_trsm_bytes, s))) return r2;
        }
        return gabo_flag_signal(f_rows, nother, base + (uint32_t)kp + 1u, s);
    };

    if (rank == 0) { if ((rc = diag_and_push(0, main_s))) return rc; }
    if ((rc = solve_and_scatter(0, main_s))) return rc;
    for (int64_t kp = 0; kp + 1 < npan; ++kp) {
        const int64_t k0 = kp * nb, k1 = k0 + nb, k2 = (k1 + nb < n) ? (k1 + nb) : n;
        const int64_t lr0 = slots_after(kp) * nb, m = rows_f - lr0;
        for (int r = 0; r < world; ++r)
            if (r != rank) { if ((rc = gabo_flag_wait_geq(flag_host[rank] + 1 + r, base + (uint32_t)kp + 1u, main_s))) return rc; }
        if (kp > 0) GABO_CUDA_TRY(cudaStreamWaitEvent(main_s, ev_side, 0));
        const double* P = P_host[3 * rank + (int)(kp % 3)];
        const int64_t rb = (lr0 / nb) * world * nb;
        if (m > 0) {   // U1: the next block column only
            if ((rc = gabo_syrk_cyclic_f64(m, k2 - k1, nb, A_loc + lr0 * lda + k0, lda, P, nb, A_loc + lr0 * lda + k1, lda, rb, k1, nb,
                                           world, rank, main_s))) return rc;
        }
        GABO_CUDA_TRY(cudaEventRecord(ev_u1, main_s));
        GABO_CUDA_TRY(cudaStreamWaitEvent(side_s, ev_u1, 0));
        if (rank == (int)((kp + 1) % world)) { if ((rc = diag_and_push(kp + 1, side_s))) return rc; }
        if ((rc = solve_and_scatter(kp + 1, side_s))) return rc;
        GABO_CUDA_TRY(cudaEventRecord(ev_side, side_s));
        if (m > 0 && k2 < n) {   // U2: the rest of the trailing matrix, under which the side stream works
            if ((rc = gabo_syrk_cyclic_f64(m, n - k2, nb, A_loc + lr0 * lda + k0, lda, P + (k2 - k1) * (int64_t)nb, nb,
                                           A_loc + lr0 * lda + k2, lda, rb, k2, nb, world, rank, main_s))) return rc;
        }
    }
    GABO_CUDA_TRY(cudaEventRecord(ev_side, side_s));
    GABO_CUDA_TRY(cudaStreamWaitEvent(main_s, ev_side, 0));
    return 0;
}

// ---- critical-chain schedule ---------------------------------------------------------------------------------------------
// Same arithmetic (per entry and in order) as gabo_potrf_cyclic_p2p_f64, different sequencing.  The true serial dependency of a
// right-looking Cholesky over panels is
//     D(k) factor diagonal block k  ->  push to the owner of panel k+1  ->  Sa(k): that owner solves ITS OWN 512 rows of block
//     column k  ->  U1a(k): updates its diagonal block  ->  D(k+1),
// everything else -- the other ranks' row solves, the rest of U1, all of U2 -- only has to keep up.  The look-ahead schedule
// above put the whole row solve, a wait for EVERY rank's rows and the whole of U1 on that chain (~0.57 ms per panel at 8 GPUs,
// more than U2 takes for the last half of the panels).  Here:
//   H  (high priority)  owner of k+1:  wait bc(k) -> Sa(k) -> flag rows_a(k) -> [U2(k-1) done] U1a(k) -> D(k+1) -> push (the owner
//                       of k+2 first, with its own flag) -> flag bc(k+1)
//   S2 (high priority)  every rank:    wait bc(k) -> Sb(k): the rest of its rows -> flag rows(k)
//   M  (caller)         every rank:    [Sb(k), rows_a(k)] U1b(k)  ->  [rows(k) of every rank] U2(k)
// U1 needs only the 512 x 512 block Sa(k) produced (flag rows_a), not everybody's rows; only U2 waits for all rows.
// Ranks are no longer throttled to within two panels of each other (the owner chain can run up to `world` panels ahead of a
// rank that owns none of them), so the block-column and diagonal-block buffers rotate over nbuf >= world + 1 instead of 3.
//   flag words per rank: [0] bc, [1 + r] rows of rank r, [9] rows_a.
namespace {
struct ChainEvents { int dev; cudaEvent_t sa, sb, u1b, u2, d; cudaStream_t s2; };
static std::mutex g_chain_mutex;
static std::vector<ChainEvents> g_chain;
static int chain_events_get(ChainEvents* out) {
    int dev = 0;
    GABO_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_chain_mutex);
    for (auto& c : g_chain)
        if (c.dev == dev) { *out = c; return 0; }
    ChainEvents c;
    c.dev = dev;
    GABO_CUDA_TRY(cudaEventCreateWithFlags(&c.sa, cudaEventDisableTiming));
    GABO_CUDA_TRY(cudaEventCreateWithFlags(&c.sb, cudaEventDisableTiming));
    GABO_CUDA_TRY(cudaEventCreateWithFlags(&c.u1b, cudaEventDisableTiming));
    GABO_CUDA_TRY(cudaEventCreateWithFlags(&c.u2, cudaEventDisableTiming));
    GABO_CUDA_TRY(cudaEventCreateWithFlags(&c.d, cudaEventDisableTiming));
    int least = 0, greatest = 0;
    GABO_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    GABO_CUDA_TRY(cudaStreamCreateWithPriority(&c.s2, cudaStreamNonBlocking, greatest));
    g_chain.push_back(c);
    *out = c;
    return 0;
}
}  // namespace

extern "C" int gabo_potrf_cyclic_chain_f64(int64_t n, int nb, int world, int rank, double* A_loc, int64_t lda, double sigma2,
                                           int nbuf, double* const* P_host, double* const* bc_host, uint32_t* const* flag_host,
                                           uint32_t base, int32_t* info_dev, double* winv_store, void* ws_diag, size_t ws_diag_bytes,
                                           void* ws_trsm, size_t ws_trsm_bytes, gabo_stream_t main_, gabo_stream_t side_) {
    if (n < 0) return -1;
    if (nb < 128 || nb > 512 || nb % 128 != 0) return -2;
    if (world < 1 || world > GABO_MAX_PEERS) return -3;
    if (rank < 0 || rank >= world) return -4;
    if (n == 0) return 0;
    if (A_loc == nullptr) return -5;
    if (lda < n) return -6;
    if (nbuf < 3 || (world > 1 && nbuf < world + 1)) return -8;
    if (P_host == nullptr) return -9;
    if (bc_host == nullptr) return -10;
    if (flag_host == nullptr) return -11;
    if (info_dev == nullptr) return -13;
    if (ws_diag == nullptr) return -15;
    if (ws_diag_bytes < gabo_potrf_workspace_bytes(nb) || ws_diag_bytes < gabo_trsm_right_workspace_bytes(nb, nb)) return -16;
    if (ws_trsm == nullptr) return -17;
    cudaStream_t M = static_cast<cudaStream_t>(main_), H = static_cast<cudaStream_t>(side_);
    ChainEvents ev;
    int rc;
    if ((rc = chain_events_get(&ev))) return rc;
    cudaStream_t S2 = ev.s2;
    const int64_t npan = (n + nb - 1) / nb;
    const int nwb = nb / 128;
    const size_t winv_elems = (size_t)nwb * 128 * 128;
    const size_t bc_bytes = ((size_t)nb * nb + winv_elems) * sizeof(double);
    int64_t rows_f = 0;
    for (int64_t p = rank; p < npan; p += world) rows_f += (n - p * nb < nb) ? (n - p * nb) : nb;
    auto slots_after = [&](int64_t kp) -> int64_t { return kp < rank ? 0 : (kp - rank) / world + 1; };
    uint32_t* f_rows[GABO_MAX_PEERS];    // flag [1 + rank] of the other ranks
    uint32_t* f_rowsa[GABO_MAX_PEERS];   // flag [9] of the other ranks
    int nother = 0;
    for (int r = 0; r < world; ++r)
        if (r != rank) { f_rows[nother] = flag_host[r] + 1 + rank; f_rowsa[nother] = flag_host[r] + 9; ++nother; }
    GABO_CUDA_TRY(cudaMemsetAsync(info_dev, 0, sizeof(int32_t) * npan, M));
    // H and S2 start after whatever the caller queued on its stream (the Gram that produced A_loc)
    GABO_CUDA_TRY(cudaEventRecord(ev.u2, M));
    GABO_CUDA_TRY(cudaStreamWaitEvent(H, ev.u2, 0));
    GABO_CUDA_TRY(cudaStreamWaitEvent(S2, ev.u2, 0));

    auto diag_and_push = [&](int64_t kp) -> int {        // on H
        const int b = (int)(kp % nbuf);
        const int64_t k0 = kp * nb, kb = (n - k0 < nb) ? (n - k0) : nb;
        double* Lkk = bc_host[nbuf * rank + b];
        double* Wv = Lkk + (size_t)nb * nb;
        const int64_t sl = kp / world;
        double* blk = A_loc + sl * nb * lda + k0;
        int r2 = gabo_potrf_diag_f64(kb, blk, lda, sigma2, info_dev + kp, Wv, ws_diag, ws_diag_bytes, H);
        if (r2) return r2;
        GABO_CUDA_TRY(cudaMemcpy2DAsync(Lkk, (size_t)nb * sizeof(double), blk, (size_t)lda * sizeof(double), (size_t)kb * sizeof(double),
                                        (size_t)kb, cudaMemcpyDeviceToDevice, H));
        if (k0 + kb < n && nother > 0) {
            // the owner of the next panel continues the chain: it gets the block and its flag first
            const int nxt = (int)((kp + 1) % world);
            if (nxt != rank) {
                GABO_CUDA_TRY(cudaMemcpyAsync(bc_host[nbuf * nxt + b], Lkk, bc_bytes, cudaMemcpyDefault, H));
                uint32_t* f1[1] = {flag_host[nxt]};
                if ((r2 = gabo_flag_signal(f1, 1, base + (uint32_t)kp + 1u, H))) return r2;
            }
            uint32_t* frest[GABO_MAX_PEERS];
            int nrest = 0;
            for (int r = 0; r < world; ++r)
                if (r != rank && r != nxt) {
                    GABO_CUDA_TRY(cudaMemcpyAsync(bc_host[nbuf * r + b], Lkk, bc_bytes, cudaMemcpyDefault, H));
                    frest[nrest++] = flag_host[r];
                }
            if (nrest > 0 && (r2 = gabo_flag_signal(frest, nrest, base + (uint32_t)kp + 1u, H))) return r2;
        }
        if (winv_store != nullptr)
            GABO_CUDA_TRY(cudaMemcpyAsync(winv_store + sl * winv_elems, Wv, winv_elems * sizeof(double), cudaMemcpyDeviceToDevice, H));
        GABO_CUDA_TRY(cudaEventRecord(ev.d, H));
        return 0;
    };
    // rows [lr, lr + m) of block column kp (local slot s0 onwards) against the diagonal block, result into every rank's P buffer
    auto solve_rows = [&](int64_t kp, int64_t lr, int64_t m, int64_t s0, void* ws, size_t ws_bytes, cudaStream_t s) -> int {
        if (m <= 0) return 0;
        const int b = (int)(kp % nbuf);
        const int64_t k0 = kp * nb, k1 = k0 + nb;
        const double* Lkk = bc_host[nbuf * rank + b];
        const double* Wv = Lkk + (size_t)nb * nb;
        double* dst[GABO_MAX_PEERS];
        for (int r = 0; r < world; ++r) dst[r] = P_host[nbuf * r + b];
        return gabo_trsm_right_scatter_f64(m, nb, Lkk, nb, Wv, A_loc + lr * lda + k0, lda, dst, world, nb,
                                           (s0 * world + rank) * nb - k1, nb, world, ws, ws_bytes, s);
    };

    if (rank == 0) { if ((rc = diag_and_push(0))) return rc; }
    for (int64_t kp = 0; kp + 1 < npan; ++kp) {
        const int64_t k0 = kp * nb, k1 = k0 + nb, k2 = (k1 + nb < n) ? (k1 + nb) : n;
        const int owner = (int)(kp % world), nxt = (int)((kp + 1) % world);
        const bool is_next = (rank == nxt);
        const int64_t s0 = slots_after(kp), lr0 = s0 * nb, m = rows_f - lr0;
        const int64_t ma = is_next ? (k2 - k1) : 0;       // my rows of panel kp+1: the diagonal block of the next step
        const int64_t mb = m - ma;
        const uint32_t fv = base + (uint32_t)kp + 1u;
        // ---- row solves of block column kp
        if (is_next) {
            if (rank != owner) { if ((rc = gabo_flag_wait_geq(flag_host[rank], fv, H))) return rc; }
            if (kp > 0) GABO_CUDA_TRY(cudaStreamWaitEvent(H, ev.u1b, 0));
            if ((rc = solve_rows(kp, lr0, ma, s0, ws_diag, ws_diag_bytes, H))) return rc;
            if (nother > 0 && (rc = gabo_flag_signal(f_rowsa, nother, fv, H))) return rc;
            GABO_CUDA_TRY(cudaEventRecord(ev.sa, H));
        }
        if (rank != owner) { if ((rc = gabo_flag_wait_geq(flag_host[rank], fv, S2))) return rc; }
        else GABO_CUDA_TRY(cudaStreamWaitEvent(S2, ev.d, 0));
        if (kp > 0) GABO_CUDA_TRY(cudaStreamWaitEvent(S2, ev.u1b, 0));
        if ((rc = solve_rows(kp, lr0 + ma, mb, s0 + (ma > 0 ? 1 : 0), ws_trsm, ws_trsm_bytes, S2))) return rc;
        if (is_next) GABO_CUDA_TRY(cudaStreamWaitEvent(S2, ev.sa, 0));
        if (nother > 0 && (rc = gabo_flag_signal(f_rows, nother, fv, S2))) return rc;
        GABO_CUDA_TRY(cudaEventRecord(ev.sb, S2));
        // ---- the chain: diagonal block of panel kp+1
        const double* P = P_host[nbuf * rank + (int)(kp % nbuf)];
        if (is_next) {
            if (kp > 0) GABO_CUDA_TRY(cudaStreamWaitEvent(H, ev.u2, 0));
            if ((rc = gabo_syrk_cyclic_f64(ma, k2 - k1, nb, A_loc + lr0 * lda + k0, lda, P, nb, A_loc + lr0 * lda + k1, lda,
                                           s0 * world * nb, k1, nb, world, rank, H))) return rc;
            if ((rc = diag_and_push(kp + 1))) return rc;
        }
        // ---- U1b, U2
        GABO_CUDA_TRY(cudaStreamWaitEvent(M, ev.sb, 0));
        if (is_next) GABO_CUDA_TRY(cudaStreamWaitEvent(M, ev.sa, 0));
        else if (world > 1) { if ((rc = gabo_flag_wait_geq(flag_host[rank] + 9, fv, M))) return rc; }
        const int64_t lrb = lr0 + ma, rbb = (lrb / nb) * world * nb;
        if (mb > 0) {
            if ((rc = gabo_syrk_cyclic_f64(mb, k2 - k1, nb, A_loc + lrb * lda + k0, lda, P, nb, A_loc + lrb * lda + k1, lda, rbb, k1,
                                           nb, world, rank, M))) return rc;
        }
        GABO_CUDA_TRY(cudaEventRecord(ev.u1b, M));
        for (int r = 0; r < world; ++r)
            if (r != rank) { if ((rc = gabo_flag_wait_geq(flag_host[rank] + 1 + r, fv, M))) return rc; }
        if (mb > 0 && k2 < n) {
            if ((rc = gabo_syrk_cyclic_f64(mb, n - k2, nb, A_loc + lrb * lda + k0, lda, P + (k2 - k1) * (int64_t)nb, nb,
                                           A_loc + lrb * lda + k2, lda, rbb, k2, nb, world, rank, M))) return rc;
        }
        GABO_CUDA_TRY(cudaEventRecord(ev.u2, M));
    }
    GABO_CUDA_TRY(cudaEventRecord(ev.d, H));
    GABO_CUDA_TRY(cudaStreamWaitEvent(M, ev.d, 0));
    GABO_CUDA_TRY(cudaEventRecord(ev.sb, S2));
    GABO_CUDA_TRY(cudaStreamWaitEvent(M, ev.sb, 0));
    return 0;
}
