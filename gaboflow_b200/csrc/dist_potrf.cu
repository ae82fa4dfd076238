// Host-side driver of the distributed right-looking Cholesky (1-D block-cyclic row panels, peer-memory panel exchange, one
// panel of look-ahead) as ONE C call per rank.  All arithmetic and all transfers are the kernels / copy-engine operations of
// linalg.cu and gabo_api.cu; this file only sequences them on the caller's main stream and a high-priority side stream.
// The same sequence exists in Python (gaboflow_b200/dist.py: potrf_block_cyclic_p2p_, kept as the readable specification and
// for GABO_DIST_PYLOOP=1): ~25 ctypes / tensor operations per panel there cost 0.2-0.3 ms of host time per panel, which is
// what the device needs for a whole panel step at 8 GPUs, so the tail of the factorisation ran at the speed of the host.
#include <cstdint>
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
