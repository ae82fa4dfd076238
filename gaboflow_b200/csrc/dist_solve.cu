// Host-side driver of the distributed forward substitution (1-D block-cyclic row panels, one NVSwitch node), as ONE C call.
// The arithmetic is in the kernels of linalg.cu / trsv.cu / gabo_api.cu; this file only sequences them.  Round 2 first drove
// the same sequence from Python: ~8 ctypes calls + tensor slicing per panel cost ~100 us of host time per step, more than the
// ~30 us the step needs on the device, so the solve was host-bound (6.6 ms at N = 32768 on 2 GPUs).
#include <cstdint>
#include "gabo_common.cuh"

// L x = b for the block-cyclically stored factor (global panel p of nb rows on rank p % world, local slot p / world; A_loc
// [local_rows, >= n], B_loc [local_rows, nrhs] in place).  Step kp: the owner solves its block against the diagonal block
// (inverses of its 128-blocks in winv_host[kp]) and publishes it into slot kp of every peer's x area (xarea_host[r], nb * nrhs
// doubles per slot) together with flag value base + kp + 1 (flag_host[r]); everybody else waits for the flag on the stream.
// The owner of panel kp + 1 updates THOSE rows first, solves and publishes, and only then updates the rest of its rows.
extern "C" int gabo_solve_cyclic_p2p_f64(int64_t n, int nb, int world, int rank, const double* A_loc, int64_t lda, double* B_loc,
                                         int64_t ldb, int64_t nrhs, const double* const* winv_host, double* const* xarea_host,
                                         uint32_t* const* flag_host, uint32_t base, void* scratch, size_t scratch_bytes,
                                         gabo_stream_t stream) {
    if (n < 0) return -1;
    if (nb < 128 || nb % 128 != 0) return -2;
    if (world < 1 || world > GABO_MAX_PEERS) return -3;
    if (rank < 0 || rank >= world) return -4;
    if (nrhs < 1) return -9;
    if (n == 0) return 0;
    if (A_loc == nullptr) return -5;
    if (B_loc == nullptr) return -7;
    if (lda < n) return -6;
    if (ldb < nrhs) return -8;
    if (winv_host == nullptr) return -10;
    if (xarea_host == nullptr) return -11;
    if (flag_host == nullptr) return -12;
    if (scratch == nullptr || scratch_bytes < 256) return -14;
    const int64_t npan = (n + nb - 1) / nb;
    // local rows of the factored block: sum of the rows of my panels < npan
    int64_t rows_f = 0;
    for (int64_t p = rank; p < npan; p += world) rows_f += (n - p * nb < nb) ? (n - p * nb) : nb;
    auto slots_after = [&](int64_t kp) -> int64_t { return kp < rank ? 0 : (kp - rank) / world + 1; };
    double* dst[GABO_MAX_PEERS];
    uint32_t* flg[GABO_MAX_PEERS];
    int nother = 0;
    for (int r = 0; r < world; ++r)
        if (r != rank) { flg[nother] = flag_host[r]; ++nother; }
    const double* xloc = xarea_host[rank];
    int rc;

    auto solve_and_publish = [&](int64_t kp) -> int {
        const int64_t k0 = kp * nb, kb = (n - k0 < nb) ? (n - k0) : nb;
        const int64_t s = kp / world;
        double* bk = B_loc + s * nb * ldb;
        if (winv_host[kp] == nullptr) return -10;
        int r2 = gabo_trsm_winv_f64(0, kb, nrhs, A_loc + s * nb * lda + k0, lda, winv_host[kp], bk, ldb, scratch, scratch_bytes, stream);
        if (r2) return r2;
        if (kp + 1 < npan && nother > 0) {                                   // the last block has no consumer
            if (ldb != nrhs) return -8;                                       // published as one contiguous block
            int j = 0;
            for (int r = 0; r < world; ++r)
                if (r != rank) dst[j++] = xarea_host[r] + kp * (int64_t)nb * nrhs;
            r2 = gabo_push_signal_f64(bk, kb * nrhs, dst, flg, nother, base + (uint32_t)kp + 1u, stream);
        }
        return r2;
    };

    if (rank == 0 % world) { if ((rc = solve_and_publish(0))) return rc; }
    for (int64_t kp = 0; kp + 1 < npan; ++kp) {
        const int64_t k0 = kp * nb, kb = (n - k0 < nb) ? (n - k0) : nb;
        const double* xk;
        if (rank != (int)(kp % world)) {
            if ((rc = gabo_flag_wait_geq(flag_host[rank], base + (uint32_t)kp + 1u, stream))) return rc;
            xk = xloc + kp * (int64_t)nb * nrhs;
        } else {
            xk = B_loc + (kp / world) * nb * ldb;
        }
        const int64_t xld = nrhs;
        int64_t lr0 = slots_after(kp) * nb;
        if (rank == (int)((kp + 1) % world)) {                               // critical path first
            const int64_t kb1 = (n - (kp + 1) * nb < nb) ? (n - (kp + 1) * nb) : nb;
            if ((rc = gabo_solve_update_f64(kb1, kb, nrhs, A_loc + lr0 * lda + k0, lda, xk, (rank == (int)(kp % world)) ? ldb : xld,
                                            B_loc + lr0 * ldb, ldb, stream))) return rc;
            if ((rc = solve_and_publish(kp + 1))) return rc;
            lr0 += kb1;
        }
        if (rows_f - lr0 > 0) {
            if ((rc = gabo_solve_update_f64(rows_f - lr0, kb, nrhs, A_loc + lr0 * lda + k0, lda, xk, (rank == (int)(kp % world)) ? ldb : xld,
                                            B_loc + lr0 * ldb, ldb, stream))) return rc;
        }
    }
    return 0;
}
