"""Design tool: NumPy emulation of the FP32 dqds start + FP64 Newton polish used by spd_gram_kernel (csrc/spd_small.cuh).

Measures, on the C5 workload's pair distribution (and harder ones), how often the FP32 starts fail the FP64 validation and
how many dqds sweeps a 32-lane warp pays.  Not part of the product or of the tests.
"""
from __future__ import annotations

import sys

import numpy as np

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from oracle import gabo_oracle as orc  # noqa: E402

f32 = np.float32


def householder_tridiag(M):
    """Same elimination order as householder_tridiag<D>: row k, columns k+2.. zeroed.  M: [P,D,D] -> d [P,D], e [P,D-1]."""
    A = M.copy()
    P, D, _ = A.shape
    d = np.zeros((P, D))
    e = np.zeros((P, D - 1))
    for k in range(D - 2):
        x = A[:, k, k + 1:].copy()
        sigma = np.sum(x * x, axis=1)
        nrm = np.sqrt(sigma)
        alpha = -np.copysign(nrm, x[:, 0])
        v = x.copy()
        v[:, 0] -= alpha
        beta = 1.0 / (sigma - alpha * x[:, 0])
        sub = A[:, k + 1:, k + 1:]
        pv = beta[:, None] * np.einsum('pij,pj->pi', sub, v)
        kk = 0.5 * beta * np.sum(pv * v, axis=1)
        w = pv - kk[:, None] * v
        sub -= v[:, :, None] * w[:, None, :] + w[:, :, None] * v[:, None, :]
        d[:, k] = A[:, k, k]
        e[:, k] = alpha
    d[:, D - 2] = A[:, D - 2, D - 2]
    e[:, D - 2] = A[:, D - 2, D - 1]
    d[:, D - 1] = A[:, D - 1, D - 1]
    return d, e


def make_pairs(n_pts=600, d=6, seed=20242, wmin=0.5, wmax=2.0, npairs=200000):
    Xm = orc.synth_spd_mandel(n_pts, d, seed=seed, wmin=wmin, wmax=wmax)
    S = orc.mandel_to_sym(Xm)
    L = np.linalg.cholesky(S)
    W = np.linalg.inv(L)
    rng = np.random.default_rng(1)
    i = rng.integers(0, n_pts, npairs)
    j = rng.integers(0, n_pts, npairs)
    keep = i != j
    i, j = i[keep], j[keep]
    M = np.einsum('pij,pjk,plk->pil', W[i], S[j], W[i])
    M = 0.5 * (M + np.swapaxes(M, 1, 2))
    return M


def dqds_start_f32(ds, e2s, maxit=8, tol_shift=20, strategy='gap', stats=None):
    """FP32 dqds on the SPD tridiagonal (ds diag, e2s squared off-diagonals), trace 1.  Returns start values [P,D] and the
    per-lane sweep counts per block size."""
    P, D = ds.shape
    q = np.zeros((P, D), dtype=f32)
    e = np.zeros((P, D), dtype=f32)   # e[k] couples k,k+1; e[D-1] dummy
    dsf = ds.astype(f32)
    e2f = e2s.astype(f32)
    with np.errstate(all='ignore'):
        q[:, 0] = dsf[:, 0]
        for k in range(D - 1):
            e[:, k] = e2f[:, k] / q[:, k]
            q[:, k + 1] = dsf[:, k + 1] - e[:, k]
        sigma = np.zeros(P, dtype=f32)
        ev = np.zeros((P, D), dtype=f32)
        sweeps = np.zeros((P, D), dtype=np.int32)
        fails = np.zeros(P, dtype=np.int32)
        tol = f32(2.0 ** -tol_shift)
        for m in range(D - 1, 0, -1):
            for it in range(maxit):
                active = ~(e[:, m - 1] <= tol * (sigma + q[:, m]))   # NaN stays active
                if not active.any():
                    break
                # ---- shift ----
                qm, qm1, em1 = q[:, m], q[:, m - 1], e[:, m - 1]
                if strategy == 'gap':
                    em2 = e[:, m - 2] if m >= 2 else np.zeros(P, dtype=f32)
                    a2 = qm1 + em1
                    gap1 = a2 - qm - f32(1.0) * em2     # crude distance of the rest of the spectrum from q_m
                    b1sq = qm * em1
                    s_asym = qm - b1sq / gap1
                    good = (gap1 > 0) & (b1sq < f32(0.25) * gap1 * gap1)
                    qmin = q[:, :m + 1].min(axis=1)
                    s = np.where(good, np.maximum(s_asym, f32(0.5) * qmin), f32(0.25) * qmin)
                    s = np.where(good & (qmin < qm), f32(0.25) * qmin, s)
                elif strategy == 'newton':
                    g = np.ones(P, dtype=f32)
                    acc = np.zeros(P, dtype=f32)
                    for k in range(m + 1):
                        rq = f32(1.0) / q[:, k]
                        acc = acc + g * rq
                        if k < m:
                            g = f32(1.0) + e[:, k] * rq * g
                    s = f32(1.0) / acc
                else:
                    raise ValueError(strategy)
                s = np.where(active, s, f32(0)).astype(f32)
                # ---- sweep, retry with s/4 then 0 on failure ----
                qn = q.copy(); en = e.copy()
                todo = active.copy()
                for attempt in range(3):
                    dcur = q[:, 0] - s
                    bad = np.zeros(P, dtype=bool)
                    qt = q.copy(); et = e.copy()
                    for k in range(m):
                        qt[:, k] = dcur + e[:, k]
                        bad |= ~(qt[:, k] > 0)
                        t = q[:, k + 1] / qt[:, k]
                        et[:, k] = e[:, k] * t
                        dcur = dcur * t - s
                    qt[:, m] = dcur
                    late_ok = dcur > -tol * (sigma + s)   # tiny negative last pivot is convergence noise
                    bad |= ~(late_ok)
                    okl = todo & ~bad
                    qn[okl] = qt[okl]; en[okl] = et[okl]
                    sigma = np.where(okl, sigma + s, sigma).astype(f32)
                    sweeps[todo, m] += 1
                    todo = todo & bad
                    if not todo.any():
                        break
                    fails[todo] += 1
                    s = np.where(todo, s * f32(0.25) if attempt == 0 else f32(0), s).astype(f32)
                q, e = qn, en
            ev[:, m] = q[:, m] + sigma
        ev[:, 0] = q[:, 0] + sigma
    if stats is not None:
        stats['fails'] = fails
    return ev, sweeps


def newton_polish(ds, e2s, x0):
    """Two FP64 steps per eigenvalue (full Newton, then a step reusing the derivative), as tridiag_eigvals_mixed."""
    P, D = ds.shape
    x = x0.astype(np.float64).copy()
    ok = np.ones(P, dtype=bool)
    d1 = None
    for it in range(2):
        p2 = np.ones((P, D)); p1 = ds[:, :1] - x
        q2 = np.zeros((P, D)); q1 = -np.ones((P, D))
        for k in range(1, D):
            t = ds[:, k:k + 1] - x
            pn = t * p1 - e2s[:, k - 1:k] * p2
            qn = t * q1 - e2s[:, k - 1:k] * q2 - p1
            p2, q2, p1, q1 = p1, q1, pn, qn
        if it == 0:
            qsave = q1
            dlt = p1 / q1
            d1 = np.abs(dlt)
        else:
            dlt = p1 / qsave
            ok &= np.all(np.abs(dlt) <= np.maximum(1e-3 * d1, 1e-15), axis=1)
        x = x - dlt
    ok &= np.abs(x.sum(axis=1) - 1.0) <= 1e-13
    return x, ok


def run(M, label, **kw):
    d, e = householder_tridiag(M)
    tr = d.sum(axis=1)
    ds = d / tr[:, None]
    es = e / tr[:, None]
    e2s = es * es
    st = {}
    ev0, sweeps = dqds_start_f32(ds, e2s, stats=st, **kw)
    x, ok = newton_polish(ds, e2s, ev0)
    lam = np.sort(x * tr[:, None], axis=1)
    ref = np.linalg.eigvalsh(M)
    relerr = np.max(np.abs(lam - ref) / ref, axis=1)
    wide = ref[:, 0] <= 0.02 * ref[:, -1]
    P = len(M)
    nw = P // 32
    per_lane = sweeps.sum(axis=1)
    # steps a warp pays: per block size m, max over lanes of sweeps, times m steps
    sw = sweeps[:nw * 32].reshape(nw, 32, -1).max(axis=1)
    D = M.shape[1]
    steps_warp = (sw * np.arange(D)[None, :]).sum(axis=1)
    steps_lane = (sweeps * np.arange(D)[None, :]).sum(axis=1)
    print(f'{label}: pairs {P}  invalid {np.mean(~ok):.2e}  (wide {np.mean(wide):.2e})  maxrel(valid) '
          f'{relerr[ok].max():.2e}  sweeps/lane {per_lane.mean():.1f}  steps/lane {steps_lane.mean():.1f}  '
          f'steps/warp {steps_warp.mean():.1f}  retry-sweeps/lane {st["fails"].mean():.3f}  '
          f'warp-invalid {np.mean((~ok[:nw*32]).reshape(nw,32).any(axis=1)):.2e}')


if __name__ == '__main__':
    M = make_pairs()
    for strat in ('gap', 'newton'):
        run(M, f'C5 d=6 [{strat}]', strategy=strat)
