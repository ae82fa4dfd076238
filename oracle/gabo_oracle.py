"""CPU oracle for the GaBOflow manifold-Gram + GP linear-algebra hot path.

TEST INFRASTRUCTURE ONLY.  This module is a NumPy/SciPy FP64 restatement of the
reference's algorithm.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package ``gaboflow_b200`` never does and fails loudly without its CUDA
library.

Parity status
-------------
* Gram matrices: PINNED.  ``tests/golden/make_golden.py`` imports the reference's
  own NumPy twins (``BoManifolds/Riemannian_utils/SPD_utils.py:57-101,180-197`` and
  ``sphere_utils.py:64-86``) in the build container, evaluates every kernel pair by
  pair on the reference's seeded example inputs, and stores the results under
  ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this oracle against
  them and against the known answers listed in SURVEY.md section 8(c).
* GP log-likelihood / posterior: PARITY UNPINNED.  That arithmetic lives in GPflow 0.5
  (``gpflow/gpr.py``, ``gpflow/densities.py``; pinned only by "GPflow 0.5" in the
  reference ``README.md:16``) on top of TensorFlow 1.x, neither of which is vendored
  in the reference nor installable here.  The functions below restate the published
  algorithm (Rasmussen & Williams Alg. 2.1 as GPflow 0.5 implements it) and are
  anchored on the reference call sites ``examples/kernels/sphere/sphere_kernels.py:131-139``
  and ``examples/kernels/spd/spd_kernels.py:146-170``.

Every function cites the reference file:line it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

_SQRT2 = np.sqrt(2.0)


# --------------------------------------------------------------------------- #
# Mandel notation  (SPD_utils_tf.py:11-92, NumPy twin SPD_utils.py:57-101)
# --------------------------------------------------------------------------- #
def matrix_dim_from_mandel(m: int) -> int:
    """Side d of the symmetric matrix for a Mandel vector of length m.

    Same closed form as ``kernels_spd_tf.py:206`` (``int((-1+sqrt(1+8m))/2)``).
    """
    return int((-1.0 + (1.0 + 8.0 * m) ** 0.5) / 2.0)


def mandel_index(d: int):
    """(row, col, is_offdiag) for each Mandel slot.

    Slot order follows ``SPD_utils_tf.py:27,34-63``: the main diagonal first, then
    super-diagonal k = 1 .. d-1, each in row order; off-diagonal slots carry sqrt(2)
    times the matrix entry.
    """
    rows, cols = [], []
    for k in range(d):
        for r in range(d - k):
            rows.append(r)
            cols.append(r + k)
    rows = np.asarray(rows)
    cols = np.asarray(cols)
    return rows, cols, rows != cols


def mandel_to_sym(V, d: int | None = None):
    """[N,m] Mandel vectors -> [N,d,d] symmetric matrices (``SPD_utils_tf.py:11-65``)."""
    V = np.atleast_2d(np.asarray(V, dtype=np.float64))
    m = V.shape[1]
    if d is None:
        d = matrix_dim_from_mandel(m)
    rows, cols, off = mandel_index(d)
    vals = np.where(off[None, :], V / _SQRT2, V)
    M = np.zeros((V.shape[0], d, d))
    M[:, rows, cols] = vals
    M[:, cols, rows] = vals
    return M


def sym_to_mandel(M):
    """[N,d,d] symmetric -> [N,m] Mandel (``SPD_utils_tf.py:68-92``)."""
    M = np.asarray(M, dtype=np.float64)
    if M.ndim == 2:
        M = M[None]
    d = M.shape[-1]
    rows, cols, off = mandel_index(d)
    V = M[:, rows, cols]
    return np.where(off[None, :], V * _SQRT2, V)


# --------------------------------------------------------------------------- #
# Sphere kernels  (sphere_utils_tf.py:11-60, kernels_sphere_tf.py:59-105,179-221)
# --------------------------------------------------------------------------- #
def sphere_distance(X1, X2, block: int = 4096):
    """All-pairs great-circle distance, ``acos(clamp(x1.x2, -1, 1))``.

    Follows ``sphere_utils_tf.py:40-60`` (the inner product of every pair, clamped,
    then ``acos``) without materialising the N1 x N2 x D tiles: row blocks of
    ``block`` points go through one BLAS product each.
    """
    X1 = np.asarray(X1, dtype=np.float64)
    X2 = np.asarray(X2, dtype=np.float64)
    out = np.empty((X1.shape[0], X2.shape[0]))
    for s in range(0, X1.shape[0], block):
        t = X1[s:s + block] @ X2.T
        np.clip(t, -1.0, 1.0, out=t)
        out[s:s + block] = np.arccos(t)
    return out


def sphere_gaussian_K(X, X2=None, beta=1.0, variance=1.0, block: int = 4096):
    """``variance * exp(-beta * d^2)`` (``kernels_sphere_tf.py:75-88``); beta is the
    effective value ``beta_shifted + beta_min_value`` (:79)."""
    if X2 is None:
        X2 = X
    d = sphere_distance(X, X2, block)
    return variance * np.exp(-(d * d) * beta)


def sphere_laplace_K(X, X2=None, beta=1.0, variance=1.0, block: int = 4096):
    """``variance * exp(-beta * d)`` (``kernels_sphere_tf.py:195-204``)."""
    if X2 is None:
        X2 = X
    d = sphere_distance(X, X2, block)
    return variance * np.exp(-d * beta)


def kdiag_const(X, variance=1.0):
    """``fill([N], variance)`` (``kernels_sphere_tf.py:105,221``;
    ``kernels_spd_tf.py:265,417,550,658``)."""
    return np.full((np.asarray(X).shape[0],), float(variance))


# --------------------------------------------------------------------------- #
# SPD distances and kernels (SPD_utils_tf.py:95-139, kernels_spd_tf.py)
# --------------------------------------------------------------------------- #
def ai_distance_pair_tfroute(S1, S2):
    """One pair, literally the reference TF route ``SPD_utils_tf.py:129-139``:
    ``inv -> sqrtm -> S1^-1/2 S2 S1^-1/2 -> singular values -> sqrt(sum log^2)``.
    Slow path; used to validate the vectorised route below."""
    S1_invhalf = np.real(sla.sqrtm(np.linalg.inv(S1)))
    mult = S1_invhalf @ S2 @ S1_invhalf
    sv = np.linalg.svd(mult, compute_uv=False)
    lg = np.log(sv)
    return np.sqrt(np.sum(lg * lg))


def ai_distance(S1, S2, block: int = 256):
    """All-pairs affine-invariant distance (``SPD_utils_tf.py:95-139``).

    The eigenvalues of ``S1^-1/2 S2 S1^-1/2`` equal those of ``W S2 W^T`` with
    ``W = chol(S1)^-1`` (a congruence by another square root of ``S1^-1``), which is
    symmetric so ``eigvalsh`` applies; this is the backward-stable route SURVEY.md
    section 7 (parity hazards ii) recommends.  Blocked over rows of S1.
    """
    S1 = np.asarray(S1, dtype=np.float64)
    S2 = np.asarray(S2, dtype=np.float64)
    n1, n2 = S1.shape[0], S2.shape[0]
    d = S1.shape[-1]
    L = np.linalg.cholesky(S1)
    eye = np.broadcast_to(np.eye(d), L.shape)
    W = np.stack([sla.solve_triangular(L[i], eye[i], lower=True) for i in range(n1)]) if n1 else L
    out = np.empty((n1, n2))
    for s in range(0, n1, block):
        Wb = W[s:s + block]                                   # [b,d,d]
        M = np.einsum('bij,njk,blk->bnil', Wb, S2, Wb, optimize=True)
        M = 0.5 * (M + np.swapaxes(M, -1, -2))
        lam = np.linalg.eigvalsh(M)
        lg = np.log(lam)
        out[s:s + block] = np.sqrt(np.sum(lg * lg, axis=-1))
    return out


def spd_ai_gaussian_K(X, X2=None, beta=1.0, variance=1.0):
    """``variance * exp(-beta * d_AI^2)`` on Mandel inputs (``kernels_spd_tf.py:229-248``)."""
    S1 = mandel_to_sym(X)
    S2 = S1 if X2 is None else mandel_to_sym(X2)
    d = ai_distance(S1, S2)
    return variance * np.exp(-(d * d) * beta)


def spd_ai_laplace_K(X, X2=None, beta=1.0, variance=1.0):
    """``variance * exp(-beta * d_AI)`` (``kernels_spd_tf.py:383-400``)."""
    S1 = mandel_to_sym(X)
    S2 = S1 if X2 is None else mandel_to_sym(X2)
    d = ai_distance(S1, S2)
    return variance * np.exp(-d * beta)


def spd_stein_K(X, X2=None, beta=1.0, variance=1.0, block: int = 512):
    """Stein kernel AS CODED in the reference (``kernels_spd_tf.py:98-112``):
    ``variance * det(X X2)^beta / det((X+X2)/2)^beta`` -- the exponent on det(X X2) is
    beta, not beta/2, so the diagonal is ``variance * det(X_i)^beta``."""
    S1 = mandel_to_sym(X)
    S2 = S1 if X2 is None else mandel_to_sym(X2)
    out = np.empty((S1.shape[0], S2.shape[0]))
    for s in range(0, S1.shape[0], block):
        A = S1[s:s + block, None]                             # [b,1,d,d]
        B = S2[None]                                          # [1,n,d,d]
        det_mult = np.linalg.det(A @ B)
        det_half = np.linalg.det(0.5 * (A + B))
        out[s:s + block] = np.power(det_mult, beta) / np.power(det_half, beta)
    return variance * out


def stein_kdiag(X, beta=1.0, variance=1.0):
    """``diag_part(K(X))`` (``kernels_spd_tf.py:129``)."""
    S = mandel_to_sym(X)
    det_mult = np.linalg.det(S @ S)
    det_half = np.linalg.det(0.5 * (S + S))
    return variance * np.power(det_mult, beta) / np.power(det_half, beta)


def _pair_sqfrob(A, B, block: int = 1024):
    """``||A_i - B_j||_F^2`` for all pairs, direct difference form
    (``kernels_spd_tf.py:520-529``: subtract, Frobenius norm, square)."""
    n1, n2 = A.shape[0], B.shape[0]
    Af = A.reshape(n1, -1)
    Bf = B.reshape(n2, -1)
    out = np.empty((n1, n2))
    for s in range(0, n1, block):
        diff = Af[s:s + block, None, :] - Bf[None, :, :]
        out[s:s + block] = np.einsum('bnk,bnk->bn', diff, diff)
    return out


def spd_frobenius_K(X, X2=None, beta=1.0, variance=1.0):
    """``variance * exp(-beta * ||X - X2||_F^2)`` (``kernels_spd_tf.py:510-533``)."""
    S1 = mandel_to_sym(X)
    S2 = S1 if X2 is None else mandel_to_sym(X2)
    return variance * np.exp(-_pair_sqfrob(S1, S2) * beta)


def spd_logm(S):
    """FP64 matrix logarithm of each SPD matrix by symmetric eigendecomposition.

    The reference takes ``tf.linalg.logm`` in complex64 (``kernels_spd_tf.py:634``),
    good to ~6e-8 only; its own commented check (``:660-666``) names
    ``scipy.linalg.logm`` in FP64 as the intended answer.  For SPD input the
    principal logarithm equals ``V diag(log w) V^T``.
    """
    w, V = np.linalg.eigh(S)
    return np.einsum('nij,nj,nkj->nik', V, np.log(w), V)


def spd_logeuclid_K(X, X2=None, beta=1.0, variance=1.0):
    """``variance * exp(-beta * ||logm X - logm X2||_F^2)`` (``kernels_spd_tf.py:618-641``),
    evaluated in FP64 (see ``spd_logm``)."""
    L1 = spd_logm(mandel_to_sym(X))
    L2 = L1 if X2 is None else spd_logm(mandel_to_sym(X2))
    return variance * np.exp(-_pair_sqfrob(L1, L2) * beta)


# --------------------------------------------------------------------------- #
# SPD exponential / logarithm maps, beta sweeps  
# --------------------------------------------------------------------------- #
def spd_expmap(U, S):
    """``Expmap_S(U) = S expm(S^-1 U)`` (``SPD_utils_tf.py:142-167``; NumPy twin ``SPD_utils.py:104-120``), evaluated
    literally with ``scipy.linalg.expm`` in FP64.  U [n,d,d] or [d,d]; S [d,d]."""
    U = np.asarray(U, dtype=np.float64)
    if U.ndim == 2:
        return np.real(S @ sla.expm(np.linalg.solve(S, U)))
    return np.stack([np.real(S @ sla.expm(np.linalg.solve(S, u))) for u in U])


def spd_logmap(X, S):
    """``Logmap_S(X) = S logm(S^-1 X)`` (``SPD_utils_tf.py:170-195``; twin ``SPD_utils.py:123-139``) with
    ``scipy.linalg.logm`` in FP64."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 2:
        return np.real(S @ sla.logm(np.linalg.solve(S, X)))
    return np.stack([np.real(S @ sla.logm(np.linalg.solve(S, x))) for x in X])


def min_eig_sweep(K_of_beta, betas):
    """Minimum (and maximum) eigenvalue of the Gram for each beta of a grid, as
    ``examples/kernels/sphere/sphere_gaussian_kernel_parameters.py:83-97`` does (build the Gram, take the extreme eigenvalue;
    the reference calls the non-symmetric ``numpy.linalg.eig`` and keeps the real parts -- ``eigvalsh`` is the symmetric
    solver for the same spectrum).  ``K_of_beta``: callable beta -> [n,n] Gram."""
    lo, hi = [], []
    for b in betas:
        w = np.linalg.eigvalsh(K_of_beta(float(b)))
        lo.append(w[0])
        hi.append(w[-1])
    return np.asarray(lo), np.asarray(hi)


def sweep_trial_sphere(dim=4, nb_samples=500, nb_sources=10, fact_cov=1.0, seed=0):
    """One trial point set of ``sphere_gaussian_kernel_parameters.py:57-77``: samples of ``nb_sources`` isotropic Gaussians
    around random unit means, projected to the unit sphere of R^dim."""
    rng = np.random.RandomState(seed)
    mean = rng.randn(nb_sources, dim)
    mean /= np.linalg.norm(mean, axis=1)[:, None]
    per = nb_samples // nb_sources
    pts = [rng.multivariate_normal(mean[i], fact_cov * np.eye(dim), per) for i in range(nb_sources)]
    X = np.concatenate(pts, axis=0)
    return X / np.linalg.norm(X, axis=1)[:, None]


def sweep_trial_spd(dim=2, nb_samples=500, nb_sources=10, fact_cov=1.0, seed=0):
    """One trial of ``spd_gaussian_kernel_parameters.py:57-106``: tangent-space Gaussian samples mapped to the manifold with
    the exponential map at ``nb_sources`` random base points (Mandel vectors), +1e-6 on the diagonal, ill-conditioned /
    far-away samples removed (norm of the Mandel vector > 1000).  Returns Mandel vectors [n', m]."""
    rng = np.random.RandomState(seed)
    m = dim * (dim + 1) // 2
    eye = np.eye(dim)
    means = [spd_expmap(mandel_to_sym(rng.randn(1, m), dim)[0], eye) + 1e-6 * eye for _ in range(nb_sources)]
    per = nb_samples // nb_sources
    out = []
    for i in range(nb_sources):
        tang = rng.multivariate_normal(np.zeros(m), fact_cov * np.eye(m), per)
        S = spd_expmap(mandel_to_sym(tang, dim), means[i]) + 1e-6 * eye
        out.append(sym_to_mandel(0.5 * (S + np.swapaxes(S, -1, -2))))
    V = np.concatenate(out, axis=0)
    return V[np.linalg.norm(V, axis=1) <= 1000.0]


# --------------------------------------------------------------------------- #
# GP linear algebra  (GPflow 0.5 GPR.build_likelihood / build_predict;  [3P])
# --------------------------------------------------------------------------- #
def gp_cholesky(K, sigma2):
    """``L = chol(K + sigma2 I)``, lower (GPflow 0.5 ``gpr.py`` build_likelihood)."""
    A = np.array(K, dtype=np.float64, copy=True)
    A[np.diag_indices_from(A)] += sigma2
    return sla.cholesky(A, lower=True, overwrite_a=True, check_finite=False)


def gp_loglik(K, y, sigma2, mean=0.0):
    """Log marginal likelihood as GPflow 0.5 ``densities.multivariate_normal``:
    ``-0.5 N P log(2 pi) - P sum(log L_ii) - 0.5 ||L^-1 (y - m)||^2``."""
    y = np.asarray(y, dtype=np.float64).reshape(len(K), -1)
    L = gp_cholesky(K, sigma2)
    alpha = sla.solve_triangular(L, y - mean, lower=True, check_finite=False)
    n, p = y.shape
    return float(-0.5 * n * p * np.log(2.0 * np.pi)
                 - p * np.sum(np.log(np.diag(L)))
                 - 0.5 * np.sum(alpha * alpha))


def gp_predict(K, Kx, Kss, y, sigma2, mean=0.0, full_cov=False):
    """Posterior mean / variance as GPflow 0.5 ``GPR.build_predict``:
    ``A = L^-1 Kx; V = L^-1 (y - m); mean = A^T V + m``;
    ``var = Kdiag(X*) - colsum(A^2)`` or ``cov = K(X*) - A^T A``.
    ``Kss`` is the [M] prior diagonal (full_cov False) or the [M,M] prior covariance."""
    y = np.asarray(y, dtype=np.float64).reshape(len(K), -1)
    L = gp_cholesky(K, sigma2)
    A = sla.solve_triangular(L, Kx, lower=True, check_finite=False)
    V = sla.solve_triangular(L, y - mean, lower=True, check_finite=False)
    fmean = A.T @ V + mean
    if full_cov:
        return fmean, Kss - A.T @ A
    return fmean, (Kss - np.sum(A * A, axis=0))[:, None]


def expected_improvement(mean, var, fmin):
    """GPflowOpt ``ExpectedImprovement.build_acquisition`` [3P]:
    ``(fmin - mu) Phi(z) + var * N(fmin; mu, sqrt(var))`` with ``var`` floored at 1e-6
    (GPflowOpt's ``tf.maximum(candidate_var, stability)``); the pdf is multiplied by
    the variance, not the standard deviation."""
    from scipy.stats import norm
    var = np.maximum(var, 1e-6)
    sd = np.sqrt(var)
    return (fmin - mean) * norm.cdf(fmin, loc=mean, scale=sd) + var * norm.pdf(fmin, loc=mean, scale=sd)


# --------------------------------------------------------------------------- #
# Synthetic workloads of BASELINE.json configs 4 and 5
# --------------------------------------------------------------------------- #
def synth_sphere(n, dim_ambient=51, seed=20240):
    """Uniform points on S^(D-1): normalised standard normals."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, dim_ambient))
    return X / np.linalg.norm(X, axis=1, keepdims=True)


def synth_spd_mandel(n, d=6, seed=20242, wmin=0.5, wmax=2.0):
    """Random SPD matrices ``Q diag(w) Q^T`` with w ~ U[wmin, wmax], Mandel [n, d(d+1)/2];
    the recipe of ``BO_utils/constrained_optimization.py:246-274``."""
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, d, d))
    Q, _ = np.linalg.qr(G)
    w = wmin + (wmax - wmin) * rng.random((n, d))
    S = np.einsum('nij,nj,nkj->nik', Q, w, Q)
    S = 0.5 * (S + np.swapaxes(S, -1, -2))
    return sym_to_mandel(S)
