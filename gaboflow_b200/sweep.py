"""Kernel-parameter selection on the device: extreme eigenvalues of the Gram matrix over a grid of beta.

Device-side form of the reference experiments ``examples/kernels/sphere/sphere_gaussian_kernel_parameters.py:52-104`` and
``examples/kernels/spd/spd_gaussian_kernel_parameters.py:57-125``: for each trial (a random point set) and each beta of a grid
the reference builds the full Gram with ``compute_K_symm`` and takes ``min(numpy.linalg.eig(K))``; ``beta_min`` is then the
smallest beta for which every trial is positive definite.

Here ONE Gram per trial (at the smallest beta of the grid) is computed by the Gram kernels; every other beta is an elementwise
power of it, formed on the fly inside the Lanczos operator (``csrc/eig_sweep.cu``), so the pairwise geometry -- the expensive
part for the affine-invariant kernel -- is never repeated and no second N x N matrix exists.  One CTA per (trial, beta); the
whole grid is one kernel launch.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .param import to_device


def kernel_eigen_sweep(make_kernel, X, betas, m_max=None, tol=0.0, seed=1234):
    """Extreme eigenvalues of ``make_kernel(beta).K(X)`` for every beta in ``betas``.

    make_kernel : callable beta -> kernel instance whose EFFECTIVE beta is ``beta`` (the reference scripts construct
                  ``SphereGaussianKernel(..., beta_min=0., beta=betas[i])``); any kernel of the path of the form
                  variance * exp(-beta * g(x, y)) -- sphere Gaussian / Laplace, SPD affine-invariant Gaussian / Laplace, Frobenius,
                  Log-Euclidean (not Stein, whose 'as coded' entries exceed the variance).
    X           : one point set [n, D] or a stack / list of trials [T, n, D] (NumPy or torch CUDA; n <= 16384).
    betas       : 1-D grid of positive betas.
    m_max       : Lanczos steps (default min(n, 2048); m_max == n gives dense-eigensolver accuracy).
    Returns a dict of NumPy arrays shaped [T, len(betas)] ([len(betas)] for a single point set): ``min_eig``, ``max_eig``,
    ``resid`` (0 where the tridiagonal matrix is exact), ``steps``.
    Entry-wise accuracy of K_beta: relative (beta / min(betas)) * 2e-16 on top of the Gram's own 1e-15.
    """
    betas = np.asarray(betas, dtype=np.float64).reshape(-1)
    if betas.size == 0 or not np.all(betas > 0):
        raise ValueError('betas must be a non-empty grid of positive values')
    single = not isinstance(X, (list, tuple)) and np.ndim(X) == 2
    trials = [X] if single else list(X)
    beta0 = float(betas.min())
    kern = make_kernel(beta0)
    variance = float(kern.variance)
    Gs = None
    for t, Xt in enumerate(trials):
        Xd = to_device(Xt)
        n = int(Xd.shape[0])
        if Gs is None:
            Gs = torch.empty((len(trials), n, n), dtype=torch.float64, device=Xd.device)
        elif n != Gs.shape[1]:
            raise ValueError('every trial must have the same number of points')
        K = kern.K(Xd, out=Gs[t])
        if K.data_ptr() != Gs[t].data_ptr():      # kernel classes without an `out` fast path
            Gs[t].copy_(K)
        ops.log_gram_(Gs[t], variance)
    ratios = torch.from_numpy(betas / beta0).to(Gs.device)
    out = ops.lanczos_sweep(Gs, ratios, variance=variance, m_max=m_max, tol=tol, seed=seed).cpu().numpy()
    res = {'min_eig': out[..., 0], 'max_eig': out[..., 1], 'resid': out[..., 2], 'steps': out[..., 3].astype(np.int64)}
    if single:
        res = {k: v[0] for k, v in res.items()}
    return res


def percentage_pd(min_eig):
    """Fraction of trials whose Gram is positive definite, per beta (sphere_gaussian_kernel_parameters.py:100-104)."""
    m = np.atleast_2d(np.asarray(min_eig))
    return np.sum(m > 0, axis=0) / m.shape[0]


def select_beta_min(betas, min_eig):
    """Smallest beta of the grid from which on EVERY trial is positive definite (the rule stated in the reference scripts'
    header: 'A minimum value of beta is set to the lowest beta value leading to all the kernel matrices to be
    positive-definite').  None if no beta qualifies."""
    betas = np.asarray(betas, dtype=np.float64)
    order = np.argsort(betas)
    ok = (percentage_pd(min_eig) >= 1.0)[order]
    bad = np.nonzero(~ok)[0]
    first = 0 if bad.size == 0 else bad[-1] + 1
    return float(betas[order][first]) if first < betas.size else None
