"""Batched manifold utilities on the B200 kernels -- the steps directly before the Gram path.

Drop-in names for ``BoManifolds/Riemannian_utils/SPD_utils_tf.py`` (``vector_to_symmetric_matrix_tf`` :11-65,
``symmetric_matrix_to_vector_tf`` :68-92, ``affine_invariant_distance_tf`` :95-139, ``expmap_tf`` :142-167, ``logmap_tf``
:170-195), their NumPy / Mandel twins (``SPD_utils.py:57-177``) and ``sphere_utils_tf.sphere_distance_tf`` (:11-60).
Inputs: torch CUDA FP64 tensors or NumPy arrays (copied to the device); outputs: torch tensors on the device.
All arithmetic happens in libgabo_b200.so; there is no CPU path.
"""
from __future__ import annotations

import torch

from . import ops
from .param import to_device


def vector_to_symmetric_matrix(vect, dim):
    """Mandel vectors [N, m] (or [m]) -> symmetric matrices [N, dim, dim] (or [dim, dim]); SPD_utils_tf.py:11-65."""
    v = to_device(vect)
    single = v.dim() == 1
    p = ops.spd_prep(v.reshape(-1, v.shape[-1]), dim, want_S=True, want_W=False, want_logdet=False, want_logm=False, want_Lt=False)
    S = p.S.reshape(-1, dim, dim)
    return S[0] if single else S


def symmetric_matrix_to_vector(mat, dim):
    """Symmetric matrices [N, dim, dim] (or [dim, dim]) -> Mandel vectors; SPD_utils_tf.py:68-92."""
    return ops.sym_to_mandel(to_device(mat), dim)


vector_to_symmetric_matrix_tf = vector_to_symmetric_matrix
symmetric_matrix_to_vector_tf = symmetric_matrix_to_vector
vector_to_symmetric_matrix_mandel = vector_to_symmetric_matrix      # SPD_utils.py:79-101
symmetric_matrix_to_vector_mandel = symmetric_matrix_to_vector      # SPD_utils.py:57-76


def expmap(U, S):
    """Expmap_S(U) = S expm(S^-1 U) for U [N, d, d] or [d, d], S [d, d] (or one base per input); SPD_utils_tf.py:142-167."""
    U, S = to_device(U), to_device(S)
    return ops.spd_expmap(U, S, int(U.shape[-1]))


def logmap(X, S):
    """Logmap_S(X) = S logm(S^-1 X); SPD_utils_tf.py:170-195."""
    X, S = to_device(X), to_device(S)
    return ops.spd_logmap(X, S, int(X.shape[-1]))


expmap_tf, logmap_tf = expmap, logmap


def _dim_from_mandel(m):
    return int((-1.0 + (1.0 + 8.0 * m) ** 0.5) / 2.0)


def expmap_mandel_vector(u, s):
    """Mandel in, Mandel out (SPD_utils.py:142-158)."""
    u, s = to_device(u), to_device(s)
    return ops.spd_expmap(u, s, _dim_from_mandel(int(u.shape[-1])), mandel=True)


def logmap_mandel_vector(x, s):
    """Mandel in, Mandel out (SPD_utils.py:161-177)."""
    x, s = to_device(x), to_device(s)
    return ops.spd_logmap(x, s, _dim_from_mandel(int(x.shape[-1])), mandel=True)


def affine_invariant_distance(S1, S2, full_dist_mat=False):
    """d_AI between SPD matrices [N1, d, d] and [N2, d, d] (SPD_utils_tf.py:95-139).  full_dist_mat=True: the [N1, N2] matrix
    (what the kernels use); False: the N1 == N2 paired distances.  Evaluated by the affine-invariant pair kernel (Laplace
    kind, beta = 1: K = exp(-d)) followed by -log on the device: absolute error ~1e-16 * (1 + d)."""
    S1, S2 = to_device(S1), to_device(S2)
    d = int(S1.shape[-1])
    A = S1.reshape(-1, d, d)
    B = S2.reshape(-1, d, d)
    p1 = ops.spd_prep(ops.sym_to_mandel(A, d), d, want_S=False, want_W=True, want_logdet=False, want_Lt=False)
    p2 = ops.spd_prep(ops.sym_to_mandel(B, d), d, want_S=False, want_W=False, want_logdet=False, want_Lt=True)
    p1.check()
    p2.check()
    K = ops.spd_gram(ops.SPD_AI_LAPLACE, p1, p2, beta=1.0, variance=1.0, uplo='full')
    D = ops.log_gram_(K, 1.0, scale=-1.0)
    if full_dist_mat:
        return D
    if A.shape[0] != B.shape[0]:
        raise ValueError('paired distances need as many matrices in S1 as in S2')
    return torch.diagonal(D).clone()


affine_invariant_distance_tf = affine_invariant_distance


def sphere_distance(x1, x2, diag=False):
    """Great-circle distances acos(clamp(x1 . x2)) (sphere_utils_tf.py:11-60): the [N1, N2] matrix, or with diag=True the
    N1 == N2 paired distances.  Evaluated by the sphere Laplace Gram (beta = 1) followed by -log on the device."""
    x1, x2 = to_device(x1), to_device(x2)
    K = ops.points_gram(ops.SPHERE_LAPLACE, x1.reshape(-1, x1.shape[-1]), x2.reshape(-1, x2.shape[-1]), beta=1.0, variance=1.0)
    D = ops.log_gram_(K, 1.0, scale=-1.0)
    return torch.diagonal(D).clone() if diag else D


sphere_distance_tf = sphere_distance
