"""SPD-manifold kernels on the B200 SPD pair kernels.

Drop-in for ``BoManifolds/kernel_utils/kernels_spd_tf.py`` (GPflow-0.5 branch, :14-658): same class names,
constructor arguments, Param attributes and methods.  Inputs are Mandel vectors [N, d(d+1)/2] (NumPy or torch CUDA).

Structure (instead of the reference's per-PAIR inverse / sqrtm / SVD / logm on N1 x N2 x d x d tensors):
  per point, once  : Mandel -> S, Cholesky whitening W = L^-1, log det, Mandel(logm S)   (gabo_spd_prep_f64)
  per pair         : eigenvalues of W_i S_j W_i^T in registers -- Householder, FP32 QL starts, validated FP64 Newton
                     refinement, FP64 QL / Jacobi when that does not validate -- (affine-invariant), an LDL^T of
                     S_i + S_j (Stein), or the DMMA squared-distance tile kernel (Frobenius / Log-Euclidean).
"""
from __future__ import annotations

import torch

from .. import ops
from ..param import Kern, Param, to_device


def _matrix_dim(input_dim):
    # kernels_spd_tf.py:53,206,358,491,599
    return int((-1.0 + (1.0 + 8.0 * input_dim) ** 0.5) / 2.0)


def _resolve_uplo(X2d, uplo, row0, row1, n):
    if X2d is not None:
        return 'full'
    if uplo == 'mirror' and (row0 != 0 or (row1 is not None and row1 != n)):
        return 'full'
    return uplo


class _SpdPairKernel(Kern):
    """Kernels evaluated by the SPD pair kernel (affine-invariant Gaussian / Laplace, Stein)."""
    _kind = None
    _need = dict(want_S=True, want_W=True, want_logdet=False, want_logm=False)

    def __init__(self, input_dim, active_dims):
        super().__init__(input_dim=input_dim, active_dims=active_dims)
        self.matrix_dim = _matrix_dim(input_dim)

    def _beta_eff(self):
        raise NotImplementedError

    def prepare(self, X, check=True):
        """Per-point preparation; reusable across K calls when X is unchanged (e.g. the training set).  With ``check`` (one
        host sync) a non-SPD input raises ValueError naming the matrix instead of propagating NaN rows into the Gram."""
        prep = ops.spd_prep(to_device(X), self.matrix_dim, **self._need)
        return prep.check() if check else prep

    def K(self, X, X2=None, uplo='mirror', out=None, row0=0, row1=None):
        p1 = X if isinstance(X, ops.SpdPrep) else self.prepare(X)
        p2 = None
        if X2 is not None:
            p2 = X2 if isinstance(X2, ops.SpdPrep) else self.prepare(X2)
        uplo = _resolve_uplo(p2, uplo, row0, row1, p1.n)
        if self.float_type == torch.float32:
            # FP32 mode: FP64 arithmetic inside the pair kernel (see param.Kern.float_type), float32 result
            K = ops.spd_gram(self._kind, p1, p2, beta=self._beta_eff(), variance=float(self.variance),
                             row0=row0, row1=row1, uplo=uplo)
            if out is not None:
                out.copy_(K)
                return out
            return K.to(torch.float32)
        return ops.spd_gram(self._kind, p1, p2, beta=self._beta_eff(), variance=float(self.variance),
                            out=out, row0=row0, row1=row1, uplo=uplo)

    def Kdiag(self, X):
        Xd = to_device(X)
        return ops.kdiag_const(int(Xd.shape[0]), float(self.variance), Xd.device).to(self.float_type)


class SpdSteinGaussianKernel(_SpdPairKernel):
    """Stein kernel exactly as the reference codes it (``kernels_spd_tf.py:14-162``):
    ``variance * det(X X2)^beta / det((X+X2)/2)^beta`` (:104-112), so ``K_ii = variance * det(X_i)^beta``."""
    _kind = ops.SPD_STEIN
    _need = dict(want_S=True, want_W=False, want_logdet=True, want_logm=False)

    def __init__(self, input_dim, active_dims, beta, variance=1.0):
        super().__init__(input_dim=input_dim, active_dims=active_dims)
        # kernels_spd_tf.py:59: beta in {j/2 : 1 <= j <= d-1} U ]0.5(d-1), inf[
        self.continuous_param_space_limit = 0.5 * (self.matrix_dim - 1)
        self.beta_shifted = Param(beta - self.continuous_param_space_limit, transform='positive')   # :67
        self.variance = Param(variance, transform='positive')

    def _beta_eff(self):
        return self.beta_shifted + self.continuous_param_space_limit   # :95

    def Kdiag(self, X):
        """``diag_part(K(X))`` (:129) without forming K: ``variance * det(X_i)^beta``."""
        p = X if isinstance(X, ops.SpdPrep) else self.prepare(X)
        return ops.spd_stein_kdiag(p, self._beta_eff(), float(self.variance)).to(self.float_type)

    def update_beta(self, beta):
        self.beta_shifted = (beta - self.continuous_param_space_limit)   # :144

    def get_beta(self):
        return self.beta_shifted.value + self.continuous_param_space_limit   # :162


class SpdAffineInvariantGaussianKernel(_SpdPairKernel):
    """``variance * exp(-beta * d_AI(X,X2)^2)``, ``beta = beta_shifted + beta_min_value``
    (``kernels_spd_tf.py:165-298``)."""
    _kind = ops.SPD_AI_GAUSSIAN

    def __init__(self, input_dim, active_dims, beta_min, beta=1., variance=1.):
        super().__init__(input_dim=input_dim, active_dims=active_dims)
        self.beta_shifted = Param(beta, transform='positive')     # :210
        self.variance = Param(variance, transform='positive')
        self.beta_min_value = beta_min

    def _beta_eff(self):
        return self.beta_shifted + self.beta_min_value            # :239

    def update_beta(self, beta):
        self.beta_shifted = beta - self.beta_min_value            # :282

    def get_beta(self):
        return self.beta_shifted.value + self.beta_min_value      # :298


class SpdAffineInvariantLaplaceKernel(SpdAffineInvariantGaussianKernel):
    """``variance * exp(-beta * d_AI(X,X2))`` (``kernels_spd_tf.py:319-450``); ``beta_min`` defaults to 0 (:342)."""
    _kind = ops.SPD_AI_LAPLACE

    def __init__(self, input_dim, active_dims, beta_min=0., beta=1., variance=1.):
        super().__init__(input_dim, active_dims, beta_min, beta=beta, variance=variance)


class _SpdSqdistKernel(Kern):
    """Gaussian kernels of a Euclidean distance between per-point feature vectors (DMMA squared-distance tiles)."""

    def __init__(self, input_dim, active_dims, beta=1., variance=1.):
        super().__init__(input_dim=input_dim, active_dims=active_dims)
        self.matrix_dim = _matrix_dim(input_dim)
        self.beta_param = Param(beta, transform='positive')
        self.variance = Param(variance, transform='positive')

    def _features(self, X):
        raise NotImplementedError

    def K(self, X, X2=None, uplo='mirror', out=None, row0=0, row1=None):
        F1 = self._features(X)
        F2 = self._features(X2) if X2 is not None else None
        if self.float_type == torch.float32:       # FP32 mode: features (FP64 logm for Log-Euclid) rounded once, TF32 tcgen05 tiles
            F1 = F1.to(torch.float32)
            F2 = F2.to(torch.float32) if F2 is not None else None
        uplo = _resolve_uplo(F2, uplo, row0, row1, int(F1.shape[0]))
        return ops.points_gram(ops.SQDIST_GAUSSIAN, F1, F2, beta=float(self.beta_param), variance=float(self.variance),
                               out=out, row0=row0, row1=row1, uplo=uplo)

    def Kdiag(self, X):
        Xd = to_device(X)
        return ops.kdiag_const(int(Xd.shape[0]), float(self.variance), Xd.device).to(self.float_type)


class SpdFrobeniusGaussianKernel(_SpdSqdistKernel):
    """``variance * exp(-beta * ||X - X2||_F^2)`` (``kernels_spd_tf.py:453-550``).  The Mandel map is an isometry
    for the Frobenius norm, so the distance is taken on the Mandel vectors directly."""

    def _features(self, X):
        return to_device(X)


class SpdLogEuclideanGaussianKernel(_SpdSqdistKernel):
    """``variance * exp(-beta * ||logm X - logm X2||_F^2)`` (``kernels_spd_tf.py:561-658``), with logm taken once per
    point in FP64 (the reference recomputes it per pair in complex64, :634)."""

    def prepare(self, X, check=True):
        prep = ops.spd_prep(to_device(X), self.matrix_dim, want_S=False, want_W=False, want_logdet=False, want_logm=True)
        return prep.check() if check else prep

    def _features(self, X):
        if isinstance(X, ops.SpdPrep):
            return X.logm
        return self.prepare(X).logm
