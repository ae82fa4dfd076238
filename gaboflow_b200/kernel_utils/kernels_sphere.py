"""Sphere-manifold kernels on the B200 Gram tile kernel.

Drop-in for ``BoManifolds/kernel_utils/kernels_sphere_tf.py`` (GPflow-0.5 branch, :15-221): same class names,
constructor arguments, Param attributes and methods.  ``K`` / ``Kdiag`` take NumPy arrays or torch CUDA tensors and
return torch CUDA tensors (FP64; float32 after ``set_float_type('float32')``: TF32 tcgen05 kernel, 1e-5); ``compute_K*`` return
NumPy like GPflow's AutoFlow wrappers.
"""
from __future__ import annotations

from .. import ops
from ..param import Kern, Param, to_device


class _SphereKernelBase(Kern):
    _kind = None

    def _beta_eff(self) -> float:
        raise NotImplementedError

    def K(self, X, X2=None, uplo='mirror', out=None, row0=0, row1=None):
        """Kernel matrix of X, or between X and X2 (``kernels_sphere_tf.py:59-88,179-204``).

        ``X2 is None`` means ``X2 = X`` (:75-76); the default ``uplo='mirror'`` then returns the full, exactly
        symmetric matrix while computing each off-diagonal tile once.  ``uplo='lower'`` writes only the tiles the
        Cholesky reads; ``row0/row1`` select a row block (multi-GPU sharding, multiples of 128).
        """
        Xd = to_device(X, self.float_type)
        X2d = to_device(X2, self.float_type) if X2 is not None else None
        if X2d is None and (row0 != 0 or (row1 is not None and row1 != Xd.shape[0])) and uplo == 'mirror':
            uplo = 'full'
        if X2d is not None:
            uplo = 'full'
        return ops.points_gram(self._kind, Xd, X2d, beta=self._beta_eff(), variance=float(self.variance),
                               out=out, row0=row0, row1=row1, uplo=uplo)

    def Kdiag(self, X):
        """``fill([N], variance)`` (``kernels_sphere_tf.py:90-105, 206-221``)."""
        Xd = to_device(X, self.float_type)
        return ops.kdiag_const(int(Xd.shape[0]), float(self.variance), Xd.device).to(self.float_type)


class SphereGaussianKernel(_SphereKernelBase):
    """``variance * exp(-beta * d_S(x,y)^2)`` with ``beta = beta_shifted + beta_min_value``
    (``kernels_sphere_tf.py:15-138``)."""
    _kind = ops.SPHERE_GAUSSIAN

    def __init__(self, input_dim, active_dims, beta_min, beta=1., variance=1.):
        super().__init__(input_dim=input_dim, active_dims=active_dims)
        # kernels_sphere_tf.py:55-57: the Param holds beta as given; beta_min is ADDED in K (:79)
        self.beta_shifted = Param(beta, transform='positive')
        self.variance = Param(variance, transform='positive')
        self.beta_min_value = beta_min

    def _beta_eff(self):
        return self.beta_shifted + self.beta_min_value

    def update_beta(self, beta):
        """``beta_shifted = beta - beta_min_value`` (``kernels_sphere_tf.py:107-122``)."""
        self.beta_shifted = beta - self.beta_min_value

    def get_beta(self):
        """``beta_shifted.value + beta_min_value`` (``kernels_sphere_tf.py:124-138``)."""
        return self.beta_shifted.value + self.beta_min_value


class SphereLaplaceKernel(_SphereKernelBase):
    """``variance * exp(-beta * d_S(x,y))`` (``kernels_sphere_tf.py:141-221``)."""
    _kind = ops.SPHERE_LAPLACE

    def __init__(self, input_dim, active_dims, beta=1., variance=1.):
        super().__init__(input_dim=input_dim, active_dims=active_dims)
        self.beta = Param(beta, transform='positive')
        self.variance = Param(variance, transform='positive')

    def _beta_eff(self):
        return float(self.beta)
