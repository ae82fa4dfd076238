"""Manifold kernel classes; same names and constructor signatures as the reference's
``BoManifolds/kernel_utils`` (GPflow-0.5 branch)."""
from .kernels_sphere import SphereGaussianKernel, SphereLaplaceKernel
from .kernels_spd import (SpdSteinGaussianKernel, SpdAffineInvariantGaussianKernel, SpdAffineInvariantLaplaceKernel,
                          SpdFrobeniusGaussianKernel, SpdLogEuclideanGaussianKernel)

__all__ = ['SphereGaussianKernel', 'SphereLaplaceKernel', 'SpdSteinGaussianKernel',
           'SpdAffineInvariantGaussianKernel', 'SpdAffineInvariantLaplaceKernel',
           'SpdFrobeniusGaussianKernel', 'SpdLogEuclideanGaussianKernel']
