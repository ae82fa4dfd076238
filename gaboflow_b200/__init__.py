"""gaboflow_b200 -- B200-native (sm_100a) manifold-kernel Gram matrices and GP linear algebra, a drop-in for that
hot path of NoemieJaquier/GaBOflow.  Importing the package does not need a GPU; calling any kernel does, and fails
loudly without the CUDA library (``python -m gaboflow_b200.build``)."""
from . import ops
from .kernel_utils import (SphereGaussianKernel, SphereLaplaceKernel, SpdSteinGaussianKernel,
                           SpdAffineInvariantGaussianKernel, SpdAffineInvariantLaplaceKernel,
                           SpdFrobeniusGaussianKernel, SpdLogEuclideanGaussianKernel)
from .gpr import GPR, Gaussian, Zero, Constant
from .acquisition import ExpectedImprovement

__version__ = '0.1.0'
