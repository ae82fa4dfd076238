"""Expected Improvement on the B200 posterior: the slice of GPflowOpt's ``ExpectedImprovement`` [third party] that the
reference's optimisers consume (``examples/bo_sphere/benchmark_examples/bo_sphere_ackley_manifold.py:145``;
``BoManifolds/BO_utils/manifold_optimization.py:250-271,427`` call ``acquisition.evaluate`` on candidate points).
``evaluate_with_gradients`` (what ``manifold_optimization.py:260-271`` asks GPflowOpt for) is available for the sphere
kernels and the SPD affine-invariant kernels: the Euclidean gradient (Mandel coordinates on S^d_++), which the manifold
optimiser projects onto the tangent space."""
from __future__ import annotations

import torch

from . import ops
from .param import to_device


class ExpectedImprovement:
    def __init__(self, model, fused=True):
        self.model = model
        self.fused = fused          # one-launch evaluation where available (sphere kernels, N <= 4096); same arithmetic
        self.fmin = None
        self.setup()

    def setup(self):
        """fmin = min of the posterior mean at the training inputs (GPflowOpt ``ExpectedImprovement._setup``)."""
        mean, _ = self.model.build_predict(self.model.X, full_cov=False)
        self.fmin = float(mean.min().item())
        return self.fmin

    def _use_fused(self):
        return self.fused and self.model.fused_acquisition_available()

    def build_acquisition(self, Xcand) -> torch.Tensor:
        if self._use_fused():
            return self.model.acquisition_fused(Xcand, self.fmin)[2].unsqueeze(1)
        mean, var = self.model.build_predict(to_device(Xcand), full_cov=False)
        return ops.expected_improvement(mean, var, self.fmin)

    def evaluate(self, Xcand):
        """NumPy in, NumPy out: EI at the candidate points, shape [M, 1]."""
        return self.build_acquisition(Xcand).cpu().numpy()

    def evaluate_with_gradients(self, Xcand):
        """(EI [M, 1], dEI/dx [M, D]) as NumPy arrays (GPflowOpt ``Acquisition.evaluate_with_gradients``); sphere and SPD
        affine-invariant kernels."""
        if self._use_fused():
            out = self.model.acquisition_fused(Xcand, self.fmin, with_grad=True)
            return out[2].unsqueeze(1).cpu().numpy(), out[5].cpu().numpy()
        mean, var, _, _, dei = self.model.build_predict_with_gradients(to_device(Xcand), fmin=self.fmin)
        ei = ops.expected_improvement(mean, var, self.fmin)
        return ei.cpu().numpy(), dei.cpu().numpy()
