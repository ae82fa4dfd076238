"""Minimal stand-ins for the GPflow 0.5 pieces the reference kernel classes lean on [3P]:
``gpflow.param.Param`` (a named, positive-transformed scalar with ``.value``) and the ``gpflow.kernels.Kern`` base
(``input_dim``, ``active_dims``, the AutoFlow wrappers ``compute_K`` / ``compute_K_symm`` / ``compute_Kdiag`` that
the reference examples call with NumPy arrays, e.g. ``examples/kernels/sphere/sphere_kernels.py:127-129``).

Only host logic lives here; all arithmetic is in the CUDA library.
"""
from __future__ import annotations

import numpy as np
import torch


class Param:
    """Scalar hyper-parameter.  ``.value`` returns a fresh 1-element float64 array like GPflow 0.5's
    ``Param.value`` (which is why the reference's ``get_beta()`` returns an array, ``kernels_sphere_tf.py:138``)."""

    def __init__(self, value, transform='positive'):
        self.transform = transform
        self._array = np.atleast_1d(np.asarray(value, dtype=np.float64)).copy()
        self.fixed = False

    @property
    def value(self):
        return self._array.copy()

    def assign(self, value):
        self._array = np.atleast_1d(np.asarray(value, dtype=np.float64)).copy()

    def __float__(self):
        return float(self._array.reshape(-1)[0])

    # arithmetic used in the kernel bodies (``self.beta_shifted + self.beta_min_value``)
    def __add__(self, other):
        return float(self) + float(other)

    __radd__ = __add__

    def __sub__(self, other):
        return float(self) - float(other)

    def __rsub__(self, other):
        return float(other) - float(self)

    def __mul__(self, other):
        return float(self) * float(other)

    __rmul__ = __mul__

    def __repr__(self):
        return f'Param({float(self)!r}, transform={self.transform!r})'


class Parameterized:
    """Assigning a number to an attribute that already holds a Param updates the Param in place, as GPflow 0.5's
    ``Parameterized.__setattr__`` does (the reference's ``update_beta`` relies on it: ``kernels_sphere_tf.py:122``)."""

    def __setattr__(self, key, value):
        cur = self.__dict__.get(key)
        if isinstance(cur, Param) and not isinstance(value, Param):
            cur.assign(value)
        else:
            object.__setattr__(self, key, value)

    def params(self):
        return {k: v for k, v in self.__dict__.items() if isinstance(v, Param)}


def _cuda_device():
    if not torch.cuda.is_available():
        raise RuntimeError('gaboflow_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
    return torch.device('cuda', torch.cuda.current_device())


def to_device(X, dtype=torch.float64):
    """NumPy array / list / CPU tensor -> CUDA tensor; CUDA tensors pass through untouched (zero-copy)."""
    if isinstance(X, torch.Tensor):
        if X.is_cuda:
            return X if X.dtype == dtype else X.to(dtype)
        return X.to(device=_cuda_device(), dtype=dtype, non_blocking=True)
    arr = np.ascontiguousarray(np.asarray(X, dtype=np.float64 if dtype == torch.float64 else np.float32))
    if arr.ndim == 1:
        arr = arr[None, :]
    return torch.from_numpy(arr).to(device=_cuda_device(), non_blocking=True)


class Kern(Parameterized):
    """Base of the manifold kernels: the slice of ``gpflow.kernels.Kern`` (0.5) that the reference uses."""

    # GPflow's ``settings.float_type``, per kernel object: torch.float64 (default) or torch.float32 ("FP32 mode", north_star
    # tolerance 1e-5): K / Kdiag then take and return float32 tensors.  Sphere and squared-distance kernels run the TF32
    # tcgen05 Gram kernel; the SPD pair kernels (affine-invariant, Stein) keep FP64 arithmetic inside -- a 6 x 6 eigen-solve in
    # FP32 does not hold 1e-5 on ill-conditioned pairs -- and store float32.  The GP algebra (GPR) is FP64 only.
    float_type = torch.float64

    def __init__(self, input_dim, active_dims=None):
        self.input_dim = int(input_dim)
        # accepted and stored, never applied -- same as the reference, whose K() bodies do not call _slice
        self.active_dims = list(active_dims) if active_dims is not None else list(range(self.input_dim))

    def set_float_type(self, float_type):
        """'float64' / 'float32' (or the torch / NumPy dtype); returns self."""
        name = str(float_type).replace('torch.', '').replace("<class 'numpy.", '').replace("'>", '')
        if name not in ('float64', 'float32'):
            raise ValueError('float_type must be float64 or float32')
        self.__dict__['float_type'] = torch.float64 if name == 'float64' else torch.float32
        return self

    # subclasses implement K(X, X2=None) and Kdiag(X) on device tensors
    def K(self, X, X2=None):
        raise NotImplementedError

    def Kdiag(self, X):
        raise NotImplementedError

    # analytic hyper-parameter gradients (gaboflow_b200.gpr): every kernel of the path has the form
    # variance * exp(beta_eff * g(x, x')) with beta_eff = <beta Param> + constant shift, which is all GPR needs to know
    def beta_param_name(self):
        for name in ('beta_shifted', 'beta', 'beta_param'):
            if isinstance(self.__dict__.get(name), Param):
                return name
        return None

    def beta_effective(self) -> float:
        if hasattr(self, '_beta_eff'):
            return float(self._beta_eff())
        return float(self.__dict__[self.beta_param_name()])

    # AutoFlow-style wrappers: NumPy in, NumPy out (host<->device copies included)
    def compute_K(self, X, Z):
        return self.K(X, Z).cpu().numpy()

    def compute_K_symm(self, X):
        return self.K(X).cpu().numpy()

    def compute_Kdiag(self, X):
        return self.Kdiag(X).cpu().numpy()
