"""GP regression on the B200 linear-algebra kernels: the slice of GPflow 0.5's ``gpflow.gpr.GPR`` [3P] that the
reference drives (``examples/kernels/sphere/sphere_kernels.py:131-139``, ``examples/kernels/spd/spd_kernels.py:146-170``,
every BO iteration through GPflowOpt).

    K  = kern.K(X) + sigma^2 I         -> Gram tile kernel, lower tiles only
    L  = chol(K)                       -> gabo_potrf_f64 (sigma^2 added in-kernel)
    V  = L^-1 (Y - m(X))               -> gabo_trsm_f64
    ll = -N P/2 log 2pi - P sum log L_ii - 1/2 ||V||^2          -> gabo_gp_loglik_f64
    A  = L^-1 K(X, X*) ; mean = A^T V + m ; var = Kdiag(X*) - colsum(A^2) ; cov = K(X*) - A^T A

GPflow 0.5 refactorises on every call; here the factor is cached and reused until a hyper-parameter or the data
changes (same results).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .param import Param, Parameterized, to_device


class Gaussian(Parameterized):
    """``gpflow.likelihoods.Gaussian``: noise variance sigma^2, default 1.0."""

    def __init__(self, variance=1.0):
        self.variance = Param(variance, transform='positive')


class Zero:
    """``gpflow.mean_functions.Zero``."""
    c = 0.0


class Constant(Parameterized):
    """``gpflow.mean_functions.Constant`` with a scalar constant."""

    def __init__(self, c=0.0):
        self.c = Param(c, transform=None)


class GPR(Parameterized):
    def __init__(self, X, Y, kern, mean_function=None):
        self.X = to_device(X)
        Yd = to_device(Y)
        if Yd.dim() == 1:
            Yd = Yd.unsqueeze(1)
        if Yd.shape[0] != self.X.shape[0]:
            Yd = Yd.t().contiguous()
        self.Y = Yd.contiguous()
        if getattr(kern, 'float_type', torch.float64) != torch.float64:
            raise ValueError('GPR factors K + sigma^2 I in FP64: the kernel must be in float64 mode (kern.set_float_type("float64"))')
        self.kern = kern
        self.mean_function = mean_function if mean_function is not None else Zero()
        self.likelihood = Gaussian()
        self._cache_key = None
        self._L = None
        self._V = None

    def set_data(self, X, Y) -> None:
        """Replace the training set (what GPflow 0.5 does by assigning ``m.X`` / ``m.Y``) and drop every cached quantity."""
        self.X = to_device(X)
        Yd = to_device(Y)
        if Yd.dim() == 1:
            Yd = Yd.unsqueeze(1)
        if Yd.shape[0] != self.X.shape[0]:
            Yd = Yd.t().contiguous()
        self.Y = Yd.contiguous()
        self._cache_key = None
        self._V = None
        self._xprep_tag = None

    # ---- state ------------------------------------------------------------------------------------
    def _mean_const(self) -> float:
        return float(self.mean_function.c)

    def _key(self):
        kp = tuple(sorted((k, float(v)) for k, v in self.kern.params().items()))
        extra = tuple(sorted((k, v) for k, v in self.kern.__dict__.items() if isinstance(v, (int, float))))
        return (kp, extra, float(self.likelihood.variance), self._mean_const(), self.X.data_ptr(), self.X._version,
                self.Y.data_ptr(), self.Y._version)

    def _kern_input(self):
        """What goes into ``kern.K`` for the training set: X itself, or -- for kernels with a per-point preparation (the SPD
        pair kernels) -- that preparation, computed once per training set instead of on every ``K(X, X*)``."""
        prepare = getattr(self.kern, 'prepare', None)
        if prepare is None:
            return self.X
        tag = (self.X.data_ptr(), self.X._version, tuple(self.X.shape))
        if getattr(self, '_xprep_tag', None) != tag:
            self._xprep = prepare(self.X)             # raises ValueError on a non-SPD input (ops.SpdPrep.check)
            self._xprep_tag = tag
        return self._xprep

    def _factorise(self):
        key = self._key()
        if self._cache_key == key and self._L is not None:
            return self._L, self._V
        # the buffer is about to be overwritten: whatever happens below, the old factor is gone (a failed Cholesky must not
        # leave the previous key pointing at a half-factored matrix)
        self._cache_key = None
        self._V = None
        n = int(self.X.shape[0])
        if self._L is None or self._L.shape[0] != n:
            self._L = self._factor_view(n)
        self.kern.K(self._kern_input(), uplo='lower', out=self._L)
        ops.potrf_(self._L, float(self.likelihood.variance), check=True)
        V = (self.Y - self._mean_const()).contiguous()
        ops.trsm_(self._L, V, trans=False)
        self._V = V
        self._cache_key = key
        return self._L, self._V

    def _factor_view(self, n: int) -> torch.Tensor:
        """[n, n] view of the factor buffer, which is allocated with head-room so that ``append`` grows in place."""
        buf = getattr(self, '_Lbuf', None)
        if buf is None or buf.shape[0] < n:
            cap = max(256, ((n + n // 8 + 255) // 256) * 256)
            new = torch.empty((cap, cap), dtype=torch.float64, device=self.X.device)
            if buf is not None and self._L is not None and self._L.shape[0] <= n:
                m = int(self._L.shape[0])
                new[:m, :m].copy_(self._L)
            self._Lbuf = buf = new
        return buf[:n, :n]

    def append(self, x_new, y_new) -> None:
        """Add ONE observation without refactorising (the reference rebuilds the whole model per BO
        iteration, ``examples/bo_sphere/...``).  With Ky' = [[Ky, k], [k^T, k** + s2]] the factor gains one row:

            l = L^-1 k  (gabo_trsm_f64, one right-hand side),   d = sqrt(k** + s2 - l^T l),   L' = [[L, 0], [l^T, d]]

        and V' = L'^-1 (Y' - m) gains the entry (y_new - m - l^T V) / d.  O(N^2) instead of O(N^3); hyper-parameters are
        unchanged, so the cached factor stays valid.  Raises LinAlgError if the extended matrix is not positive definite."""
        L, V = self._factorise()
        n, P = int(V.shape[0]), int(V.shape[1])
        xn = to_device(x_new).reshape(1, -1)
        yn = to_device(y_new).reshape(1, P)
        k = self.kern.K(self._kern_input(), xn).contiguous()          # [n, 1]
        # k(x*, x*): the kernels of the path with a constant diagonal report it without a device read
        kss = self._kdiag_scalar(xn) + float(self.likelihood.variance)
        ops.trsm_(L, k, trans=False)                                  # l
        Lnew = self._factor_view(n + 1)                               # may move the factor into a larger buffer
        vnew = torch.empty((1, P), dtype=torch.float64, device=V.device)
        # d, the new factor row and the new entry of V in one launch (gabo_gp_append_row_f64); one host read: the info word
        info = ops.gp_append_row(k, kss, V, yn, self._mean_const(), Lnew[n], vnew)
        if int(info.item()) != 0:
            raise np.linalg.LinAlgError('append: the extended covariance is not positive definite')
        self.X = torch.cat([self.X, xn], dim=0)
        self.Y = torch.cat([self.Y, yn], dim=0)
        self._L = Lnew
        self._V = torch.cat([V, vnew], dim=0)
        self._cache_key = self._key()

    def _kdiag_scalar(self, xn) -> float:
        """k(x, x) of one point: `variance` for every kernel of the path except Stein (det(X)^beta, one device read)."""
        from .kernel_utils.kernels_spd import SpdSteinGaussianKernel
        if isinstance(self.kern, SpdSteinGaussianKernel):
            return float(self.kern.Kdiag(xn)[0].item())
        return float(self.kern.variance)

    # ---- device-level graph pieces (GPflow: build_likelihood / build_predict) --------------------------
    def build_likelihood(self) -> torch.Tensor:
        L, V = self._factorise()
        return ops.gp_loglik(L, V)

    def build_predict(self, Xnew, full_cov=False):
        L, V = self._factorise()
        Xn = to_device(Xnew)
        A = self.kern.K(self._kern_input(), Xn)        # [N, M]
        ops.trsm_(L, A, trans=False)
        kd = self.kern.Kdiag(Xn)
        mean, var = ops.gp_predict(A, V, kd, self._mean_const())
        if not full_cov:
            p = int(V.shape[1])
            return mean, var.unsqueeze(1).expand(-1, p).contiguous()
        cov = self.kern.K(Xn)                          # [M, M]
        At = A.t().contiguous()                        # [M, N]
        ops.gemm_nt_sub_(cov, At, At, lower=False)
        p = int(V.shape[1])
        return mean, cov.unsqueeze(2).expand(-1, -1, p).contiguous()

    def loglik_gradients(self) -> dict:
        """Analytic derivatives of the log marginal likelihood w.r.t. the VALUES of the hyper-parameters (SURVEY section
        8f-1; the reference obtains them by TensorFlow autodiff through ``GPR.build_likelihood``):

            d ll / d theta = 1/2 sum_c alpha_c^T dK alpha_c - P/2 tr(Ky^-1 dK),   alpha = Ky^-1 (Y - m),  Ky = K + s2 I

        noise s2: dK = I;  kernel variance v: dK = K/v, which needs only tr(Ky^-1) because K = Ky - s2 I;  beta: every
        kernel of the path is v exp(beta_eff g), so dK = K log(K/v)/beta_eff, reduced against alpha alpha^T - P Ky^-1 in one
        pass (``gabo_gp_grad_beta_f64``);  constant mean c: sum(alpha).  Ky^-1 = L^-T L^-1 comes from the cached factor
        (triangular solve of the identity on the DMMA core, then one symmetric product).  Keys: 'likelihood.variance',
        'kern.variance', 'kern.<beta name>', 'mean_function.c'."""
        L, V = self._factorise()
        n, P = int(V.shape[0]), int(V.shape[1])
        dev = L.device
        alpha = V.clone()
        ops.trsm_(L, alpha, trans=True)
        Z = torch.eye(n, dtype=torch.float64, device=dev)
        ops.trsm_(L, Z, trans=False)                                  # L^-1
        Zt = Z.t().contiguous()
        neg_kinv = Z.zero_()                                          # reuse the buffer: lower triangle of -Ky^-1
        ops.gemm_nt_sub_(neg_kinv, Zt, Zt, lower=True)
        resid = self.Y - self._mean_const()
        # the four scalar reductions come back in ONE device -> host read
        tr_neg, aa, ay, asum = torch.stack([neg_kinv.diagonal().sum(), (alpha * alpha).sum(), (alpha * resid).sum(),
                                            alpha.sum()]).tolist()
        tr_kinv = -tr_neg
        s2 = float(self.likelihood.variance)
        v = float(self.kern.variance)
        grads = {'likelihood.variance': 0.5 * (aa - P * tr_kinv),
                 'kern.variance': 0.5 * ((ay - s2 * aa) - P * (n - s2 * tr_kinv)) / v,
                 'mean_function.c': asum}
        bname = self.kern.beta_param_name()
        if bname is not None:
            Kf = Zt                                                    # reuse: noise-free Gram, lower tiles
            self.kern.K(self._kern_input(), uplo='lower', out=Kf)
            grads['kern.' + bname] = float(ops.gp_grad_beta(Kf, neg_kinv, alpha, v, self.kern.beta_effective()).item())
        return grads

    def build_predict_with_gradients(self, Xnew, fmin=None):
        """Posterior mean / variance at the candidates AND their Euclidean gradients w.r.t. the candidates, for the sphere
        kernels and the SPD affine-invariant kernels (Mandel coordinates) (the reference differentiates through TensorFlow and hands the result to the manifold
        optimiser, ``BO_utils/manifold_optimization.py:260-271``).  Returns (mean [M,1], var [M], dmean [M,D], dvar [M,D],
        dEI [M,D] or None); with ``fmin`` the EI gradient is formed in the same kernel.  One output column only."""
        from .kernel_utils.kernels_spd import SpdAffineInvariantGaussianKernel
        kind = getattr(self.kern, '_kind', None)
        is_ai = isinstance(self.kern, SpdAffineInvariantGaussianKernel)          # Gaussian and (subclass) Laplace
        if not is_ai and kind not in (ops.SPHERE_GAUSSIAN, ops.SPHERE_LAPLACE):
            raise NotImplementedError('candidate gradients are implemented for the sphere and the SPD affine-invariant kernels')
        L, V = self._factorise()
        if int(V.shape[1]) != 1:
            raise NotImplementedError('candidate gradients need a single output column')
        Xn = to_device(Xnew)
        if is_ai:
            ptrain = self._kern_input()
            pcand = self.kern.prepare(Xn)
            Kx = self.kern.K(ptrain, pcand)                # [N, M]
            A = Kx.clone()
            ops.trsm_(L, A, trans=False)
            mean, var = ops.gp_predict(A, V, self.kern.Kdiag(Xn), self._mean_const())
            ops.trsm_(L, A, trans=True)                    # Ky^-1 Kx
            alpha = V.clone()
            ops.trsm_(L, alpha, trans=True)
            dmean, dvar, dei = ops.spd_ai_posterior_grad(kind, ptrain, pcand, alpha, A, self.kern.beta_effective(),
                                                         float(self.kern.variance), mean, var, fmin)
            return mean, var, dmean, dvar, dei
        Kx = self.kern.K(self.X, Xn)                       # [N, M]
        A = Kx.clone()
        ops.trsm_(L, A, trans=False)                       # L^-1 Kx
        mean, var = ops.gp_predict(A, V, self.kern.Kdiag(Xn), self._mean_const())
        ops.trsm_(L, A, trans=True)                        # Ky^-1 Kx
        alpha = V.clone()
        ops.trsm_(L, alpha, trans=True)
        dmean, dvar, dei = ops.sphere_posterior_grad(kind, self.X, Kx, alpha, A, self.kern.beta_effective(),
                                                     float(self.kern.variance), mean, var, fmin)
        return mean, var, dmean, dvar, dei

    # ---- fused acquisition ---------------------------------------------------------------------
    FUSED_MAX_N = 4096

    def fused_acquisition_available(self) -> bool:
        return (getattr(self.kern, '_kind', None) in (ops.SPHERE_GAUSSIAN, ops.SPHERE_LAPLACE)
                and hasattr(self.kern, 'beta_effective') and type(self.kern).__module__.endswith('kernels_sphere')
                and int(self.X.shape[0]) <= self.FUSED_MAX_N and int(self.Y.shape[1]) == 1 and int(self.X.shape[1]) <= 64)

    def _fused_state(self):
        """L^-1 and alpha = Ky^-1 (y - m), formed once per factorisation and kept beside the cached factor."""
        L, V = self._factorise()
        if getattr(self, '_fused_key', None) != self._cache_key:
            n = int(L.shape[0])
            Linv = torch.eye(n, dtype=torch.float64, device=L.device)
            ops.trsm_(L, Linv, trans=False)
            alpha = V.clone()
            ops.trsm_(L, alpha, trans=True)
            self._Linv, self._alpha, self._fused_key, self._fused_out = Linv, alpha.reshape(-1).contiguous(), self._cache_key, {}
        return self._Linv, V, self._alpha

    def acquisition_fused(self, Xcand, fmin, with_grad=False):
        """Posterior mean / variance, Expected Improvement and (optionally) their Euclidean gradients at the candidates in ONE
        kernel launch (``gabo_sphere_acq_fused_f64``) -- the evaluation ``manifold_optimization.py:250-271,427`` repeats thousands
        of times per BO iteration.  Sphere kernels, N <= 4096.  Returns (mean [M], var [M], ei [M], dmean, dvar, dei)."""
        if not self.fused_acquisition_available():
            raise NotImplementedError('fused acquisition: sphere kernels, one output column, N <= %d' % self.FUSED_MAX_N)
        Linv, V, alpha = self._fused_state()
        return ops.sphere_acq_fused(self.kern._kind, self.X, to_device(Xcand), Linv, V.reshape(-1), alpha, self.kern.beta_effective(),
                                    float(self.kern.variance), self._mean_const(), float(fmin), with_grad=with_grad,
                                    out=self._fused_out)

    # ---- small-N fused path (BASELINE configs 1-3: N = 5..160) ------------------------------------------------------------
    def small_path_available(self) -> bool:
        """One-launch evaluation (``gabo_gp_small_batch_f64``): N <= 160 points, <= 8 output columns, and a kernel of the form
        variance * exp(-beta g) with g(x, x) = 0 -- every kernel of the path except Stein 'as coded'."""
        from .kernel_utils.kernels_spd import SpdSteinGaussianKernel
        n, p = int(self.X.shape[0]), int(self.Y.shape[1])
        if isinstance(self.kern, SpdSteinGaussianKernel) or self.kern.beta_param_name() is None:
            return False
        return 1 <= n <= ops.GP_SMALL_MAX_N and 1 <= p <= ops.GP_SMALL_MAX_P and (n + p) * (n | 1) + (p + 2) * n + 32 <= 227 * 128

    def _log_gram(self):
        """G = log(K_beta0 / variance) of the training set, computed ONCE per training set (beta0 = the kernel's effective beta
        at that moment): every other beta is the elementwise power exp((beta/beta0) G) formed inside the fused kernel, so the
        pairwise geometry -- acos, or a generalised eigenproblem per pair -- is not repeated during hyper-parameter search."""
        tag = (self.X.data_ptr(), self.X._version, tuple(self.X.shape), id(self.kern))
        if getattr(self, '_g_tag', None) != tag:
            beta0 = float(self.kern.beta_effective())
            K = self.kern.K(self._kern_input()).clone()
            ops.log_gram_(K, float(self.kern.variance))
            if not bool(torch.isfinite(K).all().item()):
                raise ValueError('log-Gram is not finite: entries of K underflow at beta = %g; start from a smaller beta' % beta0)
            self._G, self._G_beta0, self._g_tag = K, beta0, tag
        return self._G, self._G_beta0

    def log_likelihood_batch(self, beta=None, variance=None, noise_variance=None, mean_const=None, with_grad=True) -> dict:
        """Log marginal likelihood (and its gradients) for B hyper-parameter sets in ONE kernel launch -- multi-start, grid or
        line-search candidates of the hyper-parameter fit the reference runs per BO iteration (``GPR.optimize``,
        ``examples/kernels/sphere/sphere_kernels.py:133``).  Each argument is a scalar or an array of length B (EFFECTIVE beta,
        kernel variance, noise variance, constant mean); omitted ones keep the model's current value.  Returns NumPy arrays:
        ``loglik`` [B], ``info`` [B] (0, or the index of the failing Cholesky pivot: loglik is NaN there), and with ``with_grad``
        ``grad`` = {'kern.<beta name>', 'kern.variance', 'likelihood.variance', 'mean_function.c'} -> [B] (derivatives w.r.t. the
        parameter VALUES, as ``loglik_gradients``)."""
        if not self.small_path_available():
            raise NotImplementedError('log_likelihood_batch: N <= %d, <= %d output columns, non-Stein kernels'
                                      % (ops.GP_SMALL_MAX_N, ops.GP_SMALL_MAX_P))
        G, beta0 = self._log_gram()
        cur = [self.kern.beta_effective(), float(self.kern.variance), float(self.likelihood.variance), self._mean_const()]
        cols = [np.atleast_1d(np.asarray(c if a is None else a, dtype=np.float64)) for a, c in
                zip((beta, variance, noise_variance, mean_const), cur)]
        B = max(c.size for c in cols)
        theta = np.empty((B, 4), dtype=np.float64)
        for k, c in enumerate(cols):
            if c.size not in (1, B):
                raise ValueError('hyper-parameter arrays must have one common length')
            theta[:, k] = c
        theta[:, 0] /= beta0
        out, info = ops.gp_small_batch(G, self.Y, torch.from_numpy(theta).to(G.device), want_grad=with_grad)
        o = out.cpu().numpy()
        # a failed set has NaN outputs, so the info words cost a second device->host read only when something failed
        res = {'loglik': o[:, 0].copy(),
               'info': info.cpu().numpy() if np.isnan(o[:, 0]).any() else np.zeros(B, dtype=np.int32)}
        if with_grad:
            res['grad'] = {'kern.' + self.kern.beta_param_name(): o[:, 1] / beta0, 'kern.variance': o[:, 2].copy(),
                           'likelihood.variance': o[:, 3].copy(), 'mean_function.c': o[:, 4].copy()}
        return res

    # ---- AutoFlow-style NumPy API ---------------------------------------------------------------------
    def compute_log_likelihood(self) -> float:
        return float(self.build_likelihood().item())

    def predict_f(self, Xnew):
        m, v = self.build_predict(Xnew, full_cov=False)
        return m.cpu().numpy(), v.cpu().numpy()

    def predict_f_full_cov(self, Xnew):
        m, c = self.build_predict(Xnew, full_cov=True)
        return m.cpu().numpy(), c.cpu().numpy()

    def predict_y(self, Xnew):
        m, v = self.build_predict(Xnew, full_cov=False)
        return m.cpu().numpy(), (v + float(self.likelihood.variance)).cpu().numpy()

    # ---- hyper-parameter fitting ---------------------------------------------------------------------------
    def _free_params(self):
        ps = [(f'kern.{k}', v) for k, v in sorted(self.kern.params().items()) if not getattr(v, 'fixed', False)]
        if not getattr(self.likelihood.variance, 'fixed', False):
            ps.append(('likelihood.variance', self.likelihood.variance))
        if isinstance(self.mean_function, Constant) and not getattr(self.mean_function.c, 'fixed', False):
            ps.append(('mean_function.c', self.mean_function.c))
        return ps

    def optimize(self, method='L-BFGS-B', maxiter=1000, analytic_grad=True, fused_small=True, **kwargs):
        """Maximise the log marginal likelihood over the free hyper-parameters, as ``GPR.optimize()`` of GPflow 0.5 does
        with SciPy L-BFGS-B in the unconstrained (softplus) space (reference call sites: ``sphere_kernels.py:133``,
        ``spd_kernels.py:155-187``).  Every objective evaluation runs on the GPU (Gram -> Cholesky -> solve -> log-lik);
        gradients are analytic (``loglik_gradients``, chained through the softplus transform); ``analytic_grad=False``
        falls back to SciPy's finite differences.  For N <= 160 (``small_path_available``) value and gradient come from ONE
        fused launch per evaluation (``log_likelihood_batch``; ``fused_small=False`` keeps the tiled path).  Returns the SciPy
        result."""
        from scipy.optimize import minimize
        ps = self._free_params()
        if not ps:
            return None

        def to_free(p):                      # inverse softplus for 'positive' Params (GPflow transforms.positive)
            v = float(p)
            if p.transform == 'positive':
                v = max(v, 1e-12)
                return v + np.log(-np.expm1(-v)) if v < 30 else v
            return v

        def from_free(p, x):
            if p.transform == 'positive':
                return float(np.logaddexp(0.0, x))
            return float(x)

        def set_all(x):
            for (_, p), xi in zip(ps, x):
                p.assign(from_free(p, xi))

        def objective(x):
            set_all(x)
            try:
                return -self.compute_log_likelihood()
            except np.linalg.LinAlgError:     # non-PD Gram (e.g. beta below the PD threshold): reject the step
                return 1e25

        def objective_and_grad(x):
            set_all(x)
            try:
                f = -self.compute_log_likelihood()
            except np.linalg.LinAlgError:
                return 1e25, np.zeros_like(x)
            g = self.loglik_gradients()
            out = np.empty_like(x)
            for i, (name, p) in enumerate(ps):
                dval = 1.0 if p.transform != 'positive' else -np.expm1(-float(p))    # d softplus(x)/dx = 1 - exp(-value)
                out[i] = -g[name] * dval
            return f, out

        def objective_and_grad_small(x):     # N <= 160: value and gradient in one fused launch
            set_all(x)
            r = self.log_likelihood_batch(with_grad=True)
            if int(r['info'][0]) != 0:
                return 1e25, np.zeros_like(x)
            out = np.empty_like(x)
            for i, (name, p) in enumerate(ps):
                dval = 1.0 if p.transform != 'positive' else -np.expm1(-float(p))
                out[i] = -r['grad'][name][0] * dval
            return -float(r['loglik'][0]), out

        x0 = np.array([to_free(p) for _, p in ps], dtype=np.float64)
        known = {'likelihood.variance', 'kern.variance', 'mean_function.c', 'kern.' + str(self.kern.beta_param_name())}
        if analytic_grad and fused_small and all(name in known for name, _ in ps) and self.small_path_available():
            res = minimize(objective_and_grad_small, x0, method=method, jac=True, options={'maxiter': maxiter}, **kwargs)
        elif analytic_grad and all(name in known for name, _ in ps):
            res = minimize(objective_and_grad, x0, method=method, jac=True, options={'maxiter': maxiter}, **kwargs)
        else:
            res = minimize(objective, x0, method=method, options={'maxiter': maxiter}, **kwargs)
        set_all(res.x)
        return res
