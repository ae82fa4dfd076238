"""Small-N fused GP evaluation (gabo_gp_small_batch_f64, GPR.log_likelihood_batch): the sizes of BASELINE.json configs 1-3
(N = 5..160).  Log-likelihood against the ORACLE (north_star: 1e-8 relative), gradients against the tiled analytic path and
against central differences of the oracle."""
import numpy as np
import pytest
import torch

from oracle import gabo_oracle as orc

pytestmark = pytest.mark.gpu


def _fd(f, h):
    return (f(+h) - f(-h)) / (2.0 * h)


@pytest.mark.parametrize('n,D,p', [(1, 3, 1), (2, 3, 1), (5, 3, 1), (30, 3, 1), (31, 3, 2), (64, 3, 1), (100, 4, 1), (129, 3, 3),
                                   (160, 51, 1)])
def test_batch_loglik_vs_oracle(cuda, n, D, p):
    from gaboflow_b200 import GPR, SphereGaussianKernel
    X = orc.synth_sphere(n, D, seed=100 + n)
    Y = np.random.default_rng(n).standard_normal((n, p))
    k = SphereGaussianKernel(D, range(D), beta_min=0.5, beta=1.5, variance=1.3)       # effective beta 2.0
    m = GPR(X, Y, kern=k)
    rng = np.random.default_rng(7)
    B = 37
    betas = rng.uniform(0.7, 6.0, B)
    vs = rng.uniform(0.3, 3.0, B)
    s2 = rng.uniform(1e-3, 0.5, B)
    cs = rng.uniform(-0.5, 0.5, B)
    r = m.log_likelihood_batch(beta=betas, variance=vs, noise_variance=s2, mean_const=cs, with_grad=False)
    assert np.all(r['info'] == 0)
    for b in range(B):
        want = orc.gp_loglik(orc.sphere_gaussian_K(X, beta=betas[b], variance=vs[b]), Y, s2[b], mean=cs[b])
        assert abs(r['loglik'][b] - want) <= 1e-10 * max(1.0, abs(want)), (b, r['loglik'][b], want)


def test_batch_gradients_vs_oracle_finite_differences(cuda):
    from gaboflow_b200 import GPR, SphereGaussianKernel
    from gaboflow_b200.gpr import Constant
    n = 60
    X = orc.synth_sphere(n, 3, seed=21)
    Y = np.random.default_rng(22).standard_normal((n, 2))
    bmin, b, v, s2, c = 1.5, 2.5, 1.7, 0.3, 0.2
    k = SphereGaussianKernel(3, range(3), beta_min=bmin, beta=1.0, variance=1.0)      # G is taken at another beta on purpose
    m = GPR(X, Y, kern=k, mean_function=Constant(0.0))
    r = m.log_likelihood_batch(beta=[b + bmin, 3.0], variance=[v, 1.0], noise_variance=[s2, 0.1], mean_const=[c, 0.0])
    ll = lambda bb, vv, ss, cc: orc.gp_loglik(orc.sphere_gaussian_K(X, beta=bb + bmin, variance=vv), Y, ss, mean=cc)
    ref = {'kern.beta_shifted': _fd(lambda h: ll(b + h, v, s2, c), 1e-5),
           'kern.variance': _fd(lambda h: ll(b, v + h, s2, c), 1e-5),
           'likelihood.variance': _fd(lambda h: ll(b, v, s2 + h, c), 1e-5),
           'mean_function.c': _fd(lambda h: ll(b, v, s2, c + h), 1e-5)}
    for name, want in ref.items():
        assert abs(r['grad'][name][0] - want) < 1e-6 * max(1.0, abs(want)), (name, r['grad'][name][0], want)


@pytest.mark.parametrize('which', ['sphere_gauss', 'sphere_laplace', 'spd_ai', 'spd_ai_laplace', 'spd_logeuclid', 'spd_frobenius'])
def test_batch_matches_tiled_path(cuda, which):
    """Value and gradients of the fused kernel against the tiled path (Gram -> gabo_potrf_f64 -> ... -> gabo_gp_grad_beta_f64)."""
    from gaboflow_b200 import (GPR, SphereGaussianKernel, SphereLaplaceKernel, SpdAffineInvariantGaussianKernel,
                               SpdAffineInvariantLaplaceKernel, SpdLogEuclideanGaussianKernel, SpdFrobeniusGaussianKernel)
    n = 90
    if which.startswith('sphere'):
        X = orc.synth_sphere(n, 4, seed=31)
        k = (SphereGaussianKernel(4, range(4), beta_min=0.3, beta=1.3, variance=0.9) if which == 'sphere_gauss'
             else SphereLaplaceKernel(4, range(4), beta=1.3, variance=0.9))
    else:
        X = orc.synth_spd_mandel(n, 3, seed=32)
        k = {'spd_ai': lambda: SpdAffineInvariantGaussianKernel(6, range(6), beta_min=0.2, beta=0.4, variance=1.2),
             'spd_ai_laplace': lambda: SpdAffineInvariantLaplaceKernel(6, range(6), beta=0.7, variance=1.2),
             'spd_logeuclid': lambda: SpdLogEuclideanGaussianKernel(6, range(6), beta=0.5, variance=0.8),
             'spd_frobenius': lambda: SpdFrobeniusGaussianKernel(6, range(6), beta=0.5, variance=0.8)}[which]()
    y = np.random.default_rng(33).standard_normal((n, 1))
    m = GPR(X, y, kern=k)
    m.likelihood.variance = 0.25
    r = m.log_likelihood_batch()
    want_ll = m.compute_log_likelihood()
    want_g = m.loglik_gradients()
    assert int(r['info'][0]) == 0
    assert abs(r['loglik'][0] - want_ll) <= 1e-10 * abs(want_ll)
    for name in ('kern.' + k.beta_param_name(), 'kern.variance', 'likelihood.variance', 'mean_function.c'):
        assert abs(r['grad'][name][0] - want_g[name]) <= 1e-8 * max(1.0, abs(want_g[name])), (which, name)


def test_batch_reports_non_pd_per_set(cuda):
    """A set whose Gram is not positive definite gets info > 0 and NaN outputs; its neighbours are unaffected."""
    from gaboflow_b200 import ops
    n = 40
    X = orc.synth_sphere(n, 3, seed=5)
    K = orc.sphere_gaussian_K(X, beta=2.0)
    G = torch.from_numpy(np.log(K)).cuda()
    G.fill_diagonal_(0.0)
    y = torch.from_numpy(np.random.default_rng(1).standard_normal((n, 1))).cuda()
    theta = torch.tensor([[1.0, 1.0, 0.1, 0.0], [1.0, 1.0, -1.5, 0.0], [2.0, 0.5, 0.2, 0.1]], dtype=torch.float64, device='cuda')
    out, info = ops.gp_small_batch(G, y, theta)
    info = info.cpu().numpy()
    out = out.cpu().numpy()
    assert info[0] == 0 and info[2] == 0 and info[1] > 0
    assert np.all(np.isnan(out[1])) and np.all(np.isfinite(out[0])) and np.all(np.isfinite(out[2]))
    want = orc.gp_loglik(orc.sphere_gaussian_K(X, beta=4.0, variance=0.5), y.cpu().numpy(), 0.2, mean=0.1)
    assert abs(out[2, 0] - want) <= 1e-10 * abs(want)


def test_batch_factor_and_whitened_outputs(cuda):
    from gaboflow_b200 import ops
    import scipy.linalg as sla
    n, p = 50, 2
    X = orc.synth_sphere(n, 3, seed=9)
    K = orc.sphere_gaussian_K(X, beta=1.0)
    G = torch.from_numpy(np.log(K)).cuda()
    G.fill_diagonal_(0.0)
    Y = np.random.default_rng(2).standard_normal((n, p))
    theta = torch.tensor([[3.0, 1.5, 0.05, 0.2], [1.0, 1.0, 0.3, 0.0]], dtype=torch.float64, device='cuda')
    Lout = torch.zeros((2, n, n), dtype=torch.float64, device='cuda')
    Vout = torch.zeros((2, n, p), dtype=torch.float64, device='cuda')
    ops.gp_small_batch(G, torch.from_numpy(Y).cuda(), theta, want_grad=False, Lout=Lout, Vout=Vout)
    for b, (r, v, s2, c) in enumerate(theta.cpu().numpy()):
        Lref = sla.cholesky(orc.sphere_gaussian_K(X, beta=r, variance=v) + s2 * np.eye(n), lower=True)
        assert np.max(np.abs(np.tril(Lout[b].cpu().numpy()) - Lref)) < 1e-12
        Vref = sla.solve_triangular(Lref, Y - c, lower=True)
        assert np.max(np.abs(Vout[b].cpu().numpy() - Vref)) < 1e-11


def test_optimize_uses_fused_path_and_matches_tiled_run(cuda, golden):
    from gaboflow_b200 import GPR, SphereGaussianKernel, _lib
    g = golden['sphere_s2']
    runs = []
    for fused in (True, False):
        k = SphereGaussianKernel(input_dim=3, active_dims=range(3), beta_min=7.0, beta=10.0, variance=1.)
        m = GPR(g['x_man'], g['y_train'], kern=k)
        n0 = _lib.lib().gabo_launch_count()
        res = m.optimize(fused_small=fused)
        launches = _lib.lib().gabo_launch_count() - n0
        runs.append((m.compute_log_likelihood(), res.nfev, launches))
    assert abs(runs[0][0] - runs[1][0]) < 1e-6 * abs(runs[1][0])
    assert runs[0][2] < runs[1][2] / 4          # one launch per evaluation instead of the tiled chain


def test_stein_and_large_n_are_refused(cuda):
    from gaboflow_b200 import GPR, SpdSteinGaussianKernel, SphereGaussianKernel
    X = orc.synth_spd_mandel(20, 3, seed=1)
    m = GPR(X, np.zeros((20, 1)), kern=SpdSteinGaussianKernel(6, range(6), beta=1.6))
    assert not m.small_path_available()
    with pytest.raises(NotImplementedError):
        m.log_likelihood_batch()
    Xs = orc.synth_sphere(161, 3, seed=1)
    assert not GPR(Xs, np.zeros((161, 1)), kern=SphereGaussianKernel(3, range(3), beta_min=0., beta=1.)).small_path_available()
