"""Parity of the FP64 Cholesky / triangular solves / GP log-likelihood / posterior against the oracle.
Tolerances (BASELINE.json north_star): log marginal likelihood and posterior mean/variance 1e-8 relative;
factor entries are checked far tighter on well-conditioned inputs."""
import numpy as np
import pytest
import scipy.linalg as sla
import torch

from oracle import gabo_oracle as orc

pytestmark = pytest.mark.gpu


def spd_matrix(n, seed, cond_shift=None):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, max(n // 2, 4)))
    A = G @ G.T / G.shape[1]
    A[np.diag_indices(n)] += (cond_shift if cond_shift is not None else 0.5)
    return A


@pytest.mark.parametrize('n', [1, 2, 17, 128, 129, 300, 512, 513, 640, 1100, 2049])
def test_potrf_vs_scipy(cuda, n):
    from gaboflow_b200 import ops
    A = spd_matrix(n, n)
    ref = sla.cholesky(A + 0.01 * np.eye(n), lower=True)
    d = torch.from_numpy(A).cuda()
    d_up = torch.triu(d, 1).clone()
    info = ops.potrf_(d, sigma2=0.01)
    assert int(info.item()) == 0
    got = d.cpu().numpy()
    assert np.max(np.abs(np.tril(got) - ref)) / np.max(np.abs(ref)) < 1e-13
    assert torch.equal(torch.triu(d, 1), d_up)        # strictly upper triangle untouched


def test_potrf_reports_non_pd(cuda):
    from gaboflow_b200 import ops
    n = 400
    A = spd_matrix(n, 5)
    A[250, 250] = -1.0
    d = torch.from_numpy(A).cuda()
    info = ops.potrf_(d, check=False)
    assert int(info.item()) == 251
    with pytest.raises(np.linalg.LinAlgError):
        ops.potrf_(torch.from_numpy(A).cuda(), check=True)


@pytest.mark.parametrize('n,nrhs', [(5, 1), (130, 1), (700, 3), (700, 4), (300, 5), (1000, 64), (1000, 200), (257, 130)])
@pytest.mark.parametrize('trans', [False, True])
def test_trsm_vs_scipy(cuda, n, nrhs, trans):
    from gaboflow_b200 import ops
    L = np.linalg.cholesky(spd_matrix(n, 3 * n))
    B = np.random.default_rng(n + nrhs).standard_normal((n, nrhs))
    ref = sla.solve_triangular(L, B, lower=True, trans='T' if trans else 'N')
    Ld = torch.from_numpy(L).cuda()
    Bd = torch.from_numpy(B).cuda()
    ops.trsm_(Ld, Bd, trans=trans)
    assert np.max(np.abs(Bd.cpu().numpy() - ref)) / np.max(np.abs(ref)) < 1e-12


@pytest.mark.parametrize('n,nrhs', [(1024, 32), (1025, 33), (1536, 64), (2049, 130), (3000, 257), (4096, 1024)])
def test_trsm_many_rhs_blocked_path(cuda, n, nrhs):
    """Forward solve with many right-hand sides: the two-level right-looking solve on the transposed right-hand sides
    (trsm_forward_blocked: n >= 1024, nrhs >= 32), odd sizes and leading dimensions included, against SciPy and against the
    single-level path on a strided view."""
    from gaboflow_b200 import ops
    L = np.linalg.cholesky(spd_matrix(n, 3 * n))
    B = np.random.default_rng(n + nrhs).standard_normal((n, nrhs))
    ref = sla.solve_triangular(L, B, lower=True)
    Ld = torch.from_numpy(L).cuda()
    Bd = torch.from_numpy(B).cuda()
    ops.trsm_(Ld, Bd)
    assert np.max(np.abs(Bd.cpu().numpy() - ref)) / np.max(np.abs(ref)) < 1e-12
    # L and B as views into larger buffers (leading dimensions n + 3 and nrhs + 5)
    Lbig = torch.zeros((n, n + 3), dtype=torch.float64, device='cuda')
    Lbig[:, :n] = Ld
    Bbig = torch.zeros((n, nrhs + 5), dtype=torch.float64, device='cuda')
    Bbig[:, :nrhs] = torch.from_numpy(B).cuda()
    ops.trsm_(Lbig[:, :n], Bbig[:, :nrhs])
    assert np.max(np.abs(Bbig[:, :nrhs].cpu().numpy() - ref)) / np.max(np.abs(ref)) < 1e-12
    assert float(Bbig[:, nrhs:].abs().max()) == 0.0


@pytest.mark.parametrize('m,n,k,lower', [(128, 128, 16, False), (300, 200, 77, False), (513, 513, 128, True), (260, 130, 512, True)])
def test_gemm_nt_sub(cuda, m, n, k, lower):
    from gaboflow_b200 import ops
    rng = np.random.default_rng(m + n + k)
    A, B, C = rng.standard_normal((m, k)), rng.standard_normal((n, k)), rng.standard_normal((m, n))
    ref = C - A @ B.T
    Cd = torch.from_numpy(C).cuda()
    ops.gemm_nt_sub_(Cd, torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), lower=lower)
    got = Cd.cpu().numpy()
    if lower:
        mask = np.tril(np.ones((m, n), dtype=bool))
        assert np.max(np.abs(got[mask] - ref[mask])) < 1e-12 * k
        assert np.array_equal(got[~mask], C[~mask])
    else:
        assert np.max(np.abs(got - ref)) < 1e-12 * k


def test_gpr_sphere_example_known_answers(cuda, golden):
    """examples/kernels/sphere/sphere_kernels.py:125-139 with the kernel parameters as constructed
    (no optimize()); known answers from SURVEY.md 8(c)."""
    from gaboflow_b200 import GPR, SphereGaussianKernel
    g = golden['sphere_s2']
    k = SphereGaussianKernel(input_dim=3, active_dims=range(3), beta_min=7.0, beta=10.0, variance=1.)
    m = GPR(g['x_man'], g['y_train'], kern=k, mean_function=None)
    ll = m.compute_log_likelihood()
    assert abs(ll / -23.67939289441117 - 1) < 1e-8
    mean, var = m.predict_f(g['x_test'])
    Kor = orc.sphere_gaussian_K(g['x_man'], beta=17.0)
    mo, vo = orc.gp_predict(Kor, orc.sphere_gaussian_K(g['x_man'], g['x_test'], beta=17.0), np.ones(10), g['y_train'], 1.0)
    assert np.max(np.abs(mean - mo) / np.maximum(np.abs(mo), 1e-12)) < 1e-8
    assert np.max(np.abs(var - vo) / np.abs(vo)) < 1e-8
    assert np.allclose(mean[:3, 0], [1.2854030824754458e-03, 2.7761803306368668e-08, 9.6917159395751379e-02], rtol=1e-8, atol=0)
    assert np.allclose(var[:3, 0], [0.9998564833587267, 0.9999999999999013, 0.5014023044062732], rtol=1e-8, atol=0)
    mean2, cov = m.predict_f_full_cov(g['x_test'])
    _, co = orc.gp_predict(Kor, orc.sphere_gaussian_K(g['x_man'], g['x_test'], beta=17.0),
                           orc.sphere_gaussian_K(g['x_test'], beta=17.0), g['y_train'], 1.0, full_cov=True)
    assert cov.shape == (10, 10, 1)
    assert np.max(np.abs(cov[:, :, 0] - co)) < 1e-10
    # hyper-parameter change invalidates the cached factor
    k.update_beta(12.0)
    assert abs(m.compute_log_likelihood() - orc.gp_loglik(orc.sphere_gaussian_K(g['x_man'], beta=12.0), g['y_train'], 1.0)) < 1e-8


def test_gpr_spd_example_known_answers(cuda, golden):
    """examples/kernels/spd/spd_kernels.py:176-194 (affine-invariant kernel on the C.mat fixture)."""
    from gaboflow_b200 import GPR, SpdAffineInvariantGaussianKernel
    g = golden['spd_s2pp']
    k = SpdAffineInvariantGaussianKernel(input_dim=3, active_dims=range(3), beta=1.0, variance=1., beta_min=0.6)
    m = GPR(g['x_man'], g['y'], kern=k, mean_function=None)
    assert abs(m.compute_log_likelihood() / -27621.519513872187 - 1) < 1e-8
    mean, var = m.predict_f(g['x_test'])
    Kor = orc.spd_ai_gaussian_K(g['x_man'], beta=1.6)
    mo, vo = orc.gp_predict(Kor, orc.spd_ai_gaussian_K(g['x_man'], g['x_test'], beta=1.6), np.ones(100), g['y'], 1.0)
    assert np.max(np.abs(mean - mo) / np.maximum(np.abs(mo), 1e-9)) < 1e-8
    assert np.max(np.abs(var - vo) / np.abs(vo)) < 1e-8


@pytest.mark.parametrize('n', [1500, 4096])
def test_gp_pipeline_synthetic_sphere(cuda, n):
    """BASELINE config 4 at oracle-checkable size: sphere Gaussian Gram on S^50, sigma^2 = 1e-2, loglik + posterior."""
    from gaboflow_b200 import GPR, SphereGaussianKernel
    X = orc.synth_sphere(n, 51, seed=20240)
    y = np.random.default_rng(20241).standard_normal((n, 1))
    Xs = orc.synth_sphere(300, 51, seed=99)
    k = SphereGaussianKernel(51, range(51), beta_min=0.0, beta=2.0)
    m = GPR(X, y, kern=k)
    m.likelihood.variance = 1e-2
    Kor = orc.sphere_gaussian_K(X, beta=2.0)
    ll_ref = orc.gp_loglik(Kor, y, 1e-2)
    assert abs(m.compute_log_likelihood() / ll_ref - 1) < 1e-8
    mean, var = m.predict_f(Xs)
    mo, vo = orc.gp_predict(Kor, orc.sphere_gaussian_K(X, Xs, beta=2.0), np.ones(300), y, 1e-2)
    assert np.max(np.abs(mean - mo)) / np.max(np.abs(mo)) < 1e-8
    assert np.max(np.abs(var - vo) / np.abs(vo)) < 1e-8


def test_expected_improvement_vs_oracle(cuda, golden):
    """a12 of SURVEY section 8(a): EI from the posterior, against the oracle's restatement of GPflowOpt's formula."""
    from gaboflow_b200 import GPR, SphereGaussianKernel, ExpectedImprovement
    g = golden['sphere_s2']
    k = SphereGaussianKernel(input_dim=3, active_dims=range(3), beta_min=7.0, beta=10.0, variance=1.)
    m = GPR(g['x_man'], g['y_train'], kern=k)
    acq = ExpectedImprovement(m)
    Kor = orc.sphere_gaussian_K(g['x_man'], beta=17.0)
    m_tr, _ = orc.gp_predict(Kor, Kor, np.ones(20), g['y_train'], 1.0)
    assert abs(acq.fmin - m_tr.min()) < 1e-10
    cand = orc.synth_sphere(100, 3, seed=12)          # the reference scores 100 anchor candidates (manifold_optimization.py:402-431)
    mo, vo = orc.gp_predict(Kor, orc.sphere_gaussian_K(g['x_man'], cand, beta=17.0), np.ones(100), g['y_train'], 1.0)
    ref = orc.expected_improvement(mo, vo, m_tr.min())
    got = acq.evaluate(cand)
    assert got.shape == (100, 1)
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-12)) < 1e-8


def test_gpr_optimize_improves_and_matches_cpu_optimum(cuda, golden):
    """m.optimize() (sphere_kernels.py:133): same optimum as L-BFGS-B over the oracle's log-likelihood."""
    from scipy.optimize import minimize
    from gaboflow_b200 import GPR, SphereGaussianKernel
    g = golden['sphere_s2']
    k = SphereGaussianKernel(input_dim=3, active_dims=range(3), beta_min=7.0, beta=10.0, variance=1.)
    m = GPR(g['x_man'], g['y_train'], kern=k)
    ll0 = m.compute_log_likelihood()
    res = m.optimize()
    ll1 = m.compute_log_likelihood()
    assert ll1 > ll0 + 1.0 and abs(ll1 + res.fun) < 1e-9
    assert float(k.beta_shifted) > 0 and float(k.variance) > 0 and float(m.likelihood.variance) > 0

    sp = lambda x: np.logaddexp(0.0, x)
    def nll(x):
        return -orc.gp_loglik(orc.sphere_gaussian_K(g['x_man'], beta=sp(x[0]) + 7.0, variance=sp(x[1])), g['y_train'], sp(x[2]))
    inv = lambda v: v + np.log(-np.expm1(-v))
    ref = minimize(nll, [inv(10.0), inv(1.0), inv(1.0)], method='L-BFGS-B')
    assert abs(ll1 - (-ref.fun)) < 1e-3 * abs(ref.fun)


def test_trsm_with_supplied_block_inverses(cuda):
    """gabo_potrf_diag_f64 exports the inverses of the 128-blocks; gabo_trsm_winv_f64 solves with them (no inversion)."""
    from gaboflow_b200 import ops
    n = 512
    A = spd_matrix(n, 77)
    d = torch.from_numpy(A).cuda()
    winv = torch.empty((4, 128, 128), dtype=torch.float64, device='cuda')
    info = ops.potrf_diag_(d, 0.01, winv)
    assert int(info.item()) == 0
    ref = sla.cholesky(A + 0.01 * np.eye(n), lower=True)
    assert np.max(np.abs(np.tril(d.cpu().numpy()) - ref)) < 1e-12
    for b in range(4):
        blk = ref[b * 128:(b + 1) * 128, b * 128:(b + 1) * 128]
        assert np.max(np.abs(winv[b].cpu().numpy() @ blk - np.eye(128))) < 1e-12
    for nrhs in (1, 3, 70):
        B = np.random.default_rng(nrhs).standard_normal((n, nrhs))
        for trans in (False, True):
            Bd = torch.from_numpy(B).cuda()
            ops.trsm_(d, Bd, trans=trans, winv_all=winv)
            want = sla.solve_triangular(ref, B, lower=True, trans='T' if trans else 'N')
            assert np.max(np.abs(Bd.cpu().numpy() - want)) / np.max(np.abs(want)) < 1e-12


def _fd(f, h):
    return (f(+h) - f(-h)) / (2.0 * h)


def test_loglik_gradients_vs_oracle_finite_differences(cuda):
    """GPR.loglik_gradients (analytic, gabo_gp_grad_beta_f64 + trace identities) against central finite differences of
    the ORACLE's log-likelihood: sphere Gaussian with beta_min shift, two output columns, constant mean."""
    from gaboflow_b200 import GPR, SphereGaussianKernel
    from gaboflow_b200.gpr import Constant
    n = 300
    X = orc.synth_sphere(n, 3, seed=21)
    Y = np.random.default_rng(22).standard_normal((n, 2))
    bmin, b, v, s2, c = 1.5, 2.5, 1.7, 0.3, 0.2
    k = SphereGaussianKernel(3, range(3), beta_min=bmin, beta=b, variance=v)
    m = GPR(X, Y, kern=k, mean_function=Constant(c))
    m.likelihood.variance = s2
    g = m.loglik_gradients()
    ll = lambda bb, vv, ss, cc: orc.gp_loglik(orc.sphere_gaussian_K(X, beta=bb + bmin, variance=vv), Y, ss, mean=cc)
    ref = {'kern.beta_shifted': _fd(lambda h: ll(b + h, v, s2, c), 1e-5),
           'kern.variance': _fd(lambda h: ll(b, v + h, s2, c), 1e-5),
           'likelihood.variance': _fd(lambda h: ll(b, v, s2 + h, c), 1e-5),
           'mean_function.c': _fd(lambda h: ll(b, v, s2, c + h), 1e-5)}
    for name, r in ref.items():
        assert abs(g[name] - r) < 1e-6 * max(1.0, abs(r)), (name, g[name], r)


@pytest.mark.parametrize('which', ['sphere_laplace', 'spd_ai', 'spd_stein', 'spd_logeuclid'])
def test_loglik_gradients_other_kernels(cuda, which):
    """Same check for the other kernel families, with finite differences of the GPU log-likelihood itself."""
    from gaboflow_b200 import (GPR, SphereLaplaceKernel, SpdAffineInvariantGaussianKernel, SpdSteinGaussianKernel,
                               SpdLogEuclideanGaussianKernel)
    n = 160
    if which == 'sphere_laplace':
        X = orc.synth_sphere(n, 4, seed=31)
        k = SphereLaplaceKernel(4, range(4), beta=1.3, variance=0.9)
    else:
        X = orc.synth_spd_mandel(n, 3, seed=32)
        if which == 'spd_ai':
            k = SpdAffineInvariantGaussianKernel(6, range(6), beta_min=0.2, beta=0.4, variance=1.2)
        elif which == 'spd_stein':
            k = SpdSteinGaussianKernel(6, range(6), beta=1.6, variance=1.1)
        else:
            k = SpdLogEuclideanGaussianKernel(6, range(6), beta=0.5, variance=0.8)
    y = np.random.default_rng(33).standard_normal((n, 1))
    m = GPR(X, y, kern=k)
    m.likelihood.variance = 0.25
    g = m.loglik_gradients()
    bname = k.beta_param_name()
    for name, par in (('kern.' + bname, getattr(k, bname)), ('kern.variance', k.variance),
                      ('likelihood.variance', m.likelihood.variance)):
        v0 = float(par)

        def f(h, par=par, v0=v0):
            par.assign(v0 + h)
            out = m.compute_log_likelihood()
            par.assign(v0)
            return out
        r = _fd(f, 1e-5 * max(1.0, abs(v0)))
        assert abs(g[name] - r) < 2e-6 * max(1.0, abs(r)), (which, name, g[name], r)


def test_optimize_with_analytic_gradients_matches_finite_difference_run(cuda, golden):
    from gaboflow_b200 import GPR, SphereGaussianKernel
    g = golden['sphere_s2']
    runs = []
    for analytic in (True, False):
        k = SphereGaussianKernel(input_dim=3, active_dims=range(3), beta_min=7.0, beta=10.0, variance=1.)
        m = GPR(g['x_man'], g['y_train'], kern=k)
        res = m.optimize(analytic_grad=analytic)
        runs.append((m.compute_log_likelihood(), res.nfev))
    assert abs(runs[0][0] - runs[1][0]) < 1e-4 * abs(runs[1][0])
    assert runs[0][1] < runs[1][1]            # fewer likelihood evaluations than finite differences


def test_gpr_append_matches_refactorisation(cuda):
    """GPR.append (rank-1 growth of the cached factor) against a model rebuilt from scratch on the same data."""
    from gaboflow_b200 import GPR, SphereGaussianKernel
    n0, extra = 250, 9                      # crosses the 256-row capacity of the first factor buffer
    X = orc.synth_sphere(n0 + extra, 3, seed=41)
    y = np.random.default_rng(42).standard_normal((n0 + extra, 1))
    Xs = orc.synth_sphere(17, 3, seed=43)
    mk = lambda: SphereGaussianKernel(3, range(3), beta_min=1.0, beta=2.0, variance=1.4)
    m = GPR(X[:n0], y[:n0], kern=mk())
    m.likelihood.variance = 0.05
    m.compute_log_likelihood()
    for i in range(n0, n0 + extra):
        m.append(X[i], y[i])
    full = GPR(X, y, kern=mk())
    full.likelihood.variance = 0.05
    assert abs(m.compute_log_likelihood() / full.compute_log_likelihood() - 1) < 1e-11
    La, Lf = m._factorise()[0], full._factorise()[0]
    assert float((torch.tril(La) - torch.tril(Lf)).abs().max()) < 1e-11
    ma, va = m.predict_f(Xs)
    mf, vf = full.predict_f(Xs)
    assert np.max(np.abs(ma - mf)) < 1e-10 and np.max(np.abs(va - vf)) < 1e-10
    with pytest.raises(np.linalg.LinAlgError):          # a duplicate point with zero noise is singular
        m2 = GPR(X[:50], y[:50], kern=mk())
        m2.likelihood.variance = 1e-300
        m2.compute_log_likelihood()
        m2.append(X[3], y[3])
    # a failed append leaves the model as it was
    ll_before = m.compute_log_likelihood()
    m.likelihood.variance = 0.05
    assert m.X.shape[0] == n0 + extra and abs(m.compute_log_likelihood() - ll_before) == 0.0
    # two output columns and the Stein kernel (whose diagonal is det(X)^beta, not `variance`)
    from gaboflow_b200 import SpdSteinGaussianKernel
    V = orc.synth_spd_mandel(60, 2, seed=44)
    Y2 = np.random.default_rng(45).standard_normal((60, 2))
    ks = lambda: SpdSteinGaussianKernel(3, range(3), beta=1.5)
    a = GPR(V[:55], Y2[:55], kern=ks()); a.likelihood.variance = 0.1
    a.compute_log_likelihood()
    for i in range(55, 60):
        a.append(V[i], Y2[i])
    b = GPR(V, Y2, kern=ks()); b.likelihood.variance = 0.1
    assert abs(a.compute_log_likelihood() / b.compute_log_likelihood() - 1) < 1e-11
    assert float((a._factorise()[1] - b._factorise()[1]).abs().max()) < 1e-10


@pytest.mark.parametrize('laplace', [False, True])
def test_candidate_gradients_vs_oracle_finite_differences(cuda, laplace):
    """gabo_sphere_posterior_grad_f64 (d mean/dx*, d var/dx*, d EI/dx*) against central finite differences of the ORACLE's
    posterior and EI in the ambient coordinates (the Euclidean gradient the manifold optimiser projects)."""
    from gaboflow_b200 import GPR, SphereGaussianKernel, SphereLaplaceKernel, ExpectedImprovement
    n, D, M = 220, 4, 6
    X = orc.synth_sphere(n, D, seed=51)
    y = np.sin(3.0 * X[:, :1]) + 0.1 * np.random.default_rng(52).standard_normal((n, 1))
    Xs = orc.synth_sphere(M, D, seed=53)
    beta, v, s2 = 1.8, 1.3, 0.05
    if laplace:
        k = SphereLaplaceKernel(D, range(D), beta=beta, variance=v)
        Kfun = lambda A, B=None: orc.sphere_laplace_K(A, B, beta=beta, variance=v)
    else:
        k = SphereGaussianKernel(D, range(D), beta_min=0.5, beta=beta - 0.5, variance=v)
        Kfun = lambda A, B=None: orc.sphere_gaussian_K(A, B, beta=beta, variance=v)
    m = GPR(X, y, kern=k)
    m.likelihood.variance = s2
    acq = ExpectedImprovement(m)
    fmin = acq.fmin
    mean, var, dmean, dvar, dei = m.build_predict_with_gradients(Xs, fmin=fmin)
    ei, dei2 = acq.evaluate_with_gradients(Xs)          # the fused one-launch path (sphere kernels)
    assert np.max(np.abs(dei.cpu().numpy() - dei2)) < 1e-10 * max(1.0, float(np.max(np.abs(dei2))))
    Kor = Kfun(X)
    np.fill_diagonal(Kor, v)          # K(X) has a diagonal of exactly `variance` here; the reference's carries acos noise

    def post(x):
        mu, vv = orc.gp_predict(Kor, Kfun(X, x[None, :]), np.full(1, v), y, s2)
        mu, vv = float(np.ravel(mu)[0]), float(np.ravel(vv)[0])
        return mu, vv, float(np.ravel(orc.expected_improvement(np.array([mu]), np.array([vv]), fmin))[0])

    h = 1e-6
    for c in range(M):
        mu0, v0, e0 = post(Xs[c])
        assert abs(float(mean[c, 0]) - mu0) < 1e-8 and abs(float(var[c]) - v0) < 1e-8 and abs(float(ei[c, 0]) - e0) < 1e-8
        for j in range(D):
            e = np.zeros(D); e[j] = h
            (mp, vp, ep), (mm, vm, em) = post(Xs[c] + e), post(Xs[c] - e)
            for got, ref in ((float(dmean[c, j]), (mp - mm) / (2 * h)), (float(dvar[c, j]), (vp - vm) / (2 * h)),
                             (float(dei[c, j]), (ep - em) / (2 * h))):
                assert abs(got - ref) < 2e-5 * max(1.0, abs(ref)), (laplace, c, j, got, ref)


@pytest.mark.parametrize('d,laplace', [(2, False), (2, True), (3, False), (6, False)])
def test_spd_ai_candidate_gradients_vs_oracle_finite_differences(cuda, d, laplace):
    """gabo_spd_ai_posterior_grad_f64 (d mean/dx*, d var/dx*, d EI/dx* in Mandel coordinates) against central finite
    differences of the ORACLE's posterior and EI -- what bo_spd_ackley_manifold.py:159-187 obtains from TensorFlow autodiff and
    hands to the manifold optimiser (manifold_optimization.py:260-271)."""
    from gaboflow_b200 import GPR, SpdAffineInvariantGaussianKernel, SpdAffineInvariantLaplaceKernel, ExpectedImprovement
    n, M = 90, 4
    mm = d * (d + 1) // 2
    X = orc.synth_spd_mandel(n, d, seed=61)
    y = np.log(orc.mandel_to_sym(X)[:, 0, 0])[:, None] + 0.05 * np.random.default_rng(62).standard_normal((n, 1))
    Xs = orc.synth_spd_mandel(M, d, seed=63)
    beta, v, s2 = 0.9, 1.2, 0.05
    if laplace:
        k = SpdAffineInvariantLaplaceKernel(mm, range(mm), beta=beta, variance=v)
        Kfun = lambda A, B=None: orc.spd_ai_laplace_K(A, B, beta=beta, variance=v)
    else:
        k = SpdAffineInvariantGaussianKernel(mm, range(mm), beta_min=0.4, beta=beta - 0.4, variance=v)
        Kfun = lambda A, B=None: orc.spd_ai_gaussian_K(A, B, beta=beta, variance=v)
    m = GPR(X, y, kern=k)
    m.likelihood.variance = s2
    acq = ExpectedImprovement(m)
    fmin = acq.fmin
    mean, var, dmean, dvar, dei = m.build_predict_with_gradients(Xs, fmin=fmin)
    ei, dei2 = acq.evaluate_with_gradients(Xs)
    assert np.array_equal(dei.cpu().numpy(), dei2)
    Kor = Kfun(X)
    np.fill_diagonal(Kor, v)

    def post(x):
        mu, vv = orc.gp_predict(Kor, Kfun(X, x[None, :]), np.full(1, v), y, s2)
        mu, vv = float(np.ravel(mu)[0]), float(np.ravel(vv)[0])
        return mu, vv, float(np.ravel(orc.expected_improvement(np.array([mu]), np.array([vv]), fmin))[0])

    h = 1e-6
    for c in range(M):
        mu0, v0, e0 = post(Xs[c])
        assert abs(float(mean[c, 0]) - mu0) < 1e-8 and abs(float(var[c]) - v0) < 1e-8 and abs(float(ei[c, 0]) - e0) < 1e-8
        for j in range(mm):
            e = np.zeros(mm); e[j] = h
            (mp, vp, ep), (mn, vm, em) = post(Xs[c] + e), post(Xs[c] - e)
            for got, ref in ((float(dmean[c, j]), (mp - mn) / (2 * h)), (float(dvar[c, j]), (vp - vm) / (2 * h)),
                             (float(dei[c, j]), (ep - em) / (2 * h))):
                assert abs(got - ref) < 3e-5 * max(1.0, abs(ref)), (d, laplace, c, j, got, ref)


@pytest.mark.parametrize('laplace', [False, True])
@pytest.mark.parametrize('n,D,M', [(5, 3, 1), (30, 3, 100), (500, 3, 7), (3000, 51, 33)])
def test_fused_acquisition_matches_unfused_path(cuda, laplace, n, D, M):
    """gabo_sphere_acq_fused_f64 (K(X,x*) -> L^-1 -> mean/var -> EI -> gradients in ONE launch, cached L^-1) against the
    separate launches (Gram, TRSM, predict, EI, TRSM^T, gradient kernel) and, for the values, against the oracle."""
    from gaboflow_b200 import GPR, SphereGaussianKernel, SphereLaplaceKernel, ExpectedImprovement
    X = orc.synth_sphere(n, D, seed=71)
    y = np.cos(2.0 * X[:, :1]) + 0.1 * np.random.default_rng(72).standard_normal((n, 1))
    Xs = orc.synth_sphere(M, D, seed=73)
    beta, v, s2 = 2.2, 0.9, 0.02
    if laplace:
        k = SphereLaplaceKernel(D, range(D), beta=beta, variance=v)
        Kfun = lambda A, B=None: orc.sphere_laplace_K(A, B, beta=beta, variance=v)
    else:
        k = SphereGaussianKernel(D, range(D), beta_min=0.7, beta=beta - 0.7, variance=v)
        Kfun = lambda A, B=None: orc.sphere_gaussian_K(A, B, beta=beta, variance=v)
    m = GPR(X, y, kern=k)
    m.likelihood.variance = s2
    fused, plain = ExpectedImprovement(m, fused=True), ExpectedImprovement(m, fused=False)
    assert m.fused_acquisition_available() and fused.fmin == plain.fmin
    e1, g1 = fused.evaluate_with_gradients(Xs)
    e2, g2 = plain.evaluate_with_gradients(Xs)
    assert e1.shape == e2.shape == (M, 1) and g1.shape == g2.shape == (M, D)
    scale = max(1.0, float(np.max(np.abs(g2))))
    assert np.max(np.abs(e1 - e2)) < 1e-11 and np.max(np.abs(g1 - g2)) < 1e-9 * scale
    assert np.max(np.abs(fused.evaluate(Xs) - e1)) == 0.0
    mean, var, ei = [t.cpu().numpy() for t in m.acquisition_fused(Xs, fused.fmin)[:3]]
    Kor = Kfun(X)
    np.fill_diagonal(Kor, v)
    mu_o, var_o = orc.gp_predict(Kor, Kfun(X, Xs), np.full(M, v), y, s2)
    assert np.max(np.abs(mean - mu_o[:, 0])) < 1e-8 * max(1.0, np.max(np.abs(mu_o)))
    assert np.max(np.abs(var - var_o[:, 0])) < 1e-8
    assert np.max(np.abs(ei - orc.expected_improvement(mu_o[:, 0], var_o[:, 0], fused.fmin))) < 1e-8
    # a new observation invalidates the cached inverse factor
    m.append(orc.synth_sphere(1, D, seed=74)[0], np.array([0.3]))
    f2 = ExpectedImprovement(m, fused=True)
    p2 = ExpectedImprovement(m, fused=False)
    assert np.max(np.abs(f2.evaluate(Xs) - p2.evaluate(Xs))) < 1e-11
