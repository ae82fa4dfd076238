"""batched SPD exponential / logarithm maps and the device-side beta sweep (extreme eigenvalues of
the Gram over a beta grid), against the reference-generated fixture tests/golden/spd_maps_sweep.npz and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import gabo_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'spd_maps_sweep.npz')


@pytest.fixture(scope='module')
def fx():
    return dict(np.load(GOLD))


@pytest.mark.parametrize('d', [2, 3, 6])
def test_expmap_logmap_vs_reference_fixture(cuda, fx, d):
    from gaboflow_b200 import riemannian_utils as ru
    S, U = fx[f'd{d}_S'], fx[f'd{d}_U']
    X = ru.expmap(U, S).cpu().numpy()
    scale = np.max(np.abs(fx[f'd{d}_exp']), axis=(1, 2), keepdims=True)
    assert np.max(np.abs(X - fx[f'd{d}_exp']) / scale) < 1e-12
    assert np.array_equal(X, np.swapaxes(X, 1, 2))                      # symmetric by construction
    Ub = ru.logmap(fx[f'd{d}_exp'], S).cpu().numpy()
    # log(exp(U)) = U: the reference's own round trip carries cond-amplified rounding, ours is compared with the input
    assert np.max(np.abs(Ub - U)) < 1e-11 * max(1.0, float(np.max(np.abs(U))))
    assert np.max(np.abs(Ub - fx[f'd{d}_log_of_exp'])) < 1e-10
    # Mandel in / Mandel out (SPD_utils.py:142-177)
    xm = ru.expmap_mandel_vector(fx[f'd{d}_u_mandel'], fx[f'd{d}_s_mandel']).cpu().numpy()
    assert np.max(np.abs(xm - fx[f'd{d}_exp_mandel']) / scale[:, :, 0]) < 1e-12
    um = ru.logmap_mandel_vector(fx[f'd{d}_exp_mandel'], fx[f'd{d}_s_mandel']).cpu().numpy()
    assert np.max(np.abs(um - fx[f'd{d}_u_mandel'])) < 1e-11 * max(1.0, float(np.max(np.abs(U))))
    # single matrix in, single matrix out
    x1 = ru.expmap(U[0], S).cpu().numpy()
    assert x1.shape == (d, d) and np.max(np.abs(x1 - fx[f'd{d}_exp'][0])) < 1e-12 * scale[0, 0, 0]


def test_maps_vs_oracle_all_dims_and_per_point_base(cuda):
    from gaboflow_b200 import ops
    rng = np.random.default_rng(5)
    for d in range(2, 9):
        n = 33
        A = rng.standard_normal((n, d, d))
        S = A @ np.swapaxes(A, 1, 2) + 0.5 * np.eye(d)
        U = rng.standard_normal((n, d, d))
        U = 0.25 * (U + np.swapaxes(U, 1, 2))
        Xg = ops.spd_expmap(torch.from_numpy(U).cuda(), torch.from_numpy(S).cuda(), d).cpu().numpy()
        Xo = np.stack([orc.spd_expmap(U[i], S[i]) for i in range(n)])
        assert np.max(np.abs(Xg - Xo) / np.max(np.abs(Xo), axis=(1, 2), keepdims=True)) < 1e-12, d
        Ug = ops.spd_logmap(torch.from_numpy(Xo).cuda(), torch.from_numpy(S).cuda(), d).cpu().numpy()
        assert np.max(np.abs(Ug - U)) < 1e-10, d


def test_map_info_flags_bad_inputs(cuda):
    from gaboflow_b200 import ops
    S = torch.eye(3, dtype=torch.float64, device='cuda')
    X = torch.stack([torch.eye(3, dtype=torch.float64), torch.diag(torch.tensor([1.0, -1.0, 2.0], dtype=torch.float64))]).cuda()
    out, info = ops.spd_logmap(X, S, 3, want_info=True)
    assert info.tolist()[0] == 0 and info.tolist()[1] > 3 and bool(torch.isnan(out[1]).all()) and bool((out[0] == 0).all())
    bad_base = torch.diag(torch.tensor([1.0, 0.0, 1.0], dtype=torch.float64)).cuda()
    out, info = ops.spd_expmap(X[:1], bad_base, 3, want_info=True)
    assert info.tolist() == [2] and bool(torch.isnan(out).all())


def test_mandel_roundtrip_and_distances(cuda, golden):
    from gaboflow_b200 import riemannian_utils as ru
    g = golden['spd_s6pp']
    S = ru.vector_to_symmetric_matrix(g['x'], 6)
    assert np.max(np.abs(S.cpu().numpy() - orc.mandel_to_sym(g['x'], 6))) < 1e-15
    v = ru.symmetric_matrix_to_vector(S, 6).cpu().numpy()
    assert np.max(np.abs(v - g['x'])) < 1e-15
    D = ru.affine_invariant_distance(g['S'], g['S'][:20], full_dist_mat=True).cpu().numpy()
    assert np.max(np.abs(D - g['Dai'][:, :20])) < 1e-13
    dp = ru.affine_invariant_distance(g['S'][:20], g['S'][20:40]).cpu().numpy()
    assert np.max(np.abs(dp - np.array([g['Dai'][i, 20 + i] for i in range(20)]))) < 1e-13
    s = golden['sphere_s2']
    Ds = ru.sphere_distance(s['x_man'], s['x_test']).cpu().numpy()
    assert np.max(np.abs(Ds - s['D12'])) < 1e-14


def test_beta_sweep_reference_trial(cuda, fx):
    """The reference experiment's own first trial (sphere_gaussian_kernel_parameters.py, dim = 4, 500 points, 30 betas):
    minimum eigenvalue per beta against the numpy.linalg.eig minima the reference script computes."""
    from gaboflow_b200 import SphereGaussianKernel
    from gaboflow_b200.sweep import kernel_eigen_sweep, percentage_pd, select_beta_min
    X, betas = fx['sweep_x'], fx['sweep_betas']
    r = kernel_eigen_sweep(lambda b: SphereGaussianKernel(4, range(4), beta_min=0., beta=b), X, betas)
    assert r['min_eig'].shape == (30,) and np.all(r['steps'] == 500) and np.all(r['resid'] == 0)
    lmax = r['max_eig']
    err = np.abs(r['min_eig'] - fx['sweep_min_eig'])
    assert np.max(err / lmax) < 1e-12, float(np.max(err / lmax))
    # same PD verdict wherever the reference's value is clear of its own rounding noise
    clear = np.abs(fx['sweep_min_eig']) > 1e-12 * lmax
    assert np.array_equal((r['min_eig'] > 0)[clear], (fx['sweep_min_eig'] > 0)[clear])
    assert np.array_equal(percentage_pd(r['min_eig'])[clear], (fx['sweep_min_eig'] > 0)[clear].astype(float))
    bm = select_beta_min(betas, r['min_eig'])
    assert bm is not None and bm == betas[np.nonzero(fx['sweep_min_eig'] <= 0)[0].max() + 1]


def test_beta_sweep_trials_and_kernels_vs_oracle(cuda):
    from gaboflow_b200 import SphereLaplaceKernel, SpdAffineInvariantGaussianKernel, SpdLogEuclideanGaussianKernel
    from gaboflow_b200.sweep import kernel_eigen_sweep
    # three sphere trials at once, Laplace kernel, variance != 1
    trials = [orc.sweep_trial_sphere(dim=3, nb_samples=200, nb_sources=5, seed=s) for s in (1, 2, 3)]
    betas = np.logspace(-0.5, 1.5, 7)
    r = kernel_eigen_sweep(lambda b: SphereLaplaceKernel(3, range(3), beta=b, variance=1.7), trials, betas)
    assert r['min_eig'].shape == (3, 7)
    for t, X in enumerate(trials):
        def K_exact_diag(b, X=X):
            # parity convention 1: K(X, X) has a diagonal of exactly `variance`; the reference's own Laplace
            # diagonal is variance * exp(-beta * acos(1 - k eps)) = variance * (1 - beta * 1e-8), which moves lambda_max by up
            # to 4e-8 relative on this grid (and lambda_min by 1e-11)
            K = orc.sphere_laplace_K(X, beta=b, variance=1.7)
            np.fill_diagonal(K, 1.7)
            return K
        lo, hi = orc.min_eig_sweep(K_exact_diag, betas)
        # entries of K_beta carry (beta / beta_0) * 2e-16 relative error (sweep.py), here beta / beta_0 = 100 and n = 200:
        # eigenvalue error <= n * 2e-14 = 4e-12 absolute
        assert np.max(np.abs(r['min_eig'][t] - lo)) < 4e-12 and np.max(np.abs(r['max_eig'][t] / hi - 1)) < 1e-12
        # against the reference's noisy diagonal: Weyl's bound, |shift of an eigenvalue| <= max |diagonal deviation|
        lo_ref, _ = orc.min_eig_sweep(lambda b: orc.sphere_laplace_K(X, beta=b, variance=1.7), betas)
        dev = np.array([np.max(np.abs(np.diag(orc.sphere_laplace_K(X, beta=float(b), variance=1.7)) - 1.7)) for b in betas])
        assert np.all(np.abs(r['min_eig'][t] - lo_ref) <= dev + 4e-12)
    # SPD affine-invariant (spd_gaussian_kernel_parameters.py trial, d = 2) and Log-Euclidean
    V = orc.sweep_trial_spd(dim=2, nb_samples=300, nb_sources=10, seed=4)
    betas = np.logspace(-1, 1.5, 6)
    r = kernel_eigen_sweep(lambda b: SpdAffineInvariantGaussianKernel(3, range(3), beta_min=0., beta=b), V, betas)
    lo, hi = orc.min_eig_sweep(lambda b: orc.spd_ai_gaussian_K(V, beta=b), betas)
    assert np.max(np.abs(r['min_eig'] - lo) / hi) < 1e-11, np.abs(r['min_eig'] - lo) / hi
    r = kernel_eigen_sweep(lambda b: SpdLogEuclideanGaussianKernel(3, range(3), beta=b), V, betas)
    lo, hi = orc.min_eig_sweep(lambda b: orc.spd_logeuclid_K(V, beta=b), betas)
    assert np.max(np.abs(r['min_eig'] - lo) / hi) < 1e-11


def test_beta_sweep_truncated_run_gives_upper_bound(cuda):
    from gaboflow_b200 import SphereGaussianKernel
    from gaboflow_b200.sweep import kernel_eigen_sweep
    X = orc.synth_sphere(1500, 8, seed=11)
    betas = np.array([0.5, 2.0, 8.0])
    r = kernel_eigen_sweep(lambda b: SphereGaussianKernel(8, range(8), beta_min=0., beta=b), X, betas, m_max=96)
    lo, hi = orc.min_eig_sweep(lambda b: orc.sphere_gaussian_K(X, beta=b), betas)
    assert np.all(r['steps'] == 96)
    assert np.all(r['min_eig'] >= lo - 1e-12 * hi)                    # Ritz values lie inside the spectrum
    assert np.max(np.abs(r['max_eig'] / hi - 1)) < 1e-10              # the top of the spectrum converges in a few steps
