"""GPU: edge cases of the boundary -- strided views, leading dimensions, FP32 variants, Kdiag of every kernel class,
constant mean, factor on a sub-view (what bench.py does with the leading block of a larger Gram)."""
import numpy as np
import pytest
import scipy.linalg as sla
import torch

from oracle import gabo_oracle as orc

pytestmark = pytest.mark.gpu


def test_strided_inputs_and_outputs(cuda):
    from gaboflow_b200 import ops
    X = orc.synth_sphere(300, 51, seed=1)
    X2 = orc.synth_sphere(200, 51, seed=2)
    big = torch.zeros((300, 64), dtype=torch.float64, device='cuda')
    big[:, :51] = torch.from_numpy(X).cuda()
    big2 = torch.zeros((200, 70), dtype=torch.float64, device='cuda')
    big2[:, 3:54] = torch.from_numpy(X2).cuda()
    out_big = torch.full((300, 333), -7.0, dtype=torch.float64, device='cuda')
    ops.points_gram(ops.SPHERE_GAUSSIAN, big[:, :51], big2[:, 3:54], beta=2.0, out=out_big[:, 5:205])   # odd offset: scalar stores
    ref = orc.sphere_gaussian_K(X, X2, beta=2.0)
    got = out_big.cpu().numpy()
    assert np.max(np.abs(got[:, 5:205] / ref - 1)) < 1e-12
    assert np.all(got[:, :5] == -7.0) and np.all(got[:, 205:] == -7.0)


def test_fp32_variants(cuda):
    from gaboflow_b200 import ops
    X = orc.synth_sphere(257, 51, seed=3).astype(np.float32)
    Xd = torch.from_numpy(X).cuda()
    Kl = ops.points_gram(ops.SPHERE_LAPLACE, Xd, None, beta=2.0, variance=0.5, uplo='lower').cpu().numpy()
    ref = orc.sphere_laplace_K(X.astype(np.float64), beta=2.0, variance=0.5)
    low = np.tril(np.ones_like(ref, dtype=bool), -1)
    assert np.max(np.abs(Kl[low] / ref[low] - 1)) < 1e-5
    F = np.random.default_rng(4).standard_normal((190, 21)).astype(np.float32)
    Ks = ops.points_gram(ops.SQDIST_GAUSSIAN, torch.from_numpy(F).cuda(), None, beta=0.05).cpu().numpy()
    d2 = ((F[:, None, :].astype(np.float64) - F[None, :, :]) ** 2).sum(-1)
    assert np.max(np.abs(Ks - np.exp(-0.05 * d2))) < 2e-5
    assert np.all(np.diag(Ks) == 1.0)


def test_kdiag_every_kernel(cuda):
    from gaboflow_b200 import (SphereGaussianKernel, SphereLaplaceKernel, SpdSteinGaussianKernel,
                               SpdAffineInvariantGaussianKernel, SpdAffineInvariantLaplaceKernel,
                               SpdFrobeniusGaussianKernel, SpdLogEuclideanGaussianKernel)
    Xs = orc.synth_sphere(33, 3, seed=5)
    for k in (SphereGaussianKernel(3, range(3), beta_min=6.5, variance=1.7), SphereLaplaceKernel(3, range(3), variance=1.7)):
        assert np.array_equal(k.compute_Kdiag(Xs), np.full(33, 1.7))
    Xm = orc.synth_spd_mandel(33, 3, seed=6)
    for cls in (SpdAffineInvariantLaplaceKernel, SpdFrobeniusGaussianKernel, SpdLogEuclideanGaussianKernel):
        assert np.array_equal(cls(6, range(6), variance=0.3).compute_Kdiag(Xm), np.full(33, 0.3))
    assert np.array_equal(SpdAffineInvariantGaussianKernel(6, range(6), beta_min=0.5, variance=0.3).compute_Kdiag(Xm), np.full(33, 0.3))
    ks = SpdSteinGaussianKernel(6, range(6), beta=1.5, variance=0.3)
    assert np.max(np.abs(ks.compute_Kdiag(Xm) / np.diag(orc.spd_stein_K(Xm, beta=1.5, variance=0.3)) - 1)) < 1e-12
    assert np.max(np.abs(ks.compute_Kdiag(Xm) / np.diag(ks.compute_K_symm(Xm)) - 1)) < 1e-12


def test_gpr_constant_mean_and_two_outputs(cuda):
    from gaboflow_b200 import GPR, SphereGaussianKernel, Constant
    n = 400
    X = orc.synth_sphere(n, 5, seed=7)
    Y = np.random.default_rng(8).standard_normal((n, 2)) + 3.0
    Xs = orc.synth_sphere(50, 5, seed=9)
    k = SphereGaussianKernel(5, range(5), beta_min=1.0, beta=1.0, variance=1.2)
    m = GPR(X, Y, kern=k, mean_function=Constant(3.0))
    m.likelihood.variance = 0.1
    K = orc.sphere_gaussian_K(X, beta=2.0, variance=1.2)
    assert abs(m.compute_log_likelihood() / orc.gp_loglik(K, Y, 0.1, mean=3.0) - 1) < 1e-8
    mean, var = m.predict_f(Xs)
    mo, vo = orc.gp_predict(K, orc.sphere_gaussian_K(X, Xs, beta=2.0, variance=1.2), np.full(50, 1.2), Y, 0.1, mean=3.0)
    assert mean.shape == (50, 2) and var.shape == (50, 2)
    assert np.max(np.abs(mean - mo) / np.abs(mo)) < 1e-8
    assert np.max(np.abs(var - vo) / np.abs(vo)) < 1e-8
    ym, yv = m.predict_y(Xs)
    assert np.allclose(yv, var + 0.1, rtol=1e-12)


def test_potrf_on_leading_subview(cuda):
    from gaboflow_b200 import ops
    n, nc = 1500, 1024
    G = np.random.default_rng(10).standard_normal((n, 700))
    A = G @ G.T / 700 + np.eye(n)
    d = torch.from_numpy(A).cuda()
    keep = d.clone()
    ops.potrf_(d[:nc, :nc], 0.5)
    ref = sla.cholesky(A[:nc, :nc] + 0.5 * np.eye(nc), lower=True)
    assert np.max(np.abs(np.tril(d[:nc, :nc].cpu().numpy()) - ref)) < 1e-12
    assert torch.equal(d[nc:], keep[nc:]) and torch.equal(d[:nc, nc:], keep[:nc, nc:])


def test_empty_and_tiny(cuda):
    from gaboflow_b200 import ops, SpdAffineInvariantGaussianKernel
    e = ops.spd_prep(torch.zeros((0, 6), dtype=torch.float64, device='cuda'), 3)
    assert e.n == 0
    k = SpdAffineInvariantGaussianKernel(3, range(3), beta_min=0.5)
    one = orc.synth_spd_mandel(1, 2, seed=11)
    assert k.compute_K_symm(one).tolist() == [[1.0]]
    L = torch.tensor([[2.0]], dtype=torch.float64, device='cuda')
    b = torch.tensor([[4.0]], dtype=torch.float64, device='cuda')
    ops.potrf_(L, 0.0)
    ops.trsm_(L, b)
    assert abs(float(L.item()) - np.sqrt(2.0)) < 1e-15 and abs(float(b.item()) - 4.0 / np.sqrt(2.0)) < 1e-15


def test_gpr_failed_cholesky_does_not_leave_stale_cache(cuda):
    """factorise(ok) -> factorise(non-PD, raises) -> factorise(first parameters again) must refactorise; the
    buffer holds a half-factored matrix after the failure and the old cache key may not survive it."""
    from gaboflow_b200 import GPR, SphereGaussianKernel
    X = orc.synth_sphere(300, 3, seed=12)
    y = np.random.default_rng(13).standard_normal((300, 1))
    k = SphereGaussianKernel(3, range(3), beta_min=6.5, beta=1.0)
    m = GPR(X, y, kern=k)
    m.likelihood.variance = 1e-3
    ll0 = m.compute_log_likelihood()
    mean0, var0 = m.predict_f(X[:5])
    k.variance = -1.0                                       # K + s2 I is negative definite: the Cholesky must report it
    with pytest.raises(np.linalg.LinAlgError):
        m.compute_log_likelihood()
    k.variance = 1.0                                        # back to the parameters of the cached factor
    assert abs(m.compute_log_likelihood() - ll0) <= 1e-12 * abs(ll0)
    mean1, var1 = m.predict_f(X[:5])
    assert np.array_equal(mean0, mean1) and np.array_equal(var0, var1)


def test_spd_kernel_classes_reject_non_spd_input(cuda):
    """the per-point Cholesky info is read on the host-facing paths; a non-SPD matrix raises instead of
    silently producing NaN rows."""
    from gaboflow_b200 import GPR, SpdAffineInvariantGaussianKernel, SpdLogEuclideanGaussianKernel, SpdSteinGaussianKernel
    X = orc.synth_spd_mandel(20, 3, seed=1)
    X[7, :3] = [-1.0, 1.0, 1.0]
    for k in (SpdAffineInvariantGaussianKernel(6, range(6), beta_min=0.5), SpdLogEuclideanGaussianKernel(6, range(6)),
              SpdSteinGaussianKernel(6, range(6), beta=1.5)):
        with pytest.raises(ValueError, match='matrix 7'):
            k.compute_K_symm(X)
        with pytest.raises(ValueError, match='matrix 7'):
            GPR(X, np.zeros((20, 1)), kern=k).compute_log_likelihood()


def test_no_out_of_bounds_writes(cuda):
    """Canary frames around the output of every Gram / factorisation / solve entry point at ragged sizes and with leading
    dimensions larger than the row length (compute-sanitizer is not available on the pool, profiles/r02_sanitizer_note.txt):
    nothing outside the window the call was given may change."""
    from gaboflow_b200 import ops
    pad, CAN = 3, float(np.e)

    def framed(rows, cols, dtype=torch.float64):
        big = torch.full((rows + 2 * pad, cols + 2 * pad), CAN, dtype=dtype, device='cuda')
        return big, big[pad:pad + rows, pad:pad + cols]

    def frame_intact(big, rows, cols):
        b = big.cpu().numpy()
        m = np.ones(b.shape, dtype=bool)
        m[pad:pad + rows, pad:pad + cols] = False
        return bool(np.all(b[m] == b.dtype.type(CAN)))

    # point Grams: both FP64 engines and the FP32 tensor-core kernel, every uplo mode, a row range, a cross Gram
    for (n, dim) in [(389, 37), (131, 3), (257, 64)]:
        Xn = orc.synth_sphere(n, dim, seed=n)
        X = torch.from_numpy(Xn).cuda()
        X2 = torch.from_numpy(orc.synth_sphere(200, dim, seed=n + 1)).cuda()
        for eng in ('dmma', 'i8'):
            for uplo in ('full', 'lower', 'mirror'):
                big, out = framed(n, n)
                ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=1.5, uplo=uplo, engine=eng, out=out)
                assert frame_intact(big, n, n), (n, dim, eng, uplo)
            big, out = framed(n - 128, n)
            ops.points_gram(ops.SPHERE_LAPLACE, X, None, beta=1.5, row0=128, row1=n, engine=eng, out=out)
            assert frame_intact(big, n - 128, n), (n, dim, eng, 'rows')
            big, out = framed(n, 200)
            ops.points_gram(ops.SPHERE_GAUSSIAN, X, X2, beta=1.5, engine=eng, out=out)
            assert frame_intact(big, n, 200), (n, dim, eng, 'cross')
        for uplo in ('full', 'mirror'):
            big, out = framed(n, n, torch.float32)
            ops.points_gram(ops.SPHERE_GAUSSIAN, X.float(), None, beta=1.5, uplo=uplo, out=out)
            assert frame_intact(big, n, n), (n, dim, 'f32', uplo)
    # SPD pair Grams
    for d, n in [(2, 150), (6, 77)]:
        p = ops.spd_prep(torch.from_numpy(orc.synth_spd_mandel(n, d, seed=d)).cuda(), d, want_logm=True)
        for kind in (ops.SPD_AI_GAUSSIAN, ops.SPD_STEIN):
            for uplo in ('full', 'mirror'):
                big, out = framed(n, n)
                ops.spd_gram(kind, p, None, beta=0.7, out=out, uplo=uplo)
                assert frame_intact(big, n, n), (d, n, kind, uplo)
    # factorisation in a framed window (also: strictly upper triangle untouched), solves with 1 / 3 / 70 right-hand sides
    n = 700
    G = np.random.default_rng(3).standard_normal((n, 900))
    A = G @ G.T / 900 + np.eye(n)
    big, win = framed(n, n)
    win.copy_(torch.from_numpy(A).cuda())
    ops.potrf_(win, 0.25)
    assert frame_intact(big, n, n)
    L = win.cpu().numpy()
    assert np.array_equal(np.triu(L, 1), np.triu(A, 1))
    assert np.max(np.abs(np.tril(L) - sla.cholesky(A + 0.25 * np.eye(n), lower=True))) < 1e-12
    Lc = torch.from_numpy(np.tril(L)).cuda()
    for nrhs in (1, 3, 70):
        big, B = framed(n, nrhs)
        Bn = np.random.default_rng(nrhs).standard_normal((n, nrhs))
        B.copy_(torch.from_numpy(Bn).cuda())
        ops.trsm_(Lc, B)
        assert frame_intact(big, n, nrhs), nrhs
        assert np.max(np.abs(np.tril(L) @ B.cpu().numpy() - Bn)) < 1e-10


def test_fp32_mode_of_every_kernel_class(cuda):
    """north_star: within 1e-5 in FP32 mode.  `set_float_type('float32')` on any kernel class: float32 in, float32 out; sphere
    and squared-distance kernels on the TF32 tcgen05 Gram kernel, SPD pair kernels with FP64 arithmetic inside."""
    from gaboflow_b200 import (SphereGaussianKernel, SphereLaplaceKernel, SpdAffineInvariantGaussianKernel,
                               SpdAffineInvariantLaplaceKernel, SpdSteinGaussianKernel, SpdFrobeniusGaussianKernel,
                               SpdLogEuclideanGaussianKernel, GPR)
    X = orc.synth_sphere(300, 51, seed=5)
    X2 = orc.synth_sphere(130, 51, seed=6)
    for k, ref in [(SphereGaussianKernel(51, range(51), beta_min=0.5, beta=1.5, variance=1.2), lambda a, b: orc.sphere_gaussian_K(a, b, beta=2.0, variance=1.2)),
                   (SphereLaplaceKernel(51, range(51), beta=2.0, variance=0.8), lambda a, b: orc.sphere_laplace_K(a, b, beta=2.0, variance=0.8))]:
        k.set_float_type('float32')
        K = k.K(X.astype(np.float32))
        assert K.dtype == torch.float32
        Kn = K.cpu().numpy().astype(np.float64)
        Ko = ref(X, None)
        off = ~np.eye(300, dtype=bool)
        # 1e-5 of the entry, plus the conditioning of acos at FP32 input rounding (6e-8 on the dot product)
        assert np.max(np.abs(Kn - Ko)[off] / Ko[off]) < 2e-5
        assert np.array_equal(Kn, Kn.T) and np.all(np.diag(Kn) == np.float32(float(k.variance)))
        K12 = k.compute_K(X.astype(np.float32), X2.astype(np.float32))
        assert K12.dtype == np.float32 and np.max(np.abs(K12 / ref(X, X2) - 1)) < 2e-5
        assert k.Kdiag(X).dtype == torch.float32
        with pytest.raises(ValueError):
            GPR(X, np.zeros((300, 1)), kern=k)
        k.set_float_type('float64')
        assert k.K(X).dtype == torch.float64
    V = orc.synth_spd_mandel(150, 3, seed=7)
    V2 = orc.synth_spd_mandel(60, 3, seed=8)
    cases = [(SpdAffineInvariantGaussianKernel(6, range(6), beta_min=0.2, beta=0.5), lambda a, b: orc.spd_ai_gaussian_K(a, b, beta=0.7)),
             (SpdAffineInvariantLaplaceKernel(6, range(6), beta=0.7), lambda a, b: orc.spd_ai_laplace_K(a, b, beta=0.7)),
             (SpdSteinGaussianKernel(6, range(6), beta=1.5), lambda a, b: orc.spd_stein_K(a, b, beta=1.5)),
             (SpdFrobeniusGaussianKernel(6, range(6), beta=0.3), lambda a, b: orc.spd_frobenius_K(a, b, beta=0.3)),
             (SpdLogEuclideanGaussianKernel(6, range(6), beta=0.6), lambda a, b: orc.spd_logeuclid_K(a, b, beta=0.6))]
    for k, ref in cases:
        k.set_float_type('float32')
        K = k.K(V.astype(np.float32))
        K12 = k.K(V.astype(np.float32), V2.astype(np.float32))
        assert K.dtype == torch.float32 and K12.dtype == torch.float32
        # the oracle sees the same float32-rounded inputs
        Vr, V2r = V.astype(np.float32).astype(np.float64), V2.astype(np.float32).astype(np.float64)
        Ko, K12o = ref(Vr, None), ref(Vr, V2r)
        off = ~np.eye(150, dtype=bool)
        # pair kernels: FP64 inside, only the float32 rounding of the result remains; squared-distance kernels: TF32x3 tiles,
        # relative error of exp(-beta d^2) = beta d^2 * 1e-6
        tol = 2e-5 if isinstance(k, (SpdFrobeniusGaussianKernel, SpdLogEuclideanGaussianKernel)) else 2e-7
        assert np.max(np.abs(K.cpu().numpy() - Ko)[off] / Ko[off]) < tol, type(k).__name__
        assert np.max(np.abs(K12.cpu().numpy() / K12o - 1)) < tol, type(k).__name__
        assert k.Kdiag(V.astype(np.float32)).dtype == torch.float32
