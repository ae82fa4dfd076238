"""FP64 sphere Gram with the inner products on the integer tensor cores (csrc/points_gram_i8.cu, `engine='i8'`, the default for
large sphere Grams): parity against the NumPy oracle (sphere_utils_tf.py:11-60, kernels_sphere_tf.py:59-88, 179-204) at the
north_star tolerance 1e-12, in every mode the DMMA kernel supports (full / lower / mirror, cross Gram, row ranges, block-cyclic
rows), on ragged sizes, every dimension class, points that are not unit vectors (common power-of-two scaling), clamped /
antipodal dot products, and the exponent underflow range."""
import numpy as np
import pytest
import torch

from oracle import gabo_oracle as orc

pytestmark = pytest.mark.gpu

RTOL64 = 1e-12
SENTINEL = -7.0


def _oracle(kind, X, X2, beta, variance):
    from gaboflow_b200 import ops
    f = orc.sphere_gaussian_K if kind == ops.SPHERE_GAUSSIAN else orc.sphere_laplace_K
    return f(X, X2, beta=beta, variance=variance)


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.mark.parametrize('kindname', ['gauss', 'laplace'])
@pytest.mark.parametrize('uplo', ['full', 'lower', 'mirror'])
@pytest.mark.parametrize('n,dim', [(1, 3), (129, 3), (300, 17), (1000, 51), (513, 64), (2500, 33)])
def test_i8_symmetric_modes_vs_oracle(cuda, kindname, uplo, n, dim):
    """K(X, X): written entries equal the oracle to 1e-12 (off the diagonal; the diagonal is exactly `variance`), entries a mode
    does not own stay untouched, the mirrored matrix is bitwise symmetric."""
    from gaboflow_b200 import ops
    kind = ops.SPHERE_GAUSSIAN if kindname == 'gauss' else ops.SPHERE_LAPLACE
    Xn = orc.synth_sphere(n, dim, seed=100 * n + dim)
    X = torch.from_numpy(Xn).to(cuda)
    out = torch.full((n, n), SENTINEL, dtype=torch.float64, device=cuda)
    ops.points_gram(kind, X, None, beta=2.0, variance=1.3, uplo=uplo, engine='i8', out=out)
    K = out.cpu().numpy()
    Ko = _oracle(kind, Xn, None, 2.0, 1.3)
    written = K != SENTINEL
    off = ~np.eye(n, dtype=bool)
    if uplo == 'lower':
        # tiles on / below the diagonal are written (whole 128 x 128 tiles), nothing above the diagonal tiles
        ti, tj = np.indices((n, n)) // 128
        assert np.array_equal(written, tj <= ti)
    else:
        assert written.all()
    m = written & off
    if m.any():
        assert _rel(K[m], Ko[m]) < RTOL64
    assert np.all(np.diag(K) == 1.3)
    if uplo == 'mirror':
        assert np.array_equal(K, K.T)


@pytest.mark.parametrize('n1,n2,dim,row0,row1', [(5, 7, 3, 0, 5), (130, 257, 51, 0, 130), (1500, 900, 51, 256, 1300), (700, 1, 20, 128, 700),
                                                 (640, 640, 64, 384, 640)])
def test_i8_cross_gram_and_row_ranges(cuda, n1, n2, dim, row0, row1):
    from gaboflow_b200 import ops
    Xn, X2n = orc.synth_sphere(n1, dim, seed=1), orc.synth_sphere(n2, dim, seed=2)
    X, X2 = torch.from_numpy(Xn).to(cuda), torch.from_numpy(X2n).to(cuda)
    for kind in (ops.SPHERE_GAUSSIAN, ops.SPHERE_LAPLACE):
        K = ops.points_gram(kind, X, X2, beta=3.0, variance=0.7, engine='i8', row0=row0, row1=row1).cpu().numpy()
        assert K.shape == (row1 - row0, n2)
        assert _rel(K, _oracle(kind, Xn, X2n, 3.0, 0.7)[row0:row1]) < RTOL64


def test_i8_agrees_with_dmma_engine(cuda):
    """Both engines produce the same Gram to ~1e-14 (they round the inner product differently) and the same untouched pattern."""
    from gaboflow_b200 import ops
    n = 1100
    X = torch.from_numpy(orc.synth_sphere(n, 51, seed=9)).to(cuda)
    for uplo in ('full', 'lower', 'mirror'):
        a = ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, uplo=uplo, engine='dmma',
                            out=torch.full((n, n), SENTINEL, dtype=torch.float64, device=cuda))
        b = ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, uplo=uplo, engine='i8',
                            out=torch.full((n, n), SENTINEL, dtype=torch.float64, device=cuda))
        assert torch.equal(a == SENTINEL, b == SENTINEL)
        assert float(((a - b).abs() / a.abs()).max()) < 1e-13


@pytest.mark.parametrize('scale', [2.0 ** -30, 0.37, 1.0, 5.0, 2.0 ** 40])
def test_i8_common_scaling_of_the_points(cuda, scale):
    """The slicing uses one power-of-two exponent per point set, so inputs need not be unit vectors: K(sX, X2 / s) has the dot
    products of K(X, X2) exactly when s is a power of two, and to rounding otherwise."""
    from gaboflow_b200 import ops
    Xn, X2n = orc.synth_sphere(400, 51, seed=3), orc.synth_sphere(300, 51, seed=4)
    A, B = Xn * scale, X2n / scale
    K = ops.points_gram(ops.SPHERE_GAUSSIAN, torch.from_numpy(A).to(cuda), torch.from_numpy(B).to(cuda), beta=2.0, engine='i8').cpu().numpy()
    Ko = _oracle(ops.SPHERE_GAUSSIAN, A, B, 2.0, 1.0)
    assert _rel(K, Ko) < RTOL64


def test_i8_clamped_antipodal_and_duplicate_points(cuda):
    """Dot products at and beyond +-1 (duplicates, antipodes, slightly long vectors): the clamp of sphere_utils_tf.py:59 applies;
    d = pi gives exp(-beta pi^2), d = 0 gives `variance`."""
    from gaboflow_b200 import ops
    Xn = orc.synth_sphere(200, 51, seed=5)
    X2n = np.concatenate([Xn[:50], -Xn[50:100], Xn[100:150] * (1 + 1e-9), -Xn[150:] * (1 + 1e-9)])
    K = ops.points_gram(ops.SPHERE_GAUSSIAN, torch.from_numpy(Xn).to(cuda), torch.from_numpy(X2n).to(cuda), beta=0.5, variance=2.0,
                        engine='i8').cpu().numpy()
    Ko = _oracle(ops.SPHERE_GAUSSIAN, Xn, X2n, 0.5, 2.0)
    d = np.arange(200)
    # entries with |dot| ~ 1 are ill conditioned in the dot product (dK/dt ~ 1/sqrt(1-t^2)): compare the distance instead
    dist = np.sqrt(-np.log(K[d, d] / 2.0) / 0.5)
    dist_o = np.sqrt(-np.log(Ko[d, d] / 2.0) / 0.5)
    assert np.max(np.abs(dist - dist_o)) < 2e-7          # acos near +-1: sqrt(2 * 1e-15) from the rounding of the dot product
    assert np.all(K[d[100:150], d[100:150]] == 2.0) and np.all(np.abs(K[d[150:], d[150:]] / (2.0 * np.exp(-0.5 * np.pi ** 2)) - 1) < 4e-15)
    off = ~np.eye(200, dtype=bool)
    assert _rel(K[off], Ko[off]) < RTOL64


@pytest.mark.parametrize('variance', [1e-6, 1e-3, 1.0, 1e3])
def test_i8_exp_underflow_range(cuda, variance):
    """beta d^2 swept through 650..745 (round-1 bug: the exponent field wrapped for small variances): results stay finite, positive
    or zero, and equal the oracle where it is a normal number."""
    from gaboflow_b200 import ops
    Xn, X2n = orc.synth_sphere(130, 8, seed=6), orc.synth_sphere(140, 8, seed=7)
    d = orc.sphere_distance(Xn, X2n)
    for target in (650.0, 690.0, 700.0, 720.0, 745.0):
        beta = target / float(np.median(d) ** 2)
        K = ops.points_gram(ops.SPHERE_GAUSSIAN, torch.from_numpy(Xn).to(cuda), torch.from_numpy(X2n).to(cuda), beta=beta,
                            variance=variance, engine='i8').cpu().numpy()
        Ko = _oracle(ops.SPHERE_GAUSSIAN, Xn, X2n, beta, variance)
        assert np.all(np.isfinite(K)) and np.all(K >= 0.0)
        normal = Ko > 1e-290
        # d(log K) = 2 beta d / sin(d) * d(dot): with beta d^2 ~ 700 the ~1e-15 absolute error of either side's inner product
        # (oracle: BLAS rounding; here: the dropped slice pairs, < 6e-15 worst case) is amplified ~1000-fold
        assert _rel(K[normal], Ko[normal]) < 1e-11
        assert np.all(K[~normal] <= 1e-289)


def test_i8_block_cyclic_rows_equal_single_gpu(cuda):
    """cyc_world > 0: every emulated rank's rows are bit-identical to the rows of the single-GPU Gram."""
    from gaboflow_b200 import ops
    from gaboflow_b200.dist import BlockCyclic
    n = 512 * 7                   # the cyclic mode takes whole 512-row panels (n % 512 == 0, status -15 otherwise)
    X = torch.from_numpy(orc.synth_sphere(n, 51, seed=8)).to(cuda)
    from gaboflow_b200 import _lib
    with pytest.raises(_lib.GaboError):
        ops.points_gram_cyclic_rows_i8(ops.SPHERE_GAUSSIAN, X[:n - 100], torch.empty((n, n), dtype=torch.float64, device=cuda), 2, 0)
    full = ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, uplo='full', engine='i8')
    for W in (2, 3):
        for r in range(W):
            lay = BlockCyclic(n, 512, W, r)
            buf = torch.full((lay.local_rows, n), float('nan'), dtype=torch.float64, device=cuda)
            ops.points_gram_cyclic_rows_i8(ops.SPHERE_GAUSSIAN, X, buf, W, r, beta=2.0)
            assert torch.equal(buf, full[lay.global_rows().to(cuda)])


def test_i8_engine_selection_and_errors(cuda):
    from gaboflow_b200 import ops
    assert ops.gram_engine(ops.SPHERE_GAUSSIAN, torch.float64, 51, 65536, 65536) == 'i8'
    assert ops.gram_engine(ops.SPHERE_GAUSSIAN, torch.float64, 51, 100, 100) == 'dmma'          # launch-bound sizes
    assert ops.gram_engine(ops.SPHERE_GAUSSIAN, torch.float64, 120, 65536, 65536) == 'dmma'      # dim > 64
    assert ops.gram_engine(ops.SPHERE_GAUSSIAN, torch.float32, 51, 65536, 65536) == 'dmma'
    assert ops.gram_engine(ops.SQDIST_GAUSSIAN, torch.float64, 21, 65536, 65536) == 'dmma'
    X = torch.from_numpy(orc.synth_sphere(64, 120, seed=1)).to(cuda)
    with pytest.raises(ValueError):
        ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, engine='i8')
    # non-finite coordinates do not hang or poison other rows: their rows / columns are NaN-free garbage-free zeros of the slicing,
    # every other entry is exact
    Xn = orc.synth_sphere(300, 10, seed=2)
    Xbad = Xn.copy()
    Xbad[7, 3] = np.nan
    K = ops.points_gram(ops.SPHERE_GAUSSIAN, torch.from_numpy(Xbad).to(cuda), None, beta=1.0, engine='i8').cpu().numpy()
    Ko = orc.sphere_gaussian_K(Xn, beta=1.0)
    keep = np.ones(300, dtype=bool)
    keep[7] = False
    m = np.outer(keep, keep) & ~np.eye(300, dtype=bool)
    assert _rel(K[m], Ko[m]) < RTOL64


def test_i8_default_engine_through_kernel_class(cuda):
    """SphereGaussianKernel.K on a large point set goes through the int8 engine by default and matches the oracle."""
    from gaboflow_b200 import SphereGaussianKernel, ops
    n = 2304
    assert ops.gram_engine(ops.SPHERE_GAUSSIAN, torch.float64, 51, n, n) == 'i8'
    Xn = orc.synth_sphere(n, 51, seed=11)
    k = SphereGaussianKernel(51, range(51), beta_min=0.0, beta=2.0)
    K = k.compute_K_symm(Xn)
    Ko = orc.sphere_gaussian_K(Xn, beta=2.0)
    off = ~np.eye(n, dtype=bool)
    assert _rel(K[off], Ko[off]) < RTOL64
    assert np.array_equal(K, K.T) and np.all(np.diag(K) == 1.0)


def test_i8_no_out_of_bounds_writes(cuda):
    """Canary-padded output buffers (compute-sanitizer is not available on the pool): the int8 kernel writes nothing outside
    the [row1-row0, n2] window it was given, on ragged sizes, with a leading dimension larger than n2, in every mode."""
    from gaboflow_b200 import ops
    pad = 3
    for (n, dim, uplo, row0, row1) in [(389, 37, 'full', 128, 389), (389, 37, 'mirror', 0, 389), (260, 64, 'lower', 0, 260),
                                       (1, 3, 'full', 0, 1), (300, 51, 'full', 0, 1)]:
        Xn = orc.synth_sphere(n, dim, seed=n + dim)
        X = torch.from_numpy(Xn).to(cuda)
        rows = row1 - row0
        big = torch.full((rows + 2 * pad, n + 2 * pad), np.pi, dtype=torch.float64, device=cuda)
        out = big[pad:pad + rows, pad:pad + n]
        ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=1.5, uplo=uplo, row0=row0, row1=row1, engine='i8', out=out)
        torch.cuda.synchronize()
        b = big.cpu().numpy()
        frame = np.ones_like(b, dtype=bool)
        frame[pad:pad + rows, pad:pad + n] = False
        assert np.all(b[frame] == np.pi), (n, dim, uplo)
        Ko = orc.sphere_gaussian_K(Xn, beta=1.5)[row0:row1]
        got = b[pad:pad + rows, pad:pad + n]
        m = got != np.pi
        if uplo != 'lower':
            assert m.all()
        off = m & (np.arange(row0, row1)[:, None] != np.arange(n)[None, :])
        if off.any():
            assert _rel(got[off], Ko[off]) < RTOL64
