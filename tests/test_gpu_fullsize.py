"""GPU, BASELINE.json full sizes: parity through sampled blocks against the oracle and size-independent properties
(symmetry, exact diagonal, factor residual on sampled tiles, solve residual, log-det consistency).
SURVEY.md section 8(d): 16 sampled 256-row blocks + the full diagonal instead of a 34 GB compare."""
import numpy as np
import pytest
import torch

from oracle import gabo_oracle as orc

pytestmark = pytest.mark.gpu


def test_sphere_gram_n65536_sampled_blocks(cuda):
    from gaboflow_b200 import ops
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip('needs 40 GB of free HBM')
    N, D = 65536, 51
    X = orc.synth_sphere(N, D, seed=20240)
    Xd = torch.from_numpy(X).cuda()
    K = torch.empty((N, N), dtype=torch.float64, device='cuda')
    ops.points_gram(ops.SPHERE_GAUSSIAN, Xd, None, beta=2.0, variance=1.0, out=K, uplo='mirror')
    rng = np.random.default_rng(1)
    for r0 in rng.choice(N // 256, size=16, replace=False) * 256:
        blk = K[r0:r0 + 256].cpu().numpy()
        ref = orc.sphere_gaussian_K(X[r0:r0 + 256], X, beta=2.0)
        off = np.ones_like(ref, dtype=bool)
        off[np.arange(256), r0 + np.arange(256)] = False
        assert np.max(np.abs(blk - ref)[off] / ref[off]) < 1e-12
        assert np.all(blk[np.arange(256), r0 + np.arange(256)] == 1.0)
    assert bool(torch.all(torch.diagonal(K) == 1.0))
    # symmetry of the mirrored store on sampled tile pairs (bitwise)
    for _ in range(8):
        i0, j0 = (rng.integers(0, N // 512, 2) * 512).tolist()
        assert torch.equal(K[i0:i0 + 512, j0:j0 + 512], K[j0:j0 + 512, i0:i0 + 512].t())
    # row-block sharding reproduces the same entries bitwise (what each rank of a multi-GPU run produces)
    part = ops.points_gram(ops.SPHERE_GAUSSIAN, Xd, None, beta=2.0, row0=40960, row1=40960 + 512, uplo='full')
    assert torch.equal(torch.tril(part[:, :40960 + 512], 40960), torch.tril(K[40960:40960 + 512, :40960 + 512], 40960))
    del K, part
    torch.cuda.empty_cache()


def test_cholesky_n32768_properties(cuda):
    from gaboflow_b200 import ops
    free, _ = torch.cuda.mem_get_info()
    if free < 30e9:
        pytest.skip('needs 30 GB of free HBM')
    N = 32768
    X = orc.synth_sphere(N, 51, seed=20240)
    Xd = torch.from_numpy(X).cuda()
    K = torch.empty((N, N), dtype=torch.float64, device='cuda')
    ops.points_gram(ops.SPHERE_GAUSSIAN, Xd, None, beta=2.0, out=K, uplo='mirror')
    L = K.clone()
    info = ops.potrf_(L, 1e-2)
    assert int(info.item()) == 0
    assert torch.equal(torch.triu(L, 1), torch.triu(K, 1))            # strictly upper triangle untouched
    rng = np.random.default_rng(2)
    # residual of L L^T = K + sigma^2 I on sampled 512 x 512 tiles (row block i, column block j <= i)
    worst = 0.0
    for _ in range(6):
        bi, bj = sorted(rng.integers(0, N // 512, 2).tolist(), reverse=True)
        i0, j0 = bi * 512, bj * 512
        Li = torch.tril(L[i0:i0 + 512, :j0 + 512], diagonal=i0)      # rows of L, masked at the GLOBAL diagonal
        Lj = torch.tril(L[j0:j0 + 512, :j0 + 512], diagonal=j0)
        prod = Li @ Lj.t()
        ref = K[i0:i0 + 512, j0:j0 + 512].clone()
        if bi == bj:
            ref.diagonal().add_(1e-2)
            prod = torch.tril(prod); ref = torch.tril(ref)
        worst = max(worst, float((prod - ref).abs().max().item()))
    assert worst < 5e-13
    # solve + log-likelihood against an independent route: alpha = L^-1 y  =>  L alpha = y ; ||alpha||^2 = y^T (K+s I)^-1 y
    y = torch.from_numpy(np.random.default_rng(20241).standard_normal((N, 1))).cuda()
    a = y.clone()
    ops.trsm_(L, a)
    Lt = torch.tril(L)
    assert float((Lt @ a - y).abs().max().item()) < 1e-9
    ll = float(ops.gp_loglik(L, a).item())
    half_logdet = float(ops.logdet_half(L).item())
    ref_ll = -0.5 * N * np.log(2 * np.pi) - half_logdet - 0.5 * float((a * a).sum().item())
    assert abs(ll / ref_ll - 1) < 1e-12
    assert abs(half_logdet - float(torch.log(torch.diagonal(L)).sum().item())) < 1e-8 * abs(half_logdet)
    # leading-block consistency: the factor of the leading 4096 block equals the leading block of the factor, and its
    # log-likelihood matches the CPU oracle (1e-8 relative, north_star)
    n1 = 4096
    ll_small = float(ops.gp_loglik(L[:n1, :n1], a[:n1]).item())
    ref_small = orc.gp_loglik(orc.sphere_gaussian_K(X[:n1], beta=2.0), y[:n1].cpu().numpy(), 1e-2)
    assert abs(ll_small / ref_small - 1) < 1e-8
    del K, L, Lt
    torch.cuda.empty_cache()


def test_spd_gram_n16384_sampled_blocks(cuda):
    from gaboflow_b200 import ops
    N = 16384
    Xm = orc.synth_spd_mandel(N, 6, seed=20242)
    p = ops.spd_prep(torch.from_numpy(Xm).cuda(), 6, want_logm=True)
    assert int(p.info.abs().sum().item()) == 0
    K = ops.spd_gram(ops.SPD_AI_GAUSSIAN, p, None, beta=1.0, uplo='mirror')
    Kle = ops.points_gram(ops.SQDIST_GAUSSIAN, p.logm, None, beta=1.0, uplo='mirror')
    rng = np.random.default_rng(3)
    for r0 in rng.choice(N // 64, size=4, replace=False) * 64:
        cols = rng.choice(N, size=2048, replace=False)
        ref = orc.spd_ai_gaussian_K(Xm[r0:r0 + 64], Xm[cols], beta=1.0)
        got = K[r0:r0 + 64][:, torch.from_numpy(cols).cuda()].cpu().numpy()
        same = (cols[None, :] == (r0 + np.arange(64))[:, None])
        assert np.max(np.abs(got - ref)[~same] / ref[~same]) < 1e-12
        refle = orc.spd_logeuclid_K(Xm[r0:r0 + 64], Xm[cols], beta=1.0)
        gotle = Kle[r0:r0 + 64][:, torch.from_numpy(cols).cuda()].cpu().numpy()
        assert np.max(np.abs(gotle - refle)[~same] / refle[~same]) < 1e-12
    assert torch.equal(K, K.t()) and bool(torch.all(torch.diagonal(K) == 1.0))
