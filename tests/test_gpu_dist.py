"""GPU: the CUDA building blocks of the multi-GPU factorisation against torch-CPU references, and the distributed driver
with world_size 1 (same code path, no collectives).  The world_size>1 logic is covered on CPU by test_dist_gloo.py and on
2+ GPUs by tools/dist_check.py (run under torchrun)."""
import numpy as np
import pytest
import scipy.linalg as sla
import torch

pytestmark = pytest.mark.gpu


def test_trsm_right(cuda):
    from gaboflow_b200 import ops
    rng = np.random.default_rng(0)
    for m, nb in [(70, 128), (700, 512), (300, 200), (5, 17)]:
        G = rng.standard_normal((nb, nb)); L = np.linalg.cholesky(G @ G.T / nb + np.eye(nb))
        A = rng.standard_normal((m, nb))
        ref = sla.solve_triangular(L, A.T, lower=True).T
        big = torch.from_numpy(np.concatenate([A, np.zeros((m, 6))], axis=1)).cuda()     # non-trivial leading dimension
        ops.trsm_right_(big[:, :nb], torch.from_numpy(L).cuda())
        assert np.max(np.abs(big[:, :nb].cpu().numpy() - ref)) / np.max(np.abs(ref)) < 1e-12


@pytest.mark.parametrize('world,rank', [(1, 0), (2, 1), (4, 2)])
def test_syrk_cyclic_mask(cuda, world, rank):
    from gaboflow_b200 import ops
    rng = np.random.default_rng(world + rank)
    nb, slots, k, n = 128, 3, 128, 900
    m = slots * nb
    A, B, C = rng.standard_normal((m, k)), rng.standard_normal((n, k)), rng.standard_normal((m, n))
    row_base, col_base = 2 * world * nb, 256
    r = np.arange(m)
    grow = row_base + ((r // nb) * world + rank) * nb + r % nb
    mask = (col_base + np.arange(n))[None, :] <= grow[:, None]
    ref = np.where(mask, C - A @ B.T, C)
    Cd = torch.from_numpy(C).cuda()
    ops.syrk_cyclic_(Cd, torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), row_base, col_base, nb, world, rank)
    got = Cd.cpu().numpy()
    assert np.array_equal(got[~mask], C[~mask])
    assert np.max(np.abs(got - ref)) < 1e-11


def test_solve_update(cuda):
    from gaboflow_b200 import ops
    rng = np.random.default_rng(3)
    for m, k, nrhs in [(500, 128, 1), (777, 512, 3), (300, 256, 40)]:
        Lp, X, Y = rng.standard_normal((m, k)), rng.standard_normal((k, nrhs)), rng.standard_normal((m, nrhs))
        Yd = torch.from_numpy(Y).cuda()
        ops.solve_update_(Yd, torch.from_numpy(Lp).cuda(), torch.from_numpy(X).cuda())
        assert np.max(np.abs(Yd.cpu().numpy() - (Y - Lp @ X))) < 1e-11


def test_block_cyclic_driver_world1(cuda):
    from gaboflow_b200 import SphereGaussianKernel, ops
    from gaboflow_b200.dist import BlockCyclic, gram_block_cyclic, potrf_block_cyclic_, solve_lower_block_cyclic_, gp_loglik_block_cyclic
    from oracle import gabo_oracle as orc
    n = 1536
    X = orc.synth_sphere(n, 51, seed=4)
    y = np.random.default_rng(5).standard_normal((n, 1))
    Xd = torch.from_numpy(X).cuda()
    k = SphereGaussianKernel(51, range(51), beta_min=0.0, beta=2.0)
    lay = BlockCyclic(n, 512, 1, 0)
    A = gram_block_cyclic(lambda r0, r1, out: k.K(Xd, uplo='lower', row0=r0, row1=r1, out=out), lay, device='cuda')
    assert potrf_block_cyclic_(A, lay, 1e-2, n_factor=1024) == 0
    V = torch.from_numpy(y[:1024]).cuda()
    solve_lower_block_cyclic_(A, V, lay, n_factor=1024)
    ll = gp_loglik_block_cyclic(A, V, lay, n_factor=1024)
    ref = orc.gp_loglik(orc.sphere_gaussian_K(X[:1024], beta=2.0), y[:1024], 1e-2)
    assert abs(ll / ref - 1) < 1e-8


@pytest.mark.parametrize('mod', [0, 64, 128, 256])
@pytest.mark.parametrize('world', [1, 2, 3])
def test_cyclic_mirrored_gram_single_process(cuda, world, mod):
    """gabo_points_gram_cyclic_hybrid_f64 with all `world` ranks emulated on one GPU (every 'peer' buffer is local memory):
    same kernel and addressing as the NVLink run, launched once per rank.  mod > 0: mod/256 of the cross-rank upper blocks
    are recomputed by the row owner instead of mirrored; either way every entry is written (NaN pre-fill) with the same bits."""
    from gaboflow_b200 import ops
    from gaboflow_b200.dist import BlockCyclic
    from oracle import gabo_oracle as orc
    n = 512 * 7
    X = torch.from_numpy(orc.synth_sphere(n, 51, seed=8)).cuda()
    lays = [BlockCyclic(n, 512, world, r) for r in range(world)]
    bufs = [torch.full((l.local_rows, n), float('nan'), dtype=torch.float64, device='cuda') for l in lays]
    ptrs = [b.data_ptr() for b in bufs]
    for r in range(world):
        ops.points_gram_cyclic(ops.SPHERE_GAUSSIAN, X, ptrs, n, world, r, beta=2.0, variance=1.0, recompute_frac256=mod)
    torch.cuda.synchronize()
    full = ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, uplo='mirror', engine='dmma')   # same kernel family: bit-identical
    for r in range(world):
        rows = lays[r].global_rows().cuda()
        assert torch.equal(bufs[r], full[rows])


@pytest.mark.parametrize('split,direct', [(0, 0), (1, 0), (2, 64), (1, 128), (1, 256)])
@pytest.mark.parametrize('frac', [0, 48, 128, 256])
@pytest.mark.parametrize('world', [1, 2, 3, 8])
def test_cyclic_staged_gram_single_process(cuda, world, frac, split, direct):
    """gabo_points_gram_cyclic_staged_f64 with all `world` ranks emulated on one GPU: mirror blocks bound for other ranks go
    through the local staging buffer and copy-engine 2-D copies released by the kernel's per-chunk flags (`direct`/256 of the
    slots: remote stores from the kernel instead).  Every entry is written
    (NaN pre-fill) with the bits of the single-GPU mirrored Gram; run twice to exercise the epoch / counter reset."""
    from gaboflow_b200 import ops
    from gaboflow_b200.dist import BlockCyclic
    from oracle import gabo_oracle as orc
    n = 512 * 11
    X = torch.from_numpy(orc.synth_sphere(n, 51, seed=8)).cuda()
    lays = [BlockCyclic(n, 512, world, r) for r in range(world)]
    bufs = [torch.full((max(l.local_rows, 1), n), float('nan'), dtype=torch.float64, device='cuda') for l in lays]
    ptrs = [b.data_ptr() for b in bufs]
    states = [ops.StagedGramState(n, world, r, frac, split, 'cuda', direct_frac256=direct) for r in range(world)]
    full = ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, uplo='mirror', engine='dmma')   # same kernel family: bit-identical
    for rep in range(2):
        for b in bufs:
            b.fill_(float('nan'))
        for r in range(world):
            ops.points_gram_cyclic_staged(ops.SPHERE_GAUSSIAN, X, ptrs, n, world, r, states[r], beta=2.0, variance=1.0)
        torch.cuda.synchronize()
        for r in range(world):
            if lays[r].local_rows == 0:
                continue
            rows = lays[r].global_rows().cuda()
            assert torch.equal(bufs[r][:lays[r].local_rows], full[rows]), (rep, r)


@pytest.mark.parametrize('kind_name', ['ai_gauss', 'ai_laplace', 'stein'])
@pytest.mark.parametrize('world', [1, 2, 3, 8])
def test_cyclic_spd_gram_single_process(cuda, world, kind_name):
    """gabo_spd_gram_cyclic_f64 with all `world` ranks emulated on one GPU (every 'peer' buffer is local memory): same kernel
    and addressing as the NVLink run, launched once per rank.  Every entry is written exactly once (NaN pre-fill), with the
    bits of the single-GPU mirrored Gram; n is not a multiple of the panel (ragged last panel)."""
    from gaboflow_b200 import ops
    from gaboflow_b200.dist import BlockCyclic
    from oracle import gabo_oracle as orc
    n = 512 * 3 + 200
    kind, beta = {'ai_gauss': (ops.SPD_AI_GAUSSIAN, 1.0), 'ai_laplace': (ops.SPD_AI_LAPLACE, 0.7), 'stein': (ops.SPD_STEIN, 3.0)}[kind_name]
    prep = ops.spd_prep(torch.from_numpy(orc.synth_spd_mandel(n, 6, seed=31)).cuda(), 6)
    lays = [BlockCyclic(n, 512, world, r) for r in range(world)]
    bufs = [torch.full((max(l.local_rows, 1), n), float('nan'), dtype=torch.float64, device='cuda') for l in lays]
    ptrs = [b.data_ptr() for b in bufs]
    for r in range(world):
        ops.spd_gram_cyclic(kind, prep, ptrs, n, world, r, beta=beta, variance=1.2)
    torch.cuda.synchronize()
    full = ops.spd_gram(kind, prep, None, beta=beta, variance=1.2, uplo='mirror')
    for r in range(world):
        if lays[r].local_rows == 0:
            continue
        rows = lays[r].global_rows().cuda()
        assert torch.equal(bufs[r][:lays[r].local_rows], full[rows])


def test_flag_signal_and_stream_wait(cuda):
    """gabo_flag_signal / gabo_flag_wait_geq: a stream stalled on a flag resumes when another stream raises it."""
    from gaboflow_b200 import ops
    flags = torch.zeros(8, dtype=torch.int32, device='cuda')
    out = torch.zeros(1, dtype=torch.float64, device='cuda')
    # every kernel used below is launched once first: a first launch loads its module, which synchronises the context and
    # would block the host for as long as the waiter stream is stalled
    out.add_(1.0).zero_()
    ops.flag_signal([flags.data_ptr() + 28], 1)
    float(out.item()); flags.cpu()
    torch.cuda.synchronize()
    waiter, signaller = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(waiter):
        ops.flag_wait_geq(flags.data_ptr() + 4, 7)
        out.add_(1.0)                                  # must not run before the flag is raised
    with torch.cuda.stream(signaller):
        ops.flag_signal([flags.data_ptr() + 4, flags.data_ptr() + 8], 6)      # too small: the waiter stays blocked
    signaller.synchronize()
    assert not waiter.query()
    assert float(out.item()) == 0.0
    with torch.cuda.stream(signaller):
        ops.flag_signal([flags.data_ptr() + 4], 9)     # >= 7: releases the waiter
    waiter.synchronize()
    assert float(out.item()) == 1.0
    assert flags.cpu().tolist()[:3] == [0, 9, 6]


@pytest.mark.parametrize('world,rank', [(1, 0), (3, 1)])
def test_trsm_right_scatter(cuda, world, rank):
    """gabo_trsm_right_scatter_f64: same solve as gabo_trsm_right_winv_f64, and every destination receives the rows in
    global (block-cyclic) order."""
    from gaboflow_b200 import ops
    rng = np.random.default_rng(11 + world)
    nb, m = 512, 3 * 512 - 40                           # ragged last local panel
    L = np.tril(rng.standard_normal((nb, nb))) * 0.05 + np.eye(nb) * 2.0
    A = rng.standard_normal((m, nb))
    Ld = torch.from_numpy(L).cuda()
    Ad = torch.from_numpy(A).cuda()
    Wv = torch.empty((4, 128, 128), dtype=torch.float64, device='cuda')
    blk = (torch.from_numpy(L @ L.T).cuda()).contiguous()
    ops.potrf_diag_(blk, 0.0, Wv)                       # factor of L L^T = L; block inverses in Wv
    Aref = Ad.clone()
    ops.trsm_right_(Aref, torch.tril(blk), Wv)
    s0 = 2                                              # first local slot below the panel
    row_off = (s0 * world + rank) * nb - 1024           # panel kp = 1 -> k1 = 1024
    nrows = row_off + ((m - 1) // nb) * world * nb + (m - 1) % nb + 1
    dsts = [torch.full((nrows, nb), np.nan, dtype=torch.float64, device='cuda') for _ in range(2)]
    ops.trsm_right_scatter_(Ad, torch.tril(blk), Wv, [d.data_ptr() for d in dsts], nb, row_off, nb, world)
    assert torch.equal(Ad, Aref)
    i = np.arange(m)
    prow = torch.from_numpy(row_off + (i // nb) * world * nb + i % nb).cuda()
    for d in dsts:
        assert torch.equal(d[prow], Aref)
        untouched = torch.ones(nrows, dtype=torch.bool, device='cuda')
        untouched[prow] = False
        assert bool(torch.isnan(d[untouched]).all())


def test_p2p_driver_world1_matches_collective_driver(cuda):
    """potrf_block_cyclic_p2p_ (peer-memory exchange + look-ahead) gives the same bits as potrf_block_cyclic_."""
    from gaboflow_b200 import SphereGaussianKernel
    from gaboflow_b200.dist import BlockCyclic, gram_block_cyclic, potrf_block_cyclic_, potrf_block_cyclic_p2p_, PanelExchange
    from oracle import gabo_oracle as orc
    n = 2560
    Xd = torch.from_numpy(orc.synth_sphere(n, 51, seed=4)).cuda()
    k = SphereGaussianKernel(51, range(51), beta_min=0.0, beta=2.0)
    lay = BlockCyclic(n, 512, 1, 0)
    A = gram_block_cyclic(lambda r0, r1, out: k.K(Xd, uplo='lower', row0=r0, row1=r1, out=out), lay, device='cuda')
    B = A.clone()
    assert potrf_block_cyclic_(A, lay, 1e-2, n_factor=2048) == 0
    xch = PanelExchange(lay, 2048, 'cuda')
    try:
        for _ in range(2):                              # twice: flag epochs, buffer reuse
            C2 = B.clone()
            store = {}
            assert potrf_block_cyclic_p2p_(C2, lay, 1e-2, xch, n_factor=2048, winv_store=store) == 0
            assert torch.equal(torch.tril(C2[:2048, :2048]), torch.tril(A[:2048, :2048]))
            assert sorted(store) == [0, 1, 2, 3]
    finally:
        xch.close()


def test_p2p_driver_ragged_last_panel(cuda):
    """Peer-memory driver on a size that is not a multiple of the panel (last panel 252 rows), whole matrix factored."""
    from gaboflow_b200 import SphereGaussianKernel
    from gaboflow_b200.dist import BlockCyclic, gram_block_cyclic, potrf_block_cyclic_, potrf_block_cyclic_p2p_, PanelExchange
    from oracle import gabo_oracle as orc
    n = 4 * 512 + 252
    Xd = torch.from_numpy(orc.synth_sphere(n, 51, seed=14)).cuda()
    k = SphereGaussianKernel(51, range(51), beta_min=0.0, beta=2.0)
    lay = BlockCyclic(n, 512, 1, 0)
    A = gram_block_cyclic(lambda r0, r1, out: k.K(Xd, uplo='lower', row0=r0, row1=r1, out=out), lay, device='cuda')
    B = A.clone()
    assert potrf_block_cyclic_(A, lay, 1e-2) == 0
    xch = PanelExchange(lay, n, 'cuda')
    try:
        assert potrf_block_cyclic_p2p_(B, lay, 1e-2, xch) == 0
        assert torch.equal(torch.tril(B), torch.tril(A))
        ref = np.linalg.cholesky(orc.sphere_gaussian_K(Xd.cpu().numpy(), beta=2.0) + 1e-2 * np.eye(n))
        assert float(np.max(np.abs(torch.tril(B).cpu().numpy() - ref))) < 1e-12
    finally:
        xch.close()


def test_cyclic_staged_gram_fresh_state_on_recycled_memory(cuda):
    """A StagedGramState built on recycled device memory (stale flag words of an earlier state) while the stream still has a
    backlog: the copy streams must not evaluate the flags before the state's zero-initialisation has run (they are ordered
    after the caller's stream by an event since round 2; before that the copies could fire early and ship unwritten staging
    memory)."""
    from gaboflow_b200 import ops
    from gaboflow_b200.dist import BlockCyclic
    from oracle import gabo_oracle as orc
    world, frac, split, direct = 3, 48, 1, 128
    n = 512 * 11
    X = torch.from_numpy(orc.synth_sphere(n, 51, seed=8)).cuda()
    lays = [BlockCyclic(n, 512, world, r) for r in range(world)]
    bufs = [torch.full((max(l.local_rows, 1), n), float('nan'), dtype=torch.float64, device='cuda') for l in lays]
    ptrs = [b.data_ptr() for b in bufs]
    full = ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, uplo='mirror', engine='dmma')
    big = torch.from_numpy(orc.synth_sphere(8192, 51, seed=9)).cuda()
    scratch = torch.empty((8192, 8192), dtype=torch.float64, device='cuda')
    for attempt in range(4):
        # a state that leaves flag words = 2 behind, then goes back to the allocator
        states = [ops.StagedGramState(n, world, r, frac, split, 'cuda', direct_frac256=direct) for r in range(world)]
        for rep in range(2):
            for r in range(world):
                ops.points_gram_cyclic_staged(ops.SPHERE_GAUSSIAN, X, ptrs, n, world, r, states[r], beta=2.0, variance=1.0)
        torch.cuda.synchronize()
        del states
        for b in bufs:
            b.fill_(float('nan'))
        # backlog on the stream, then fresh states (same sizes: the caching allocator hands the same blocks back) used at once
        for _ in range(3):
            ops.points_gram(ops.SPHERE_GAUSSIAN, big, None, beta=2.0, out=scratch, uplo='mirror', engine='dmma')
        states = [ops.StagedGramState(n, world, r, frac, split, 'cuda', direct_frac256=direct) for r in range(world)]
        for r in range(world):
            ops.points_gram_cyclic_staged(ops.SPHERE_GAUSSIAN, X, ptrs, n, world, r, states[r], beta=2.0, variance=1.0)
        torch.cuda.synchronize()
        for r in range(world):
            rows = lays[r].global_rows().cuda()
            assert torch.equal(bufs[r][:lays[r].local_rows], full[rows]), (attempt, r)
