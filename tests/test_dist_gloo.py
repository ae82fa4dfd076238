"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (block-cyclic ownership, panel broadcast,
block-column all-gather and its reordering, masks in global coordinates, pipelined substitution, additive log-likelihood).
The local arithmetic is replaced by a torch-CPU stand-in that lives only in this test (the product has CUDA only)."""
import math
import os

import numpy as np
import pytest
import scipy.linalg as sla
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaboflow_b200.dist import (BlockCyclic, gram_block_cyclic, potrf_block_cyclic_, solve_lower_block_cyclic_,
                                gp_loglik_block_cyclic)


class TorchCpuLocalOps:
    @staticmethod
    def potrf_(A, sigma2):
        n = A.shape[0]
        M = torch.tril(A) + torch.tril(A, -1).T + sigma2 * torch.eye(n, dtype=A.dtype)
        L, info = torch.linalg.cholesky_ex(M)
        A.copy_(torch.tril(L) + torch.triu(A, 1))
        return info.to(torch.int32).reshape(1)

    @staticmethod
    def potrf_diag_(A, sigma2, winv_all):
        return TorchCpuLocalOps.potrf_(A, sigma2)

    @staticmethod
    def trsm_right_(A, L, winv_all=None):
        A.copy_(torch.linalg.solve_triangular(torch.tril(L), A.T.contiguous(), upper=False).T)
        return A

    @staticmethod
    def syrk_cyclic_(C, A, B, row_base, col_base, nb, world, rank):
        m, n = C.shape
        r = torch.arange(m)
        grow = row_base + ((r // nb) * world + rank) * nb + (r % nb) if nb > 0 else row_base + r
        gcol = col_base + torch.arange(n)
        mask = gcol[None, :] <= grow[:, None]
        upd = A @ B.T
        C.sub_(torch.where(mask, upd, torch.zeros_like(upd)))
        return C

    @staticmethod
    def trsm_left_(L, B, winv_all=None):
        B.copy_(torch.linalg.solve_triangular(torch.tril(L), B, upper=False))
        return B

    @staticmethod
    def solve_update_(Y, Lp, X):
        Y.sub_(Lp @ X)
        return Y

    @staticmethod
    def loglik_block(L, V):
        n, p = V.shape
        return torch.tensor([-0.5 * n * p * math.log(2 * math.pi) - p * torch.log(torch.diagonal(L)).sum() - 0.5 * (V ** 2).sum()],
                            dtype=torch.float64)


def test_block_cyclic_layout():
    lay = BlockCyclic(n=2304, nb=512, world=2, rank=1)
    assert lay.npanels == 5 and lay.my_panels() == [1, 3] and lay.local_rows == 1024
    assert BlockCyclic(2304, 512, 2, 0).local_rows == 512 + 512 + 256
    assert lay.slots_after(0) == 0 and lay.slots_after(1) == 1 and lay.slots_after(2) == 1 and lay.slots_after(3) == 2
    rows = torch.cat([BlockCyclic(2304, 512, 2, r).global_rows() for r in range(2)])
    assert sorted(rows.tolist()) == list(range(2304))


def _worker(rank, world, port, n, nb, nfac, tmp):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        G = rng.standard_normal((n, n // 2))
        K = G @ G.T / (n // 2)
        y = rng.standard_normal((n, 2))
        lay = BlockCyclic(n, nb, world, rank)
        Kt = torch.from_numpy(K)

        def gram_rows(r0, r1, out):      # lower part only is meaningful; poison the strictly upper part
            blk = Kt[r0:r1].clone()
            cols = torch.arange(n)[None, :]
            rows = torch.arange(r0, r1)[:, None]
            blk[cols > rows] = float('nan')
            out.copy_(blk)

        A = gram_block_cyclic(gram_rows, lay, dtype=torch.float64, device='cpu')
        info = potrf_block_cyclic_(A, lay, 0.25, n_factor=nfac, local=TorchCpuLocalOps)
        assert info == 0
        nf = nfac or n
        Lref = sla.cholesky(K[:nf, :nf] + 0.25 * np.eye(nf), lower=True)
        grow = lay.global_rows()
        sel = grow < nf
        got = A[sel][:, :nf].numpy()
        ref = Lref[grow[sel].numpy()]
        mask = np.arange(nf)[None, :] <= grow[sel].numpy()[:, None]
        assert np.max(np.abs(np.where(mask, got - ref, 0.0))) < 1e-12
        V = torch.from_numpy(y[grow[sel].numpy()]).clone()
        solve_lower_block_cyclic_(A, V, lay, n_factor=nfac, local=TorchCpuLocalOps)
        vref = sla.solve_triangular(Lref, y[:nf], lower=True)
        assert np.max(np.abs(V.numpy() - vref[grow[sel].numpy()])) < 1e-11
        ll = gp_loglik_block_cyclic(A, V, lay, n_factor=nfac, local=TorchCpuLocalOps)
        llref = -0.5 * nf * 2 * math.log(2 * math.pi) - 2 * np.log(np.diag(Lref)).sum() - 0.5 * np.sum(vref ** 2)
        assert abs(ll - llref) < 1e-9 * abs(llref)
        # a non-PD matrix is reported with the same global index on every rank
        Kbad = K.copy(); Kbad[300, 300] = -5.0
        Ktb = torch.from_numpy(Kbad)
        Ab = gram_block_cyclic(lambda r0, r1, out: out.copy_(Ktb[r0:r1]), lay, device='cpu')
        assert potrf_block_cyclic_(Ab, lay, 0.0, local=TorchCpuLocalOps) == 301
        open(os.path.join(tmp, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n,nb,nfac', [(1152, 128, None), (1024, 256, 512), (640, 128, None)])
def test_distributed_cholesky_world2_gloo(tmp_path, n, nb, nfac):
    port = 29500 + (os.getpid() + n + nb) % 2000
    mp.spawn(_worker, args=(2, port, n, nb, nfac, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / 'ok0').exists() and (tmp_path / 'ok1').exists()
