"""Multi-GPU correctness check, run under torchrun (one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py [N]
Distributed Gram (block-cyclic row panels) + Cholesky + solve + loglik against the single-GPU path on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, '.')
from gaboflow_b200 import SphereGaussianKernel, ops  # noqa: E402
from gaboflow_b200.dist import (BlockCyclic, gram_block_cyclic, potrf_block_cyclic_, solve_lower_block_cyclic_,  # noqa: E402
                                gp_loglik_block_cyclic, PanelExchange, potrf_block_cyclic_p2p_, solve_lower_block_cyclic_p2p_)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
rng = np.random.default_rng(20240)
X = rng.standard_normal((N, 51)); X /= np.linalg.norm(X, axis=1, keepdims=True)
y = np.random.default_rng(20241).standard_normal((N, 1))
Xd = torch.from_numpy(X).to(dev)
k = SphereGaussianKernel(51, range(51), beta_min=0.0, beta=2.0)
lay = BlockCyclic(N, 512, world, rank)
A = gram_block_cyclic(lambda r0, r1, out: k.K(Xd, uplo='lower', row0=r0, row1=r1, out=out), lay, device=dev)
A2 = A.clone()
info = potrf_block_cyclic_(A, lay, 1e-2)
# same factorisation with the panel exchange through peer memory + look-ahead: must agree bit for bit (same kernels, same
# order of operations per entry); run twice to exercise the flag epochs and the buffer reuse
xch = PanelExchange(lay, N, dev)
A3 = A2.clone()
info2 = potrf_block_cyclic_p2p_(A2, lay, 1e-2, xch)
store = {}
info3 = potrf_block_cyclic_p2p_(A3, lay, 1e-2, xch, winv_store=store)
gmask = torch.arange(N, device=dev)[None, :] <= lay.global_rows().to(dev)[:, None]
p2p_err = max(float(torch.where(gmask, (A2 - A).abs(), torch.zeros_like(A)).max().item()),
              float(torch.where(gmask, (A3 - A).abs(), torch.zeros_like(A)).max().item()))
grow = lay.global_rows()
V = torch.from_numpy(y[grow.numpy()]).to(dev)
solve_lower_block_cyclic_(A, V, lay)
ll = gp_loglik_block_cyclic(A, V, lay)
# the substitution through peer memory (push + flag per block) must reproduce the collective one; twice for the epochs
for _ in range(2):
    V2 = torch.from_numpy(y[grow.numpy()]).to(dev)
    solve_lower_block_cyclic_p2p_(A, V2, lay, xch)
    ll2 = gp_loglik_block_cyclic(A, V2, lay)
    p2p_err = max(p2p_err, float((V2 - V).abs().max().item()))     # same kernels, same order per entry: bit-identical
# ... and as one C call per rank, with the block inverses the factorisation left behind (differs by rounding only)
V3 = torch.from_numpy(y[grow.numpy()]).to(dev)
solve_lower_block_cyclic_p2p_(A3, V3, lay, xch, winv_store=store)
c_err = float((V3 - V).abs().max().item())
if c_err > 1e-11:
    p2p_err = max(p2p_err, c_err)
xch.close()
# single-GPU reference on every rank (cheap at this size)
Kf = torch.empty((N, N), dtype=torch.float64, device=dev)
k.K(Xd, uplo='lower', out=Kf)
ops.potrf_(Kf, 1e-2)
v1 = torch.from_numpy(y).to(dev)
ops.trsm_(Kf, v1)
ll1 = float(ops.gp_loglik(Kf, v1).item())
Lref = torch.tril(Kf)[grow.to(dev)]
mask = torch.arange(N, device=dev)[None, :] <= grow.to(dev)[:, None]
err = float(torch.where(mask, (A - Lref).abs(), torch.zeros_like(A)).max().item())
verr = float((V - v1[grow.to(dev)]).abs().max().item())
ok = info == 0 and info2 == 0 and info3 == 0 and p2p_err == 0.0 and err < 1e-10 and verr < 1e-9 and abs(ll / ll1 - 1) < 1e-10
print(f'[rank {rank}/{world}] info={info} p2p-vs-collective max diff={p2p_err:.1e} max|L-Lref|={err:.2e} max|v-vref|={verr:.2e} loglik {ll:.6f} vs {ll1:.6f}  {"OK" if ok else "FAIL"}', flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
