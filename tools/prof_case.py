"""Small, fixed case for ncu captures: one sphere Gram (mirror) + one Cholesky + solve, optionally one SPD AI Gram.
usage: python tools/prof_case.py [N_gram] [N_chol] [N_spd]"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from gaboflow_b200 import ops  # noqa: E402

Ng = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
Nc = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
Ns = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rng = np.random.default_rng(20240)
X = rng.standard_normal((Ng, 51)); X /= np.linalg.norm(X, axis=1, keepdims=True)
Xd = torch.from_numpy(X).cuda()
K = torch.empty((Ng, Ng), dtype=torch.float64, device='cuda')
for _ in range(2):
    ops.points_gram(ops.SPHERE_GAUSSIAN, Xd, None, beta=2.0, out=K, uplo='mirror')
torch.cuda.synchronize()
if Nc > 0:
    A = K[:Nc, :Nc]
    ops.potrf_(A, 1e-2, check=False)
    v = torch.randn((Nc, 1), dtype=torch.float64, device='cuda')
    ops.trsm_(A, v)
    print('loglik', float(ops.gp_loglik(A, v).item()))
if Ns > 0:
    G = rng.standard_normal((Ns, 6, 6)); Q, _ = np.linalg.qr(G); w = 0.5 + 1.5 * rng.random((Ns, 6))
    S = np.einsum('nij,nj,nkj->nik', Q, w, Q)
    r, c = np.triu_indices(6)
    from oracle.gabo_oracle import sym_to_mandel
    Xm = torch.from_numpy(sym_to_mandel(S)).cuda()
    p = ops.spd_prep(Xm, 6)
    Ks = ops.spd_gram(ops.SPD_AI_GAUSSIAN, p, None, beta=1.0, uplo='mirror')
    print('spd sum', float(Ks.sum().item()))
torch.cuda.synchronize()
print('done')
