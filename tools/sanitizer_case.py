"""Small shapes of every shared-memory / mbarrier / peer-store kernel for compute-sanitizer (racecheck, memcheck, synccheck):
    compute-sanitizer --tool racecheck python tools/sanitizer_case.py
The shapes are ragged on purpose (edges of the TMA boxes, partially filled tiles)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from gaboflow_b200 import ops  # noqa: E402
from oracle import gabo_oracle as orc  # noqa: E402

dev = 'cuda'
X = torch.from_numpy(orc.synth_sphere(300, 51)).to(dev)
for uplo in ('full', 'mirror', 'lower'):
    K = ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, uplo=uplo)
K32 = ops.points_gram(ops.SPHERE_GAUSSIAN, X.float(), None, beta=2.0, uplo='mirror')
X3 = torch.from_numpy(orc.synth_sphere(200, 3)).to(dev)
ops.points_gram(ops.SPHERE_LAPLACE, X3, X3[:77].contiguous(), beta=1.0)
Xm = torch.from_numpy(orc.synth_spd_mandel(200, 6)).to(dev)
p = ops.spd_prep(Xm, 6, want_logm=True)
for kind in (ops.SPD_AI_GAUSSIAN, ops.SPD_AI_LAPLACE, ops.SPD_STEIN):
    ops.spd_gram(kind, p, None, beta=1.0, uplo='mirror')
Xm3 = torch.from_numpy(orc.synth_spd_mandel(90, 3)).to(dev)
p3 = ops.spd_prep(Xm3, 3)
ops.spd_gram(ops.SPD_AI_GAUSSIAN, p3, None, beta=1.0, uplo='full')
A = ops.points_gram(ops.SPHERE_GAUSSIAN, torch.from_numpy(orc.synth_sphere(700, 51)).to(dev), None, beta=2.0, uplo='lower')
ops.potrf_(A, 1e-2)
b = torch.randn((700, 1), dtype=torch.float64, device=dev)
ops.trsm_(A, b)
ops.trsm_(A, b, trans=True)
B = torch.randn((700, 33), dtype=torch.float64, device=dev)
ops.trsm_(A, B)
print('loglik', float(ops.gp_loglik(A, b).item()))
torch.cuda.synchronize()
print('done')
