"""Fixed case for ncu captures of the round-2 kernels: FP32 tensor-core sphere Gram (tcgen05), one-kernel TRSV, fused
acquisition.  usage: python tools/prof_case2.py [N]"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from gaboflow_b200 import ops, GPR, SphereGaussianKernel, ExpectedImprovement  # noqa: E402
from oracle import gabo_oracle as orc  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
X = torch.from_numpy(orc.synth_sphere(N, 51)).cuda()
X32 = X.float()
K32 = torch.empty((N, N), dtype=torch.float32, device='cuda')
for _ in range(2):
    ops.points_gram(ops.SPHERE_GAUSSIAN, X32, None, beta=2.0, out=K32, uplo='mirror')
torch.cuda.synchronize()
del K32
n = 8192
K = torch.empty((n, n), dtype=torch.float64, device='cuda')
ops.points_gram(ops.SPHERE_GAUSSIAN, X[:n], None, beta=2.0, out=K, uplo='lower')
ops.potrf_(K, 1e-2, check=False)
v = torch.randn((n, 1), dtype=torch.float64, device='cuda')
for _ in range(2):
    ops.trsm_(K, v)
torch.cuda.synchronize()
print('done')
