"""One int8-engine Gram for ncu: python tools/i8_prof_case.py [N] [uplo]"""
import sys
import torch
sys.path.insert(0, '.')
from gaboflow_b200 import ops
from oracle import gabo_oracle as orc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
uplo = sys.argv[2] if len(sys.argv) > 2 else 'mirror'
X = torch.from_numpy(orc.synth_sphere(N, 51, seed=20240)).cuda()
K = torch.empty((N, N), dtype=torch.float64, device='cuda')
for _ in range(2):
    ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, out=K, uplo=uplo, engine='i8')
torch.cuda.synchronize()
