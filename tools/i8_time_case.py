import sys, torch
sys.path.insert(0, '.')
from gaboflow_b200 import ops
from oracle import gabo_oracle as orc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
X = torch.from_numpy(orc.synth_sphere(N, 51, seed=20240)).cuda()
K = torch.empty((N, N), dtype=torch.float64, device='cuda')
for uplo in ('lower', 'mirror'):
    for _ in range(2):
        ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, out=K, uplo=uplo, engine='i8')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, out=K, uplo=uplo, engine='i8')
    e1.record(); torch.cuda.synchronize()
    print(f'N={N} i8 {uplo}: {e0.elapsed_time(e1) / 5:.3f} ms', flush=True)
