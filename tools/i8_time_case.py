"""Timing of the int8 Gram kernel under its diagnostic knobs (GABO_I8_DEBUG bit mask, see points_gram_i8.cu).
    python tools/i8_time_case.py [N] [mask ...]"""
import os, sys, torch
sys.path.insert(0, '.')
from gaboflow_b200 import ops
from oracle import gabo_oracle as orc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
masks = [int(a) for a in sys.argv[2:]] or [0]
X = torch.from_numpy(orc.synth_sphere(N, 51, seed=20240)).cuda()
K = torch.empty((N, N), dtype=torch.float64, device='cuda')
for m in masks:
    os.environ['GABO_I8_DEBUG'] = str(m)
    for uplo in ('lower',):
        for _ in range(1 if m & 16 else 2):
            ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, out=K, uplo=uplo, engine='i8')
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        reps = 1 if m & 16 else 5
        for _ in range(reps):
            ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, out=K, uplo=uplo, engine='i8')
        e1.record(); torch.cuda.synchronize()
        print(f'N={N} i8 {uplo} debug={m}: {e0.elapsed_time(e1) / reps:.3f} ms', flush=True)
