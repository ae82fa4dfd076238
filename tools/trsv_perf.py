"""Developer timing of the one-right-hand-side triangular solve (persistent kernel vs GABO_TRSV_CHAIN=1 launch chain)."""
import os, sys, subprocess
import numpy as np, torch
sys.path.insert(0, '.')
from gaboflow_b200 import ops
from oracle import gabo_oracle as orc
tag = 'chain' if os.environ.get('GABO_TRSV_CHAIN') else 'persistent'
for N in (4096, 32768):
    X = torch.from_numpy(orc.synth_sphere(N, 51)).cuda()
    A = torch.empty((N, N), dtype=torch.float64, device='cuda')
    ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, out=A, uplo='lower')
    ops.potrf_(A, 1e-2)
    y = torch.randn((N, 1), dtype=torch.float64, device='cuda')
    for trans in (False, True):
        v = y.clone()
        def f():
            v.copy_(y); ops.trsm_(A, v, trans=trans)
        f(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        Lt = torch.tril(A)
        r = ((Lt.t() if trans else Lt) @ v - y).abs().max().item() if N <= 8192 else float('nan')
        print(f'[{tag}] N={N} trans={trans}: {min(ts):.3f} ms  ({4*N*N/min(ts)*1e-6:.0f} GB/s of L)  resid {r:.2e}', flush=True)
    del A, X
if tag == 'persistent':
    subprocess.run([sys.executable, __file__], env=dict(os.environ, GABO_TRSV_CHAIN='1'))
