"""Developer A/B harness for the SPD affine-invariant pair kernel (not the bench contract): parity of a sampled block
against the oracle and the time of the full N = 16384, d = 6 Gram; GABO_SPD_AI_V1=1 selects the first-generation kernel.
usage: python tools/spd_ab.py [N]"""
import os
import subprocess
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from gaboflow_b200 import ops  # noqa: E402
from oracle import gabo_oracle as orc  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
tag = 'v1' if os.environ.get('GABO_SPD_AI_V1') else 'v2'
dev = torch.device('cuda:0')
Xm_h = orc.synth_spd_mandel(N, 6)
Xm = torch.from_numpy(Xm_h).to(dev)
p = ops.spd_prep(Xm, 6)
K = torch.empty((N, N), dtype=torch.float64, device=dev)


def timeit(f, reps=3):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for uplo in ('mirror', 'lower', 'full'):
    t = timeit(lambda: ops.spd_gram(ops.SPD_AI_GAUSSIAN, p, None, beta=1.0, out=K, uplo=uplo))
    pairs = N * (N + 1) / 2 if uplo != 'full' else N * N
    print(f'[{tag}] AI gauss {uplo} N={N}: {t:.2f} ms  {pairs / t * 1e-6:.2f} Gpairs/s  '
          f'declared-roofline frac {2532 * pairs / (t * 1e-3) / 37.2e12:.3f}', flush=True)
ops.spd_gram(ops.SPD_AI_GAUSSIAN, p, None, beta=1.0, out=K, uplo='mirror')
rows = np.r_[0:64, N // 2:N // 2 + 64, N - 64:N]
ref = orc.spd_ai_gaussian_K(Xm_h[rows], Xm_h, beta=1.0)
got = K[torch.from_numpy(rows).to(dev)].cpu().numpy()
mask = np.ones_like(ref, dtype=bool)
mask[np.arange(len(rows)), rows] = False
print(f'[{tag}] max rel err vs oracle on {len(rows)} rows: {np.max(np.abs(got - ref)[mask] / ref[mask]):.3e}; '
      f'symmetric {bool(torch.equal(K, K.t()))}', flush=True)
if tag == 'v2' and '--no-v1' not in sys.argv:
    env = dict(os.environ, GABO_SPD_AI_V1='1')
    subprocess.run([sys.executable, __file__, str(N)], env=env, check=False)
