"""Int8-tensor-core FP64 Gram (gabo_points_gram_i8_f64) against the DMMA kernel and the oracle: accuracy + timing.
    python tools/i8_gram_check.py [--big]"""
import sys
import time
import numpy as np
import torch

sys.path.insert(0, '.')
from gaboflow_b200 import ops  # noqa: E402
from oracle import gabo_oracle as orc  # noqa: E402

big = '--big' in sys.argv
ok = True
for (n, D, kind, uplo) in [(300, 3, ops.SPHERE_GAUSSIAN, 'full'), (1000, 51, ops.SPHERE_GAUSSIAN, 'mirror'), (1000, 51, ops.SPHERE_LAPLACE, 'lower'),
                           (4096, 51, ops.SPHERE_GAUSSIAN, 'mirror'), (2500, 64, ops.SPHERE_GAUSSIAN, 'full'), (777, 17, ops.SPHERE_LAPLACE, 'mirror')]:
    X = torch.from_numpy(orc.synth_sphere(n, D, seed=n + D)).cuda()
    ref = ops.points_gram(kind, X, None, beta=2.0, variance=1.3, uplo=uplo, engine='dmma', out=torch.full((n, n), -7.0, dtype=torch.float64, device='cuda'))
    got = ops.points_gram(kind, X, None, beta=2.0, variance=1.3, uplo=uplo, engine='i8', out=torch.full((n, n), -7.0, dtype=torch.float64, device='cuda'))
    torch.cuda.synchronize()
    rel = float(((got - ref).abs() / ref.abs().clamp_min(1e-300)).max().item())
    same_untouched = bool(((got == -7.0) == (ref == -7.0)).all().item())
    Ko = orc.sphere_gaussian_K(X.cpu().numpy(), beta=2.0, variance=1.3) if kind == ops.SPHERE_GAUSSIAN else orc.sphere_laplace_K(X.cpu().numpy(), beta=2.0, variance=1.3)
    g = got.cpu().numpy()
    mask = (g != -7.0) & ~np.eye(n, dtype=bool)
    relo = float(np.max(np.abs(g - Ko)[mask] / Ko[mask]))
    sym = float((got - got.t()).abs().max().item()) if uplo == 'mirror' else 0.0
    good = rel < 1e-12 and relo < 1e-12 and same_untouched and sym == 0.0
    ok = ok and good
    print(f'n={n} D={D} kind={kind} uplo={uplo}: max rel diff vs DMMA {rel:.2e}, vs oracle {relo:.2e}, untouched pattern equal {same_untouched}, asym {sym:.1e}  {"OK" if good else "FAIL"}', flush=True)
# cross Gram + row range
X = torch.from_numpy(orc.synth_sphere(1500, 51, seed=1)).cuda()
X2 = torch.from_numpy(orc.synth_sphere(900, 51, seed=2)).cuda()
ref = ops.points_gram(ops.SPHERE_GAUSSIAN, X, X2, beta=2.0, engine='dmma', row0=256, row1=1300)
got = ops.points_gram(ops.SPHERE_GAUSSIAN, X, X2, beta=2.0, engine='i8', row0=256, row1=1300)
rel = float(((got - ref).abs() / ref.abs()).max().item())
print(f'cross Gram rows 256:1300 x 900: max rel diff {rel:.2e}  {"OK" if rel < 1e-12 else "FAIL"}', flush=True)
ok = ok and rel < 1e-12
# cyclic full rows, 3 emulated ranks
n = 512 * 7
X = torch.from_numpy(orc.synth_sphere(n, 51, seed=8)).cuda()
full = ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, uplo='full', engine='i8')
from gaboflow_b200.dist import BlockCyclic  # noqa: E402
for W in (2, 3):
    for r in range(W):
        lay = BlockCyclic(n, 512, W, r)
        buf = torch.full((lay.local_rows, n), float('nan'), dtype=torch.float64, device='cuda')
        ops.points_gram_cyclic_rows_i8(ops.SPHERE_GAUSSIAN, X, buf, W, r, beta=2.0)
        eq = bool(torch.equal(buf, full[lay.global_rows().cuda()]))
        ok = ok and eq
        print(f'cyclic rows W={W} rank={r}: equal to the single-GPU rows {eq}', flush=True)
for N in ([16384, 65536] if big else [16384]):
    X = torch.from_numpy(orc.synth_sphere(N, 51, seed=20240)).cuda()
    K = torch.empty((N, N), dtype=torch.float64, device='cuda')
    for eng in ('dmma', 'i8'):
        for uplo in ('mirror', 'lower', 'full'):
            ws = None
            for _ in range(2):
                ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, out=K, uplo=uplo, engine=eng)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(3):
                ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, out=K, uplo=uplo, engine=eng)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f'N={N} {eng} {uplo}: {ms:.3f} ms  {N * N / ms * 1e-6:.1f} G entries/s', flush=True)
sys.exit(0 if ok else 1)
