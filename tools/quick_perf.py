"""Developer timing harness (not the bench contract): times each kernel of the hot path at the BASELINE sizes.
Usage on the GPU box: python tools/quick_perf.py [sphere] [spd] [chol] [solve]  (default: all)"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from gaboflow_b200 import ops, SphereGaussianKernel, SpdAffineInvariantGaussianKernel  # noqa: E402
from oracle import gabo_oracle as orc  # noqa: E402

what = set(sys.argv[1:]) or {'sphere', 'spd', 'chol', 'solve'}
dev = torch.device('cuda:0')
res = {}


def timeit(f, reps=3, warm=1):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


if 'sphere' in what:
    for N in (16384, 65536):
        X = torch.from_numpy(orc.synth_sphere(N, 51)).to(dev)
        K = torch.empty((N, N), dtype=torch.float64, device=dev)
        ws = torch.empty(int(ops._lib.lib().gabo_points_gram_workspace_bytes(N, N, 51)), dtype=torch.uint8, device=dev)
        for uplo in ('full', 'mirror', 'lower'):
            for kind, nm in ((ops.SPHERE_GAUSSIAN, 'gauss'), (ops.SPHERE_LAPLACE, 'laplace')):
                if nm == 'laplace' and uplo != 'mirror':
                    continue
                best, med = timeit(lambda: ops.points_gram(kind, X, None, beta=2.0, out=K, uplo=uplo, workspace=ws))
                res[f'sphere_{nm}_{uplo}_N{N}_ms'] = best
                res[f'sphere_{nm}_{uplo}_N{N}_entries_per_s'] = N * N / best * 1e3
                print(f'sphere {nm} {uplo} N={N}: {best:.2f} ms  {N*N/best*1e3:.3e} entries/s  write {8*N*N/best*1e-6:.0f} GB/s', flush=True)
        if N == 65536:
            X32 = X.float(); K32 = torch.empty((N, N), dtype=torch.float32, device=dev)
            for uplo in ('full', 'mirror'):
                best, med = timeit(lambda: ops.points_gram(ops.SPHERE_GAUSSIAN, X32, None, beta=2.0, out=K32, uplo=uplo))
                res[f'sphere_gauss_f32_{uplo}_N{N}_ms'] = best
                print(f'sphere gauss FP32 {uplo} N={N}: {best:.2f} ms  {N*N/best*1e3:.3e} entries/s', flush=True)
            del K32
        del K, X

if 'spd' in what:
    N = 16384
    Xm = torch.from_numpy(orc.synth_spd_mandel(N, 6)).to(dev)
    K = torch.empty((N, N), dtype=torch.float64, device=dev)
    best, _ = timeit(lambda: ops.spd_prep(Xm, 6, want_logm=True))
    print(f'spd prep N={N}: {best:.3f} ms', flush=True)
    res['spd_prep_ms'] = best
    p = ops.spd_prep(Xm, 6, want_logm=True)
    for kind, nm in ((ops.SPD_AI_GAUSSIAN, 'ai_gauss'), (ops.SPD_STEIN, 'stein')):
        for uplo in ('full', 'mirror'):
            best, _ = timeit(lambda: ops.spd_gram(kind, p, None, beta=1.0 if nm != 'stein' else 3.0, out=K, uplo=uplo), reps=2)
            res[f'spd_{nm}_{uplo}_ms'] = best
            print(f'spd {nm} {uplo} N={N}: {best:.2f} ms  {N*N/best*1e3:.3e} entries/s', flush=True)
    for uplo in ('full', 'mirror'):
        best, _ = timeit(lambda: ops.points_gram(ops.SQDIST_GAUSSIAN, p.logm, None, beta=1.0, out=K, uplo=uplo))
        res[f'spd_logeuclid_{uplo}_ms'] = best
        print(f'spd logeuclid pair stage {uplo} N={N}: {best:.2f} ms  {N*N/best*1e3:.3e} entries/s', flush=True)
    del K

if 'chol' in what or 'solve' in what:
    for N in (8192, 16384, 32768):
        X = torch.from_numpy(orc.synth_sphere(N, 51)).to(dev)
        K = torch.empty((N, N), dtype=torch.float64, device=dev)
        kern = SphereGaussianKernel(51, range(51), beta_min=0.0, beta=2.0)
        kern.K(X, uplo='lower', out=K)
        A = K.clone()

        def run():
            A.copy_(K)
            ops.potrf_(A, 1e-2, check=False)
        tcopy, _ = timeit(lambda: A.copy_(K))
        best, med = timeit(run, reps=3)
        t = best - tcopy
        res[f'potrf_N{N}_s'] = t * 1e-3
        res[f'potrf_N{N}_tflops'] = N ** 3 / 3 / t * 1e-9
        print(f'potrf N={N}: {t:.1f} ms  {N**3/3/t*1e-9:.2f} TFLOP/s  (copy {tcopy:.1f} ms excluded)', flush=True)
        if 'solve' in what:
            y = torch.randn((N, 1), dtype=torch.float64, device=dev)
            v = y.clone()
            def solve1():
                v.copy_(y); ops.trsm_(A, v)
            best, _ = timeit(solve1)
            res[f'trsv_N{N}_ms'] = best
            print(f'trsv N={N}: {best:.2f} ms   ({4*N*N/best*1e-6:.0f} GB/s of L)', flush=True)
            B0 = torch.randn((N, 1024), dtype=torch.float64, device=dev); B = B0.clone()
            def solveM():
                B.copy_(B0); ops.trsm_(A, B)
            best, _ = timeit(solveM)
            res[f'trsm1024_N{N}_ms'] = best
            print(f'trsm nrhs=1024 N={N}: {best:.2f} ms  {N*N*1024/best*1e-9:.2f} TFLOP/s', flush=True)
            del B0, B
        # correctness at scale: residual of the factor on a sampled block
        Lt = torch.tril(A[:2048, :2048])
        Kb = torch.tril(K[:2048, :2048]); Kb = Kb + torch.tril(Kb, -1).t(); Kb.diagonal().add_(1e-2)
        r = (Lt @ Lt.t() - Kb).abs().max().item()
        print(f'  leading 2048 block residual max|LL^T - K| = {r:.2e}', flush=True)
        res[f'potrf_N{N}_resid'] = r
        del A, K, X
print(json.dumps(res))
