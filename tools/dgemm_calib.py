"""Calibrator only (never on the product path): cuBLAS FP64 GEMM / SYRK-shaped / POTRF rates on this B200.
Gives the measured FP64 'tensor' peak that roofline fractions are quoted against."""
import json, sys, time, torch
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda:0')
out = {}
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for n in (4096, 8192):
    a = torch.randn(n, n, device=dev, dtype=torch.float64); b = torch.randn(n, n, device=dev, dtype=torch.float64)
    c = torch.empty_like(a)
    ms = t(lambda: torch.matmul(a, b, out=c))
    out[f'dgemm_{n}_tflops'] = 2 * n**3 / ms * 1e-9
    ms = t(lambda: torch.matmul(a, b.t(), out=c))
    out[f'dgemm_nt_{n}_tflops'] = 2 * n**3 / ms * 1e-9
# sustained 3 s
n = 8192
a = torch.randn(n, n, device=dev, dtype=torch.float64); b = torch.randn(n, n, device=dev, dtype=torch.float64); c = torch.empty_like(a)
torch.cuda.synchronize(); t0 = time.time(); k = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
while time.time() - t0 < 3.0:
    for _ in range(5): torch.matmul(a, b, out=c)
    k += 5; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
out['dgemm_8192_sustained_tflops'] = k * 2 * n**3 / e0.elapsed_time(e1) * 1e-9
# skinny-K (panel update shape): C[n,n] -= A[n,k] A[n,k]^T
for k_ in (128, 256, 512):
    n = 16384
    A = torch.randn(n, k_, device=dev, dtype=torch.float64); C = torch.zeros(n, n, device=dev, dtype=torch.float64)
    ms = t(lambda: torch.addmm(C, A, A.t(), beta=1.0, alpha=-1.0, out=C))
    out[f'dgemm_rank{k_}_update_16384_tflops'] = 2 * n * n * k_ / ms * 1e-9
del A, C
for n in (8192, 16384, 32768):
    X = torch.randn(n, 64, device=dev, dtype=torch.float64)
    K = X @ X.t(); K.diagonal().add_(n)
    ms = t(lambda: torch.linalg.cholesky(K), reps=2)
    out[f'cusolver_potrf_{n}_s'] = ms * 1e-3
    out[f'cusolver_potrf_{n}_tflops'] = n**3 / 3 / ms * 1e-9
    del K, X
# copy bandwidth f64
x = torch.empty(1 << 29, device=dev, dtype=torch.float64); y = torch.empty_like(x)
ms = t(lambda: y.copy_(x)); out['copy_gbs'] = 2 * x.numel() * 8 / ms * 1e-6
ms = t(lambda: y.fill_(1.0)); out['fill_gbs'] = x.numel() * 8 / ms * 1e-6
print(json.dumps(out, indent=1))
