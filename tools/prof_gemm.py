"""One standalone launch of the dominant kernel at the bench shape: trailing update C -= A A^T (lower), m = 32256, k = 512.
usage: python tools/prof_gemm.py [m] [k]"""
import sys
import torch
sys.path.insert(0, '.')
from gaboflow_b200 import ops  # noqa: E402
m = int(sys.argv[1]) if len(sys.argv) > 1 else 32256
k = int(sys.argv[2]) if len(sys.argv) > 2 else 512
A = torch.randn((m, k), dtype=torch.float64, device='cuda')
C = torch.zeros((m, m), dtype=torch.float64, device='cuda')
for _ in range(2):
    ops.gemm_nt_sub_(C, A, A, lower=True)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); ops.gemm_nt_sub_(C, A, A, lower=True); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
tiles = sum(min((bi * 128 + 127) // 64 + 1, (m + 63) // 64) for bi in range((m + 127) // 128))
fl = tiles * 128 * 64 * k * 2
print(f'm={m} k={k}: {ms:.3f} ms, {tiles} tiles, {fl/ms*1e-9:.2f} TFLOP/s')
