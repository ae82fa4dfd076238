import sys, numpy as np
sys.path.insert(0, '.')
from oracle import gabo_oracle as orc
from gaboflow_b200 import SpdAffineInvariantGaussianKernel
for d in (3, 6, 8):
    rng = np.random.default_rng(700 + d)
    m = d * (d + 1) // 2
    base = np.linspace(0.6, 1.9, d)
    spectra = [np.full(d, 1.7), np.full(d, 1.0), base, base * 3.0,
               1.0 + 1e-9 * np.arange(d), 1.3 + 1e-6 * np.arange(d), 0.8 + 1e-4 * np.arange(d), 1.1 + 1e-2 * np.arange(d),
               np.r_[np.full(d - 1, 2.0), 0.5], np.r_[np.full(d // 2, 0.7), np.full(d - d // 2, 1.4)],
               np.r_[base[:-2], base[-3] * (1 + 3e-8), base[-3] * (1 - 2e-8)],
               np.geomspace(1e-3, 30.0, d), np.geomspace(0.01, 1.0, d), np.geomspace(1.0, 45.0, d)]
    n1 = 40
    A = rng.standard_normal((n1, d, d))
    S = A @ A.transpose(0, 2, 1) + d * np.eye(d)
    L = np.linalg.cholesky(S)
    S2, d2 = [], []
    for lam in spectra:
        for _ in range(3):
            Q, _ = np.linalg.qr(rng.standard_normal((d, d)))
            S2.append(Q @ np.diag(lam) @ Q.T); d2.append(np.sum(np.log(lam) ** 2))
    S2 = np.array(S2); d2 = np.array(d2)
    X = orc.sym_to_mandel(S)
    k = SpdAffineInvariantGaussianKernel(m, range(m), beta_min=0.0, beta=0.3, variance=1.0)
    eo = np.zeros(len(spectra)); ee = np.zeros(len(spectra)); oe = np.zeros(len(spectra))
    for i in range(0, n1, 8):
        X2 = orc.sym_to_mandel(np.einsum('ab,jbc,dc->jad', L[i], S2, L[i]))
        Krow = k.compute_K(X[i:i + 1], X2)[0]
        Kor = orc.spd_ai_gaussian_K(X[i:i + 1], X2, beta=0.3)[0]
        ref = np.exp(-0.3 * d2)
        eo = np.maximum(eo, (np.abs(Krow - Kor) / Kor).reshape(-1, 3).max(1))
        ee = np.maximum(ee, (np.abs(Krow - ref) / ref).reshape(-1, 3).max(1))
        oe = np.maximum(oe, (np.abs(Kor - ref) / ref).reshape(-1, 3).max(1))
    print('d', d)
    for j in range(len(spectra)):
        print(f'  spectrum {j:2d}: gpu-vs-oracle {eo[j]:.1e}  gpu-vs-exact {ee[j]:.1e}  oracle-vs-exact {oe[j]:.1e}')
