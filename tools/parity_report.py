"""Measured parity of every Gram kernel on the golden fixtures: max relative error of the CUDA result against (a) the
golden matrices generated from the reference's NumPy code, (b) the 60-digit mpmath values of the defining formulas
(tests/golden/spd_truth_mp.npz), (c) the NumPy oracle.  Run on the GPU box; writes one JSON object (commit it under
profiles/ so that every tolerance in tests/test_gpu_gram.py is backed by a number).
    python tools/parity_report.py > gpurun_out/r02_parity_errors.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gaboflow_b200 import (SphereGaussianKernel, SphereLaplaceKernel, SpdAffineInvariantGaussianKernel,   # noqa: E402
                           SpdAffineInvariantLaplaceKernel, SpdSteinGaussianKernel, SpdFrobeniusGaussianKernel,
                           SpdLogEuclideanGaussianKernel)
from oracle import gabo_oracle as orc   # noqa: E402

G = os.path.join(ROOT, 'tests', 'golden')


def rel(a, b, mask=None):
    e = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    if mask is not None:
        e = e[mask]
    return float(np.max(e))


def main():
    out = {}
    g = dict(np.load(os.path.join(G, 'sphere_s2.npz')))
    k = SphereGaussianKernel(3, range(3), beta_min=7.0, beta=10.0)
    kl = SphereLaplaceKernel(3, range(3), beta=10.0)
    off = ~np.eye(20, dtype=bool)
    out['sphere_s2'] = {
        'gauss11_offdiag_vs_golden': rel(k.compute_K_symm(g['x_man']), g['Kgauss11'], off),
        'gauss12_vs_golden': rel(k.compute_K(g['x_man'], g['x_test']), g['Kgauss12']),
        'laplace11_offdiag_vs_golden': rel(kl.compute_K_symm(g['x_man']), g['Klaplace11'], off),
        'laplace12_vs_golden': rel(kl.compute_K(g['x_man'], g['x_test']), g['Klaplace12']),
    }
    t = dict(np.load(os.path.join(G, 'spd_truth_mp.npz')))
    g = dict(np.load(os.path.join(G, 'spd_s2pp.npz')))
    x, xt = g['x_man'], g['x_test']
    off = ~np.eye(x.shape[0], dtype=bool)
    kerns = {
        'Kai': (SpdAffineInvariantGaussianKernel(3, range(3), beta=1.0, variance=1., beta_min=0.6), lambda a, b=None: orc.spd_ai_gaussian_K(a, b, beta=1.6)),
        'Kailap': (SpdAffineInvariantLaplaceKernel(3, range(3), beta=1.0), lambda a, b=None: orc.spd_ai_laplace_K(a, b, beta=1.0)),
        'Kstein': (SpdSteinGaussianKernel(3, range(3), beta=1.0), lambda a, b=None: orc.spd_stein_K(a, b, beta=1.0)),
        'Kfrob': (SpdFrobeniusGaussianKernel(3, range(3), beta=1.0), lambda a, b=None: orc.spd_frobenius_K(a, b, beta=1.0)),
        'Klogeuc': (SpdLogEuclideanGaussianKernel(3, range(3), beta=1.0), lambda a, b=None: orc.spd_logeuclid_K(a, b, beta=1.0)),
    }
    r = {}
    for name, (kern, ofun) in kerns.items():
        K11 = kern.compute_K_symm(x)
        K12 = kern.compute_K(x, xt)
        r[name] = {
            '11_vs_golden': rel(K11, g[name + '11']), '11_offdiag_vs_golden': rel(K11, g[name + '11'], off),
            '11_vs_exact': rel(K11, t['s2pp_' + name + '11']), '11_offdiag_vs_exact': rel(K11, t['s2pp_' + name + '11'], off),
            '11_vs_oracle': rel(K11, ofun(x)),
            '12_vs_golden': rel(K12, g[name + '12']), '12_vs_exact': rel(K12, t['s2pp_' + name + '12']),
            '12_vs_oracle': rel(K12, ofun(x, xt)),
            'golden11_vs_exact': rel(g[name + '11'], t['s2pp_' + name + '11']),
            'oracle11_vs_exact': rel(ofun(x), t['s2pp_' + name + '11']),
        }
        if name == 'Kailap':
            # the Laplace kernel is linear in d: report the distance itself too (near-duplicate neighbours, d ~ 1e-6)
            d_gpu = -np.log(K11)
            far = t['s2pp_Dai11'] > 1e-3
            r[name]['d_abs_err_max'] = float(np.max(np.abs(d_gpu - t['s2pp_Dai11'])))
            r[name]['11_far_vs_exact'] = rel(K11, t['s2pp_Kailap11'], far)
        if name == 'Kstein':
            r[name]['kdiag_vs_exact'] = rel(kern.compute_Kdiag(x), np.diag(t['s2pp_Kstein11']))
    out['spd_s2pp'] = r
    g = dict(np.load(os.path.join(G, 'spd_s6pp.npz')))
    x = g['x']
    kerns = {
        'Kai': (SpdAffineInvariantGaussianKernel(21, range(21), beta_min=0.5, beta=0.5), lambda a: orc.spd_ai_gaussian_K(a, beta=1.0)),
        'Kstein': (SpdSteinGaussianKernel(21, range(21), beta=3.0), lambda a: orc.spd_stein_K(a, beta=3.0)),
        'Klogeuc': (SpdLogEuclideanGaussianKernel(21, range(21), beta=1.0), lambda a: orc.spd_logeuclid_K(a, beta=1.0)),
    }
    r = {}
    for name, (kern, ofun) in kerns.items():
        K = kern.compute_K_symm(x)
        r[name] = {'vs_golden': rel(K, g[name]), 'vs_exact': rel(K, t['s6pp_' + name]), 'vs_oracle': rel(K, ofun(x)),
                   'golden_vs_exact': rel(g[name], t['s6pp_' + name]), 'oracle_vs_exact': rel(ofun(x), t['s6pp_' + name])}
    out['spd_s6pp'] = r
    # synthetic, all dimensions, vs oracle (as tests/test_gpu_gram.py::test_spd_ai_all_dims_vs_oracle)
    r = {}
    for d in range(2, 9):
        m = d * (d + 1) // 2
        X = orc.synth_spd_mandel(150, d, seed=100 + d)
        X2 = orc.synth_spd_mandel(70, d, seed=200 + d)
        beta = 0.5 * (d - 1) + 0.5
        r[f'd{d}'] = {
            'ai_gauss': rel(SpdAffineInvariantGaussianKernel(m, range(m), beta_min=0.25, beta=0.75, variance=1.3).compute_K(X, X2),
                            orc.spd_ai_gaussian_K(X, X2, beta=1.0, variance=1.3)),
            'ai_laplace': rel(SpdAffineInvariantLaplaceKernel(m, range(m), beta=1.0).compute_K(X, X2), orc.spd_ai_laplace_K(X, X2, beta=1.0)),
            'stein': rel(SpdSteinGaussianKernel(m, range(m), beta=beta).compute_K(X, X2), orc.spd_stein_K(X, X2, beta=beta)),
            'stein_kdiag': rel(SpdSteinGaussianKernel(m, range(m), beta=beta).compute_Kdiag(X), orc.stein_kdiag(X, beta=beta)),
            'frobenius': rel(SpdFrobeniusGaussianKernel(m, range(m), beta=0.3).compute_K(X, X2), orc.spd_frobenius_K(X, X2, beta=0.3)),
            'log_euclid': rel(SpdLogEuclideanGaussianKernel(m, range(m), beta=1.0).compute_K(X, X2), orc.spd_logeuclid_K(X, X2, beta=1.0)),
        }
    out['spd_synthetic_vs_oracle'] = r
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
