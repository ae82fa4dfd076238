"""Latency of the small configurations (BASELINE.json configs 1-3: N = 5..100+, the reference's real workload is thousands of
tiny likelihood / acquisition evaluations per BO run): microseconds per evaluation, GPU paths beside the CPU oracle.

    python tools/small_n_bench.py [--out gpurun_out/r02_small_n.json]

Reported per training-set size N (S^2, D = 3, sphere Gaussian kernel; SPD affine-invariant on S^2_++ for one size):
  tiled_*        the general path (Gram kernel -> gabo_potrf_f64 -> gabo_trsm_f64 -> gabo_gp_loglik_f64 [+ analytic gradients]),
                 one fresh hyper-parameter set per call (the factor cache never hits), through GPR.compute_log_likelihood;
  fused_b1_*     GPR.log_likelihood_batch with ONE set: value + 4 gradients in one launch (host round trip included);
  fused_bB_*     the same launch with B sets (multi-start / grid): microseconds per set, device time and host round trip;
  acq_*          ExpectedImprovement.evaluate_with_gradients at M = 1 and M = 100 candidates (fused one-launch acquisition);
  cpu_*          the oracle (NumPy / SciPy, gabo_oracle.py) doing the same evaluation on the host: Gram + Cholesky + solve +
                 log-density (value only -- the reference adds a TensorFlow gradient pass on top), posterior + EI.
Times are medians over repeated calls after warm-up; GPU numbers through the public Python API include the host<->device
round trip of the hyper-parameters and the result."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def median_us(fn, reps, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts) * 1e6)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='gpurun_out/r02_small_n.json')
    ap.add_argument('--reps', type=int, default=200)
    args = ap.parse_args()
    import torch
    from gaboflow_b200 import GPR, SphereGaussianKernel, SpdAffineInvariantGaussianKernel, ExpectedImprovement, _lib, ops
    from oracle import gabo_oracle as orc
    assert torch.cuda.is_available()
    lib = _lib.lib()
    res = {'device': torch.cuda.get_device_name(0), 'host_threads': os.cpu_count(), 'cases': []}
    rng = np.random.default_rng(0)

    def run_case(name, X, y, make_kernel, K_oracle, acq=True):
        n = X.shape[0]
        case = {'case': name, 'n': int(n)}
        k = make_kernel()
        m = GPR(X, y, kern=k)
        m.likelihood.variance = 0.05
        bname = k.beta_param_name()
        par = getattr(k, bname)
        state = {'i': 0}

        def fresh():
            state['i'] += 1
            par.assign(1.0 + 0.001 * (state['i'] % 500))

        def tiled_ll():
            fresh()
            return m.compute_log_likelihood()

        def tiled_ll_grad():
            fresh()
            v = m.compute_log_likelihood()
            return v, m.loglik_gradients()

        def fused_b1():
            fresh()
            return m.log_likelihood_batch(with_grad=True)

        n0 = lib.gabo_launch_count(); tiled_ll(); case['tiled_loglik_launches'] = int(lib.gabo_launch_count() - n0)
        n0 = lib.gabo_launch_count(); tiled_ll_grad(); case['tiled_loglik_grad_launches'] = int(lib.gabo_launch_count() - n0)
        fused_b1()
        n0 = lib.gabo_launch_count(); fused_b1(); case['fused_launches'] = int(lib.gabo_launch_count() - n0)
        case['tiled_loglik_us'] = median_us(tiled_ll, args.reps)
        case['tiled_loglik_grad_us'] = median_us(tiled_ll_grad, args.reps)
        case['fused_b1_value_grad_us'] = median_us(fused_b1, args.reps)
        # parity of what is being timed
        r = m.log_likelihood_batch(with_grad=True)
        want = orc.gp_loglik(K_oracle(X, float(k.beta_effective()), float(k.variance)), y, 0.05)
        case['fused_loglik_rel_err_vs_oracle'] = float(abs(r['loglik'][0] - want) / abs(want))
        for B in (148, 1184):
            betas = rng.uniform(0.5, 4.0, B)
            vs = rng.uniform(0.5, 2.0, B)
            s2 = rng.uniform(0.01, 0.2, B)
            case[f'fused_b{B}_api_us_per_set'] = median_us(
                lambda: m.log_likelihood_batch(beta=betas, variance=vs, noise_variance=s2, with_grad=True), 30) / B
            G, beta0 = m._log_gram()
            theta = torch.from_numpy(np.stack([betas / beta0, vs, s2, np.zeros(B)], axis=1)).cuda()
            out = torch.empty((B, 5), dtype=torch.float64, device='cuda')
            info = torch.empty((B,), dtype=torch.int32, device='cuda')
            for want_grad, key in ((True, 'value_grad'), (False, 'value')):
                for _ in range(3):
                    ops.gp_small_batch(G, m.Y, theta, want_grad=want_grad, out=out, info=info)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(20):
                    ops.gp_small_batch(G, m.Y, theta, want_grad=want_grad, out=out, info=info)
                e1.record()
                torch.cuda.synchronize()
                case[f'fused_b{B}_{key}_device_us_per_set'] = e0.elapsed_time(e1) * 1e3 / 20 / B
        # device time of ONE set
        theta1 = torch.tensor([[1.0, 1.0, 0.05, 0.0]], dtype=torch.float64, device='cuda')
        out1 = torch.empty((1, 5), dtype=torch.float64, device='cuda')
        info1 = torch.empty((1,), dtype=torch.int32, device='cuda')
        G, _ = m._log_gram()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            ops.gp_small_batch(G, m.Y, theta1, want_grad=True, out=out1, info=info1)
        e1.record()
        torch.cuda.synchronize()
        case['fused_b1_value_grad_device_us'] = e0.elapsed_time(e1) * 1e3 / 50

        # CPU oracle: value only
        def cpu_ll():
            state['i'] += 1
            b = 1.0 + 0.001 * (state['i'] % 500)
            return orc.gp_loglik(K_oracle(X, b, 1.0), y, 0.05)
        case['cpu_loglik_us'] = median_us(cpu_ll, max(20, args.reps // 4))

        if acq:
            par.assign(2.0)
            ei = ExpectedImprovement(m)
            D = X.shape[1]
            for M in (1, 100):
                Xc = orc.synth_sphere(M, D, seed=77)
                Xc_d = torch.from_numpy(Xc).cuda()
                case[f'acq_grad_m{M}_us'] = median_us(lambda: ei.evaluate_with_gradients(Xc_d), args.reps)
                n0 = lib.gabo_launch_count(); ei.evaluate_with_gradients(Xc_d)
                case[f'acq_grad_m{M}_launches'] = int(lib.gabo_launch_count() - n0)
                Kt = K_oracle(X, 2.0, 1.0)
                fmin = ei.fmin

                def cpu_acq():
                    Kx = orc.sphere_gaussian_K(X, Xc, beta=2.0)
                    mean, var = orc.gp_predict(Kt, Kx, np.ones(M), y, 0.05)
                    return orc.expected_improvement(mean, var, fmin)
                case[f'cpu_acq_value_m{M}_us'] = median_us(cpu_acq, max(20, args.reps // 4))
        res['cases'].append(case)
        print(json.dumps(case), flush=True)

    for n in (30, 100, 160):
        X = orc.synth_sphere(n, 3, seed=n)
        y = rng.standard_normal((n, 1))
        run_case(f'sphere_gaussian_S2_n{n}', X, y,
                 lambda: SphereGaussianKernel(3, range(3), beta_min=0.0, beta=1.0, variance=1.0),
                 lambda X_, b, v: orc.sphere_gaussian_K(X_, beta=b, variance=v))
    n = 60
    Xm = orc.synth_spd_mandel(n, 2, seed=3)
    y = rng.standard_normal((n, 1))
    run_case(f'spd_affine_invariant_S2pp_n{n}', Xm, y,
             lambda: SpdAffineInvariantGaussianKernel(3, range(3), beta_min=0.0, beta=1.0, variance=1.0),
             lambda X_, b, v: orc.spd_ai_gaussian_K(X_, beta=b, variance=v), acq=False)
    os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
    with open(args.out, 'w') as f:
        json.dump(res, f, indent=1)
    print('wrote', args.out)


if __name__ == '__main__':
    main()
