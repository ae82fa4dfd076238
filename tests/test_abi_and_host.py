"""CPU: the C-ABI library loads and exports every symbol include/gabo_b200.h declares (no compute calls), and the
host-side mirror of the reference interface keeps the reference's semantics.  No GPU needed."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'gabo_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gabo_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_expected_surface():
    names = declared_functions()
    for must in ['gabo_points_gram_f64', 'gabo_points_gram_f32', 'gabo_spd_prep_f64', 'gabo_spd_gram_f64', 'gabo_potrf_f64',
                 'gabo_trsm_f64', 'gabo_logdet_f64', 'gabo_gp_loglik_f64', 'gabo_gp_predict_f64', 'gabo_kdiag_const_f64']:
        assert must in names


def test_library_builds_and_exports_every_declared_symbol():
    from gaboflow_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.lib()
    names = declared_functions()
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.gabo_abi_version() == _lib.ABI_VERSION == 2
    assert lib.gabo_status_string(0) == b'ok'
    assert b'argument' in lib.gabo_status_string(-3)
    # pure size queries run on the host
    assert lib.gabo_points_gram_workspace_bytes(65536, 65536, 51) >= 2 * 65536 * 52 * 8
    assert lib.gabo_potrf_workspace_bytes(32768) >= 128 * 128 * 8
    assert lib.gabo_trsm_workspace_bytes(32768, 1) >= 256 * 128 * 128 * 8


def test_library_is_sm100a_only():
    import subprocess
    from gaboflow_b200 import build
    out = subprocess.run(['cuobjdump', '-lelf', build.build()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    archs = set(re.findall(r'sm_\d+a?', out.stdout))
    assert archs == {'sm_100a'}, archs


def test_kernel_classes_mirror_reference_parameters():
    from gaboflow_b200 import (SphereGaussianKernel, SphereLaplaceKernel, SpdSteinGaussianKernel,
                               SpdAffineInvariantGaussianKernel, SpdAffineInvariantLaplaceKernel,
                               SpdFrobeniusGaussianKernel, SpdLogEuclideanGaussianKernel)
    k = SphereGaussianKernel(input_dim=3, active_dims=range(3), beta_min=7.0, beta=10.0, variance=1.)
    # kernels_sphere_tf.py:55-57,79: the Param holds beta as given, beta_min is added in K -> effective 17
    assert float(k.beta_shifted) == 10.0 and k.beta_min_value == 7.0 and k._beta_eff() == 17.0
    gb = k.get_beta()
    assert isinstance(gb, np.ndarray) and gb.shape == (1,) and gb[0] == 17.0           # :138 returns Param.value + beta_min
    k.update_beta(9.5)                                                                    # :122
    assert float(k.beta_shifted) == 2.5 and k.get_beta()[0] == 9.5
    k.variance = 2.0                                                                      # Param assignment updates in place
    assert float(k.variance) == 2.0 and 'variance' in k.params()
    kl = SphereLaplaceKernel(3, range(3), beta=4.0)
    assert float(kl.beta) == 4.0 and not hasattr(kl, 'beta_min_value')
    ks = SpdSteinGaussianKernel(input_dim=21, active_dims=range(21), beta=3.0)
    assert ks.matrix_dim == 6 and ks.continuous_param_space_limit == 2.5                  # kernels_spd_tf.py:53,59
    assert float(ks.beta_shifted) == 0.5 and ks.get_beta()[0] == 3.0                      # :67,162
    ks.update_beta(4.0)
    assert float(ks.beta_shifted) == 1.5
    ka = SpdAffineInvariantGaussianKernel(input_dim=3, active_dims=range(3), beta=1.0, variance=1., beta_min=0.6)
    assert ka.matrix_dim == 2 and abs(ka._beta_eff() - 1.6) < 1e-15
    kal = SpdAffineInvariantLaplaceKernel(3, range(3))
    assert kal.beta_min_value == 0.0 and kal._beta_eff() == 1.0                           # :342 default beta_min = 0
    for cls in (SpdFrobeniusGaussianKernel, SpdLogEuclideanGaussianKernel):
        kk = cls(input_dim=6, active_dims=range(6), beta=0.7)
        assert kk.matrix_dim == 3 and float(kk.beta_param) == 0.7
    assert k.active_dims == [0, 1, 2] and k.input_dim == 3


def test_no_cpu_fallback():
    """Without a CUDA device every product call fails loudly (and nothing under gaboflow_b200 imports the oracle)."""
    import torch
    import gaboflow_b200
    from gaboflow_b200 import SphereGaussianKernel, ops
    pkg = os.path.dirname(gaboflow_b200.__file__)
    for fn in os.listdir(pkg) + [os.path.join('kernel_utils', f) for f in os.listdir(os.path.join(pkg, 'kernel_utils'))]:
        if fn.endswith('.py'):
            assert 'oracle' not in open(os.path.join(pkg, fn)).read(), fn
    with pytest.raises(TypeError):
        ops.points_gram(ops.SPHERE_GAUSSIAN, torch.zeros((4, 3), dtype=torch.float64))    # CPU tensor rejected
    if not torch.cuda.is_available():
        k = SphereGaussianKernel(3, range(3), beta_min=1.0)
        with pytest.raises(RuntimeError):
            k.compute_K_symm(np.eye(3))


def test_bench_reference_arm_runs_small():
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                          '--n-gram', '2048', '--n-chol', '1024', '--ref-gram-rows', '256', '--ref-chol-n', '512', '--ref-spd-rows', '2'],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['value'] > 0 and line['cpu_baseline']['kind'] == 'port'
    assert line['e2e']['h2d_bytes_per_step'] == 0
    # both arms print the same `config` dict (the driver compares them): workload, parameters, seeds and the L2 note only
    assert set(line['config']) == {'workload', 'beta', 'variance', 'sigma2', 'seeds', 'l2'}
    assert line['config']['workload'].startswith('sphere Gaussian Gram N=2048 on S^50 (D=51) + FP64 Cholesky of K+1e-2 I at N=1024')
    assert line['metric'] == 'sphere_gram_entries_per_s' and line['unit'] == 'entries/s' and line['higher_is_better'] is True


def test_gram_engine_and_multi_gpu_mode_selection():
    """Host-side dispatch only (no device work): which Gram kernel / multi-GPU mode a problem gets."""
    import torch
    from gaboflow_b200 import ops
    from gaboflow_b200.dist import default_gram_mode
    f64 = torch.float64
    assert ops.gram_engine(ops.SPHERE_GAUSSIAN, f64, 51, 65536, 65536) == 'i8'
    assert ops.gram_engine(ops.SPHERE_LAPLACE, f64, 64, 4096, 4096) == 'i8'
    assert ops.gram_engine(ops.SPHERE_GAUSSIAN, f64, 65, 65536, 65536) == 'dmma'         # more than 64 coordinates
    assert ops.gram_engine(ops.SPHERE_GAUSSIAN, f64, 51, 30, 30) == 'dmma'               # launch-bound sizes (configs 1-3)
    assert ops.gram_engine(ops.SQDIST_GAUSSIAN, f64, 21, 16384, 16384) == 'dmma'         # not a sphere kind
    assert ops.gram_engine(ops.SPHERE_GAUSSIAN, torch.float32, 51, 65536, 65536) == 'dmma'
    assert ops.gram_engine(ops.SPHERE_GAUSSIAN, f64, 51, 65536, 65536, engine='dmma') == 'dmma'
    with pytest.raises(ValueError):
        ops.gram_engine(ops.SPHERE_GAUSSIAN, f64, 100, 65536, 65536, engine='i8')
    assert default_gram_mode(2, 65536, 51) == 'staged'
    assert default_gram_mode(4, 65536, 51) == 'i8rows' and default_gram_mode(8, 65536, 51) == 'i8rows'
    assert default_gram_mode(8, 65536 + 100, 51) == 'direct' and default_gram_mode(8, 65536, 100) == 'direct'
    assert default_gram_mode(3, 65536, 51) == 'direct'
