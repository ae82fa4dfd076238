"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own code.

Run in the build container only (needs /root/reference; the GPU box has no reference):

    python tests/golden/make_golden.py

What runs from the reference (imported, unmodified, from /root/reference):
  * BoManifolds/Riemannian_utils/sphere_utils.py : sphere_distance (:64-86), logmap
  * BoManifolds/Riemannian_utils/SPD_utils.py    : vector_to_symmetric_matrix_mandel (:79-101),
    symmetric_matrix_to_vector_mandel (:57-76), affine_invariant_distance (:180-197), expmap (:104-120)
The TF kernel classes themselves cannot run here (TensorFlow 1 / GPflow 0.5 absent), so each
kernel matrix is assembled pair by pair from the reference distance functions with the
elementwise formula of the kernel class (file:line quoted at each block) -- the same
loop the reference keeps, commented out, as its own check (kernels_spd_tf.py:300-316,
552-558, 660-666).  Inputs are the reference examples' seeded data
(examples/kernels/sphere/sphere_kernels.py:79-114, examples/kernels/spd/spd_kernels.py:84-132).
"""
import os
import sys
import warnings

import numpy as np
import scipy.linalg as sla
from scipy.io import loadmat

REF = '/root/reference'
sys.path.insert(0, REF)
warnings.simplefilter('ignore')
from BoManifolds.Riemannian_utils import sphere_utils as ref_sphere   # noqa: E402
from BoManifolds.Riemannian_utils import SPD_utils as ref_spd         # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sphere_fixture():
    # sphere_kernels.py:79-114 (same statements, same RNG call order)
    np.random.seed(1234)
    mu_test_fct = np.array([1 / np.sqrt(2), 1 / np.sqrt(2), 0])
    nb_data, dim = 20, 3
    mean = np.array([1, 0, 0])
    mean = mean / np.linalg.norm(mean)
    cov = 0.1 * np.eye(dim)
    data = np.random.multivariate_normal(mean, cov, nb_data)
    x_man = data / np.linalg.norm(data, axis=1)[:, None]

    sigma = np.array([[0.6, 0.2, 0], [0.2, 0.3, -0.01], [0, -0.01, 0.2]])   # sphere_kernels.py:33

    def test_function(x, mu):                                               # sphere_kernels.py:30-41
        xp = ref_sphere.logmap(x, mu)
        return np.exp(-0.5 * np.dot(xp.T, np.dot(np.linalg.inv(sigma), xp))) / np.sqrt(
            (2 * np.pi) ** dim * np.linalg.det(sigma))

    y_train = np.zeros((nb_data, 1))
    for n in range(nb_data):
        y_train[n] = test_function(x_man[n], mu_test_fct)
    nb_test = 10
    mean_test = mu_test_fct / np.linalg.norm(mean)
    data = np.random.multivariate_normal(mean_test, 0.1 * np.eye(dim), nb_test)
    x_test = data / np.linalg.norm(data, axis=1)[:, None]

    def dist(A, B):
        D = np.zeros((A.shape[0], B.shape[0]))
        for i in range(A.shape[0]):
            for j in range(B.shape[0]):
                D[i, j] = float(np.squeeze(ref_sphere.sphere_distance(A[i], B[j])))
        return D

    out = dict(x_man=x_man, x_test=x_test, y_train=y_train)
    for tag, A, B in (('11', x_man, x_man), ('12', x_man, x_test), ('22', x_test, x_test)):
        D = dist(A, B)
        out['D' + tag] = D
        # kernels_sphere_tf.py:79-88, beta_eff = beta + beta_min = 10 + 7 (sphere_kernels.py:125)
        out['Kgauss' + tag] = 1.0 * np.exp(-(D * D) * 17.0)
        # kernels_sphere_tf.py:195-204, beta = 10 (sphere_kernels.py:147 uses the same style)
        out['Klaplace' + tag] = 1.0 * np.exp(-D * 10.0)
    np.savez(os.path.join(HERE, 'sphere_s2.npz'), **out)
    return out


def spd_fixture():
    # spd_kernels.py:84-132
    data_demos = loadmat(os.path.join(REF, 'data/2Dletters/C.mat'))['demos'][0]
    demos = [data_demos[i]['pos'][0][0] for i in range(data_demos.shape[0])]
    nb_init = demos[0].shape[1]
    time = np.hstack([np.arange(0, nb_init) * 1.0] * data_demos.shape[0])
    data_eucl = np.vstack((time, np.hstack(demos)))[:, :nb_init]
    spd_mandel = [ref_spd.symmetric_matrix_to_vector_mandel(
        ref_spd.expmap(0.01 * np.dot(data_eucl[1:, n][:, None], data_eucl[1:, n][None]), np.eye(2)))[:, None]
        for n in range(data_eucl.shape[1])]
    spd_mandel = np.vstack((data_eucl[0], np.concatenate(spd_mandel, axis=1)))
    data = spd_mandel[:, ::2]
    data = np.delete(data, np.hstack((np.arange(24, 37), np.arange(68, 76))), axis=1)
    y = data[0][:, None]
    x_man = np.ascontiguousarray(np.real(data[1:]).T)                    # [79,3] Mandel rows
    y_test = spd_mandel[0, ::2][:, None]
    x_test = np.ascontiguousarray(np.real(spd_mandel[1:, ::2]).T)        # [100,3]

    def mats(V):
        return np.stack([ref_spd.vector_to_symmetric_matrix_mandel(v) for v in V])

    S1, S2 = mats(x_man), mats(x_test)
    out = dict(x_man=x_man, x_test=x_test, y=np.real(y), y_test=np.real(y_test), S_man=S1, S_test=S2)
    for tag, A, B in (('11', S1, S1), ('12', S1, S2)):
        n1, n2 = A.shape[0], B.shape[0]
        Dai = np.zeros((n1, n2)); Kst = np.zeros((n1, n2)); Dfr = np.zeros((n1, n2)); Dle = np.zeros((n1, n2))
        logB = [np.real(sla.logm(B[j])) for j in range(n2)]
        for i in range(n1):
            logA = np.real(sla.logm(A[i]))
            for j in range(n2):
                Dai[i, j] = np.real(ref_spd.affine_invariant_distance(A[i], B[j]))     # SPD_utils.py:196-197
                # kernels_spd_tf.py:104-112 (as coded), beta = 1
                Kst[i, j] = np.linalg.det(A[i] @ B[j]) ** 1.0 / np.linalg.det(0.5 * (A[i] + B[j])) ** 1.0
                Dfr[i, j] = np.linalg.norm(A[i] - B[j])                                # kernels_spd_tf.py:552-558
                Dle[i, j] = np.linalg.norm(logA - logB[j])                             # kernels_spd_tf.py:660-666
        out['Dai' + tag] = Dai
        out['Kai' + tag] = np.exp(-(Dai * Dai) * 1.6)        # kernels_spd_tf.py:239-248, beta_eff 1.0+0.6 (spd_kernels.py:177)
        out['Kailap' + tag] = np.exp(-Dai * 1.0)             # kernels_spd_tf.py:393-400
        out['Kstein' + tag] = Kst
        out['Kfrob' + tag] = np.exp(-(Dfr * Dfr) * 1.0)      # kernels_spd_tf.py:520-533
        out['Klogeuc' + tag] = np.exp(-(Dle * Dle) * 1.0)    # kernels_spd_tf.py:628-641 in FP64
    np.savez(os.path.join(HERE, 'spd_s2pp.npz'), **out)
    return out


def spd6_fixture():
    """48 well-conditioned 6x6 SPD matrices (recipe of constrained_optimization.py:246-274,
    legacy RNG seed 1234), pairwise through the reference NumPy distance."""
    np.random.seed(1234)
    n, d = 48, 6
    w = 0.5 + 1.5 * np.random.rand(n, d)
    V = []
    for i in range(n):
        u, _ = np.linalg.qr(np.random.randn(d, d))
        V.append(ref_spd.symmetric_matrix_to_vector_mandel(np.dot(u, np.dot(np.diag(w[i]), u.T))))
    V = np.array(V)
    S = np.stack([ref_spd.vector_to_symmetric_matrix_mandel(v) for v in V])
    Dai = np.zeros((n, n)); Dle = np.zeros((n, n)); Kst = np.zeros((n, n))
    logS = [np.real(sla.logm(S[j])) for j in range(n)]
    for i in range(n):
        for j in range(n):
            Dai[i, j] = np.real(ref_spd.affine_invariant_distance(S[i], S[j]))
            Dle[i, j] = np.linalg.norm(logS[i] - logS[j])
            Kst[i, j] = np.linalg.det(S[i] @ S[j]) ** 3.0 / np.linalg.det(0.5 * (S[i] + S[j])) ** 3.0
    np.savez(os.path.join(HERE, 'spd_s6pp.npz'), x=V, S=S, Dai=Dai,
             Kai=np.exp(-Dai * Dai), Klogeuc=np.exp(-Dle * Dle), Kstein=Kst)


def spd_maps_fixture():
    """Exponential / logarithm maps through the reference's NumPy twins (SPD_utils.py:104-177: expmap, logmap,
    expmap_mandel_vector, logmap_mandel_vector) for d = 2, 3, 6 -- random symmetric tangent matrices at a random SPD base
    (legacy RNG seed 4321) -- and one beta-sweep trial of examples/kernels/sphere/sphere_gaussian_kernel_parameters.py:57-97
    (dim = 4, 500 samples, seed 0; minimum numpy.linalg.eig value of the Gram assembled from the reference sphere_distance)."""
    np.random.seed(4321)
    out = {}
    for d in (2, 3, 6):
        n = 40
        A = np.random.randn(d, d)
        S = A @ A.T + 0.5 * np.eye(d)
        U = np.random.randn(n, d, d)
        U = 0.5 * (U + np.swapaxes(U, 1, 2))
        X = np.stack([np.real(ref_spd.expmap(U[i], S)) for i in range(n)])
        Uback = np.stack([np.real(ref_spd.logmap(X[i], S)) for i in range(n)])
        s_m = ref_spd.symmetric_matrix_to_vector_mandel(S)
        u_m = np.stack([ref_spd.symmetric_matrix_to_vector_mandel(U[i]) for i in range(n)])
        x_m = np.stack([np.real(ref_spd.expmap_mandel_vector(u_m[i], s_m)) for i in range(n)])
        ub_m = np.stack([np.real(ref_spd.logmap_mandel_vector(x_m[i], s_m)) for i in range(n)])
        out.update({f'd{d}_S': S, f'd{d}_U': U, f'd{d}_exp': X, f'd{d}_log_of_exp': Uback, f'd{d}_s_mandel': s_m,
                    f'd{d}_u_mandel': u_m, f'd{d}_exp_mandel': x_m, f'd{d}_log_of_exp_mandel': ub_m})
    # one sweep trial, sphere_gaussian_kernel_parameters.py:57-97 with dim = 4
    np.random.seed(0)
    dim, nb_samples, nb_sources = 4, 500, 10
    mean = np.array([np.random.randn(dim) for _ in range(nb_sources)])
    mean = mean / np.linalg.norm(mean, axis=1)[:, None]
    data = [np.random.multivariate_normal(mean[i], np.eye(dim), nb_samples // nb_sources).T for i in range(nb_sources)]
    pts = np.array([data[i][:, n] / np.linalg.norm(data[i][:, n]) for i in range(nb_sources) for n in range(nb_samples // nb_sources)])
    D = np.zeros((nb_samples, nb_samples))
    for i in range(nb_samples):
        for j in range(i + 1, nb_samples):
            D[i, j] = D[j, i] = float(np.squeeze(ref_sphere.sphere_distance(pts[i], pts[j])))
    betas = np.logspace(0, 2, 30)                                    # :50 (dim == 4)
    mins = []
    for b in betas:
        eigvals, _ = np.linalg.eig(np.exp(-b * D * D))                # kernels_sphere_tf.py:79-88 + :93-96 of the script
        mins.append(np.min(np.real(eigvals)))
    out.update(sweep_x=pts, sweep_betas=betas, sweep_min_eig=np.array(mins))
    np.savez(os.path.join(HERE, 'spd_maps_sweep.npz'), **out)
    return out


if __name__ == '__main__':
    s = sphere_fixture()
    print('sphere: x_man[0] =', s['x_man'][0].tolist(), ' K[0,1] =', s['Kgauss11'][0, 1])
    p = spd_fixture()
    print('spd: X[0] =', p['x_man'][0].tolist(), ' Kai[0,40] =', p['Kai11'][0, 40])
    spd6_fixture()
    m = spd_maps_fixture()
    print('sweep: min eig at beta[0], beta[-1] =', m['sweep_min_eig'][0], m['sweep_min_eig'][-1])
    print('wrote', sorted(f for f in os.listdir(HERE) if f.endswith('.npz')))
