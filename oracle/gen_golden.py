"""ORACLE (test infrastructure): generate tests/golden/*.npz by running the REFERENCE'S OWN
files (/root/reference, unmodified) through oracle/ref_shim.  Run in the build container:

    python oracle/gen_golden.py

/root/reference does not exist on the GPU box, so the fixtures are committed.  Every stochastic
primitive is recorded or injected so the fixtures are deterministic functions of their inputs:
  * omega: reg.omega is patched to return E[PG(1,psi)] (SURVEY Appendix D convention) or a
    recorded vector;
  * the a-scan permutation (regression.py:286) and the categorical uniforms (:315) are recorded
    by wrapping numpy.random.permutation and the shim's sample_discrete_from_log;
  * the Gaussian draw's normals (:334) are recorded by wrapping numpy.random.randn.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "ref_shim"))
import sitecustomize_shim  # noqa: E402,F401  (numpy aliases + sys.path for the reference)

import warnings  # noqa: E402
warnings.filterwarnings("ignore", category=SyntaxWarning)

import numpy as np  # noqa: E402
import pybasicbayes.util.stats as shim_stats  # noqa: E402
import pyglm.regression as ref_reg  # noqa: E402
from pyglm.models import SparseBernoulliGLM, NonlinearAutoregressiveModel  # noqa: E402
from pyglm.regression import SparseBernoulliRegression  # noqa: E402
from pyglm.utils.basis import cosine_basis, convolve_with_basis  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def omega_mean(psi):
    out = np.full_like(psi, 0.25)
    nz = np.abs(psi) > 1e-12
    out[nz] = np.tanh(psi[nz] / 2.0) / (2.0 * psi[nz])
    return out


class Recorder(object):
    """Records perm / uniforms / normals consumed by the reference's resample()."""

    def __enter__(self):
        self.perms, self.us, self.zs, self.lps = [], [], [], []
        self._perm, self._randn = np.random.permutation, np.random.randn
        self._sdfl = ref_reg.sample_discrete_from_log

        def permutation(n):
            p = self._perm(n)
            self.perms.append(np.array(p))
            return p

        def randn(*shape):
            z = self._randn(*shape)
            self.zs.append(np.array(z))
            return z

        def sdfl(lps):
            u = np.random.random()
            self.us.append(u)
            self.lps.append(np.array(lps))
            saved = np.random.random
            np.random.random = lambda size=None: np.full(size, u)
            try:
                return self._sdfl(lps)
            finally:
                np.random.random = saved

        np.random.permutation = permutation
        np.random.randn = randn
        ref_reg.sample_discrete_from_log = sdfl
        return self

    def __exit__(self, *exc):
        np.random.permutation, np.random.randn = self._perm, self._randn
        ref_reg.sample_discrete_from_log = self._sdfl


def kat_case(N, B, L, T, store_arrays):
    """SURVEY Appendix D recipe, evaluated by the reference's own functions."""
    Y = (np.random.default_rng(0).random((T, N)) < 0.05).astype(float)
    basis = cosine_basis(B, L) / L
    X = convolve_with_basis(Y, basis)
    np.random.seed(1)
    reg = SparseBernoulliRegression(N, B, S_w=10.0, mu_b=-2.)
    reg.W = 0.1 * (np.arange(N)[:, None] + 1) * ((-1.0) ** np.arange(B))[None, :]
    reg.b = np.array([-2.0])
    reg.a = np.array([False] + [True] * (N - 1))
    y = Y[:, N - 1]
    psi = reg.activation(X)
    om = omega_mean(psi)
    reg.omega = lambda X_, y_: om
    ll = reg.log_likelihood((X, y)).sum()
    J_l, h_l = reg._lkhd_sufficient_statistics([(X, y)])
    J_0, h_0 = reg._prior_sufficient_statistics()
    ml_a = reg._marginal_likelihood(J_0, h_0, J_0 + J_l, h_0 + h_l)
    a_keep = reg.a.copy()
    reg.a = np.ones(N, dtype=bool)
    ml_ones = reg._marginal_likelihood(J_0, h_0, J_0 + J_l, h_0 + h_l)
    reg.a = a_keep.copy()      # the scan below mutates reg.a in place
    out = dict(N=N, B=B, L=L, T=T, sumY=Y.sum(), sumX=X.sum(), maxX=X.max(), sumpsi=psi.sum(), ll=ll,
               trJ=np.trace(J_l), froJ=np.linalg.norm(J_l), Jcorner=J_l[-1, -1], J10=J_l[1, 0],
               sumh=h_l.sum(), h0=h_l[0], ml_a=ml_a, ml_ones=ml_ones,
               basis=basis, W=reg.W.copy(), b=reg.b.copy(), a=a_keep.copy())
    if store_arrays:
        out.update(Y=Y.astype(np.uint8), X=X, psi=psi, omega=om, J_lkhd=J_l, h_lkhd=h_l,
                   J_prior=J_0, h_prior=h_0,
                   ll_terms=reg.log_likelihood((X, y)), mean=reg.mean(X))
        # a-scan + W draw with recorded randomness, rho = 0.5 (regression.py:282-340)
        np.random.seed(7)
        with Recorder() as rec:
            reg._collapsed_resample_a(J_0, h_0, J_0 + J_l, h_0 + h_l)
            a_scan = reg.a.copy()
            reg._resample_W(J_0 + J_l, h_0 + h_l)
        z_full = np.zeros(N * B + 1)
        z_full[np.concatenate((np.repeat(a_scan, B), [1])).astype(bool)] = rec.zs[0]
        out.update(scan_a0=a_keep, scan_perm=rec.perms[0], scan_us=np.array(rec.us),
                   scan_lps=np.array(rec.lps), scan_a=a_scan, draw_z=z_full,
                   draw_W=reg.W.copy(), draw_b=reg.b.copy())
    else:
        # only a sparse sample of X rows for the big case
        rows = np.array([0, 1, 2, 99, 100, 101, 5000, T - 1])
        out.update(X_rows=rows, X_sample=X[rows])
    return out


def reference_tests():
    """The two assertions of test/test_generate.py, with their inputs/outputs recorded."""
    out = {}
    np.random.seed(3)
    N, B, L = 2, 3, 10
    basis = cosine_basis(B, L=L) / L
    regs = [SparseBernoulliRegression(N, B, mu_b=-2, S_b=0.1) for _ in range(N)]
    model = NonlinearAutoregressiveModel(N, regs, basis=basis)
    X, Y = model.generate(T=1000, keep=False)
    model.add_data(Y)
    Xtest = model.data_list[0][0]
    assert np.allclose(X, Xtest)                                   # test_generate.py:24
    means = model.means
    model.data_list[0] = (X, Y)
    assert np.allclose(means, model.means)                          # test_generate.py:26-29
    out.update(tm_basis=basis, tm_Y=Y.astype(np.uint8), tm_X_generate=X, tm_X_conv=Xtest,
               tm_means=means[0], tm_A=model.adjacency, tm_W=model.weights, tm_b=model.biases,
               tm_ll=model.log_likelihood())
    np.random.seed(4)
    regs = [SparseBernoulliRegression(N, B, mu_b=-2, S_b=0.1) for _ in range(N)]
    model = NonlinearAutoregressiveModel(N, regs, B=B)
    X, Y = model.generate(T=1000, keep=False)
    for n in range(N):
        for b in range(B):
            assert np.allclose(Y[:-(b + 1), n], X[(b + 1):, n, b])  # test_generate.py:55
    out.update(tb_Y=Y.astype(np.uint8), tb_X=X, tb_W=model.weights, tb_b=model.biases,
               tb_A=model.adjacency)
    return out


def full_sweep_case():
    """One SparseBernoulliGLM.resample_model() (models.py:224-236) on README-like data with
    omega = E[PG] and all other draws recorded, neuron by neuron."""
    np.random.seed(11)
    T, N, B, L = 2000, 5, 2, 20
    basis = cosine_basis(B=B, L=L) / L
    true = SparseBernoulliGLM(N, basis=basis, regression_kwargs=dict(S_w=10.0, mu_b=-2.))
    for n in range(N):
        true.regressions[n].a[n] = True
        true.regressions[n].W[n, :] = -2.0
    _, Y = true.generate(T=T, keep=True)
    m = SparseBernoulliGLM(N, basis=basis, regression_kwargs=dict(S_w=10.0, mu_b=-2., rho=0.3))
    m.add_data(Y)
    out = dict(N=N, B=B, L=L, T=T, basis=basis, Y=Y.astype(np.uint8),
               A0=m.adjacency.copy(), W0=m.weights.copy(), b0=m.biases.copy(),
               ll0=m.log_likelihood(), means0=m.means[0])
    omegas, perms, us, zs = [], [], [], []
    for n, reg in enumerate(m.regressions):
        X, Yd = m.data_list[0]
        psi = reg.activation(X)
        om = omega_mean(psi)
        reg.omega = (lambda o: (lambda X_, y_: o))(om)
        omegas.append(om)
        with Recorder() as rec:
            reg.resample([(X, Yd[:, n])])
        perms.append(rec.perms[0])
        us.append(np.array(rec.us))
        z_full = np.zeros(N * B + 1)
        z_full[np.concatenate((np.repeat(reg.a, B), [1])).astype(bool)] = rec.zs[0]
        zs.append(z_full)
    out.update(omega=np.array(omegas).T, perm=np.array(perms), us=np.array(us), z=np.array(zs),
               A1=m.adjacency.copy(), W1=m.weights.copy(), b1=m.biases.copy(),
               ll1=m.log_likelihood(), means1=m.means[0])
    # hyper-parameters the regressions were constructed with
    r0 = m.regressions[0]
    out.update(rho=r0.rho, mu_w=r0.mu_w, S_w=r0.S_w, mu_b=r0.mu_b, S_b=r0.S_b)
    # network step (networks.py:132-149) shapes + push-down (models.py:232-236)
    np.random.seed(5)
    m.resample_network()
    out.update(net_mu_W=m.network.mu_W, net_sigma_W=m.network.sigma_W, net_rho=m.network.rho,
               reg0_S_w=m.regressions[0].S_w, reg0_mu_w=m.regressions[0].mu_w)
    return out


def generate_case():
    """The reference's own generate() (models.py:98-151) on a network with excitatory and inhibitory weights; the
    uniforms it consumed are numpy.random.rand(T, N) under the same seed (rvs draws N per step, regression.py:539)."""
    N, B, L, T = 5, 2, 15, 400
    rng = np.random.default_rng(42)
    basis = cosine_basis(B=B, L=L) / L
    model = SparseBernoulliGLM(N, basis=basis, regression_kwargs=dict(S_w=10.0, mu_b=-2.))
    for n, reg in enumerate(model.regressions):
        reg.a = rng.random(N) < 0.6
        reg.W = rng.standard_normal((N, B)) * 2.0
        reg.W[n, :] = -2.0
        reg.b = np.array([-1.5 + 0.2 * n])
    np.random.seed(1234)
    X, Y = model.generate(T=T, keep=False)
    np.random.seed(1234)
    U = np.random.rand(T, N)
    return dict(basis=basis, weights=model.weights, biases=model.biases, adjacency=model.adjacency, U=U, X=X,
                Y=np.asarray(Y, dtype=np.float64))


def gaussian_case():
    """SparseGaussianGLM (regression.py:380-446, models.py:274-276): the reference's own sufficient statistics,
    log-likelihood, means, one resample() per regression with every draw recorded (permutation, categorical
    uniforms, Gaussian normals, and the (alpha, beta) handed to sample_invgamma together with the eta it returned)."""
    from pyglm.models import SparseGaussianGLM
    np.random.seed(21)
    T, N, B, L = 1500, 4, 2, 15
    basis = cosine_basis(B=B, L=L) / L
    true = SparseGaussianGLM(N, basis=basis, regression_kwargs=dict(S_w=4.0, mu_b=0.2, eta=0.3))
    for n, reg in enumerate(true.regressions):
        reg.a[:] = False
        reg.W[:] = 0.0
        reg.a[n] = True
        reg.W[n, :] = 0.2
        reg.a[(n + 1) % N] = True
        reg.W[(n + 1) % N, :] = -0.15
        reg.b[:] = 0.2
    _, Y = true.generate(T=T, keep=False)
    m = SparseGaussianGLM(N, basis=basis, regression_kwargs=dict(S_w=4.0, mu_b=0.2, rho=0.4, a_0=3.0, b_0=2.5))
    m.add_data(Y)
    X = m.data_list[0][0]
    out = dict(N=N, B=B, L=L, T=T, basis=basis, Y=Y, X=X, A0=m.adjacency.copy(), W0=m.weights.copy(),
               b0=m.biases.copy(), eta0=np.array([r.eta for r in m.regressions]), ll0=m.log_likelihood(),
               means0=m.means[0], a_0=3.0, b_0=2.5)
    Js, hs, perms, us, zs, alphas, betas, etas = [], [], [], [], [], [], [], []
    for n, reg in enumerate(m.regressions):
        J, h = reg._lkhd_sufficient_statistics([(X, Y[:, n])])
        Js.append(J)
        hs.append(h)
        saved = ref_reg.sample_invgamma

        def rec_invgamma(alpha, beta):
            eta = saved(alpha, beta)
            alphas.append(alpha)
            betas.append(beta)
            etas.append(eta)
            return eta

        ref_reg.sample_invgamma = rec_invgamma
        try:
            with Recorder() as rec:
                reg.resample([(X, Y[:, n])])
        finally:
            ref_reg.sample_invgamma = saved
        perms.append(rec.perms[0])
        us.append(np.array(rec.us))
        z_full = np.zeros(N * B + 1)
        z_full[np.concatenate((np.repeat(reg.a, B), [1])).astype(bool)] = rec.zs[0]
        zs.append(z_full)
    r0 = m.regressions[0]
    out.update(J=np.array(Js), h=np.array(hs), perm=np.array(perms), us=np.array(us), z=np.array(zs),
               alpha=np.array(alphas), beta=np.array(betas), eta1=np.array(etas),
               A1=m.adjacency.copy(), W1=m.weights.copy(), b1=m.biases.copy(), ll1=m.log_likelihood(),
               means1=m.means[0], rho=r0.rho, mu_w=r0.mu_w, S_w=r0.S_w, mu_b=r0.mu_b, S_b=r0.S_b)
    return out


def basis_table():
    out = {}
    for (B, L) in [(1, 100), (2, 100), (3, 100), (3, 10), (5, 50)]:
        out["B%d_L%d" % (B, L)] = cosine_basis(B, L)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(OUT, "basis.npz"), **basis_table())
    np.savez_compressed(os.path.join(OUT, "kat_small.npz"), **kat_case(3, 2, 10, 50, True))
    np.savez_compressed(os.path.join(OUT, "kat_readme.npz"), **kat_case(4, 1, 100, 10000, True))
    np.savez_compressed(os.path.join(OUT, "kat_cfg2.npz"), **kat_case(27, 3, 100, 100000, False))
    np.savez_compressed(os.path.join(OUT, "reference_tests.npz"), **reference_tests())
    np.savez_compressed(os.path.join(OUT, "full_sweep.npz"), **full_sweep_case())
    np.savez_compressed(os.path.join(OUT, "generate.npz"), **generate_case())
    np.savez_compressed(os.path.join(OUT, "gaussian.npz"), **gaussian_case())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
    k = np.load(os.path.join(OUT, "kat_cfg2.npz"))
    for key in ["sumY", "sumX", "maxX", "sumpsi", "ll", "trJ", "froJ", "Jcorner", "J10", "sumh", "h0",
                "ml_a", "ml_ones"]:
        print(key, repr(float(k[key])))
