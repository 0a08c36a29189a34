"""Gaussian-observation siblings (regression.py:380-456, models.py:270-276) on the GPU: the same Gram
and spike-and-slab kernels with omega = 1/eta, kappa = y/eta, checked against fixtures produced by the reference's own
SparseGaussianGLM (oracle/gen_golden.py: gaussian_case) and against the oracle."""
import numpy as np
import pytest

from oracle import pyglm_oracle as O

pytestmark = pytest.mark.gpu


def _golden_model(g):
    from pyglm_b200.models import SparseGaussianGLM
    N = int(g["N"])
    np.random.seed(0)
    m = SparseGaussianGLM(N, basis=g["basis"], regression_kwargs=dict(S_w=4.0, mu_b=0.2, rho=0.4, a_0=3.0, b_0=2.5))
    for n, reg in enumerate(m.regressions):
        reg.a, reg.W, reg.b, reg.eta = g["A0"][n], g["W0"][n], g["b0"][n:n + 1], float(g["eta0"][n])
    m.add_data(g["Y"])
    return m


def test_gaussian_model_matches_reference(golden):
    g = golden("gaussian.npz")
    N, B, T = int(g["N"]), int(g["B"]), int(g["T"])
    m = _golden_model(g)
    np.testing.assert_allclose(m.data_list[0][0], g["X"], atol=1e-13)
    assert m.log_likelihood() == pytest.approx(float(g["ll0"]), rel=1e-11)
    np.testing.assert_allclose(m.means[0], g["means0"], rtol=1e-10, atol=1e-13)
    # one resample_regressions() on the reference's recorded draws
    m.engine.inject = dict(omega=None, perm=g["perm"], us=g["us"], z=g["z"])
    np.random.seed(1)
    m.resample_regressions()
    m.engine.inject = None
    assert np.array_equal(m.adjacency, g["A1"])
    np.testing.assert_allclose(m.weights, g["W1"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(m.biases, g["b1"], rtol=1e-8)
    # what _resample_eta hands to sample_invgamma (regression.py:432-445)
    rss = m.engine.residual_ss(m._device_datasets(), m.adjacency, m.weights, m.biases)
    np.testing.assert_allclose(float(g["b_0"]) + rss, g["beta"], rtol=1e-10)
    assert np.all(np.array([r.eta for r in m.regressions]) > 0)
    for n, reg in enumerate(m.regressions):
        reg.eta = float(g["eta1"][n])
    assert m.log_likelihood() == pytest.approx(float(g["ll1"]), rel=1e-10)
    m.resample_model()                                      # network step included


@pytest.mark.parametrize("N,B,L,T", [(3, 2, 10, 600), (20, 2, 30, 1500)])
def test_gaussian_generate_replays_the_reference_recursion(N, B, L, T):
    from pyglm_b200.models import SparseGaussianGLM
    from pyglm_b200.utils.basis import cosine_basis
    rng = np.random.default_rng(N)
    basis = cosine_basis(B, L=L) / L
    np.random.seed(2)
    m = SparseGaussianGLM(N, basis=basis, regression_kwargs=dict(S_w=1.0, mu_b=0.1, eta=0.4), seed=9)
    for n, reg in enumerate(m.regressions):
        reg.a = rng.random(N) < 0.4
        reg.W = reg.a[:, None] * rng.standard_normal((N, B)) * (0.3 / np.sqrt(N))
        reg.b = np.array([0.1 * rng.standard_normal()])
    X, Y, Z = m.generate(T=T, keep=True, return_uniforms=True)
    Xo, Yo = O.generate_gaussian(m.weights, m.biases, [r.eta for r in m.regressions], basis, T, Z)
    np.testing.assert_allclose(Y, Yo, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(X, Xo, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(X.reshape(T, -1), O.convolve_with_basis(Y, basis).reshape(T, -1), atol=1e-13)
    assert abs(Z.mean()) < 5 / np.sqrt(Z.size) and abs(Z.std() - 1) < 0.05
    assert np.isfinite(m.log_likelihood())


def test_gaussian_chain_recovers_noise_and_structure():
    """Statistical sanity on synthetic data: eta concentrates on the true noise variance, the strong self-connections
    are found, the log-likelihood rises."""
    from pyglm_b200.models import SparseGaussianGLM
    from pyglm_b200.utils.basis import cosine_basis
    N, B, L, T = 5, 2, 15, 20000
    basis = cosine_basis(B, L=L) / L
    np.random.seed(4)
    true = SparseGaussianGLM(N, basis=basis, regression_kwargs=dict(S_w=1.0, eta=0.25), seed=3)
    for n, reg in enumerate(true.regressions):
        reg.a[:] = False
        reg.W[:] = 0.0
        reg.a[n] = True
        reg.W[n, :] = 0.35
        reg.b[:] = 0.1
    _, Y = true.generate(T=T, keep=False)
    m = SparseGaussianGLM(N, basis=basis, regression_kwargs=dict(S_w=1.0, rho=0.3), seed=5)
    m.add_data(Y)
    ll = [m.log_likelihood()]
    etas, As = [], []
    for k in range(60):
        m.resample_model()
        ll.append(m.log_likelihood())
        if k >= 20:
            etas.append([r.eta for r in m.regressions])
            As.append(m.adjacency.astype(float))
    assert ll[-1] > ll[0]
    # the reference's beta omits the factor 1/2 (regression.py:443), so eta concentrates near 2 x the noise variance
    np.testing.assert_allclose(np.mean(etas, 0), 2 * 0.25, rtol=0.1)
    P = np.mean(As, 0)
    assert np.all(np.diag(P) > 0.9) and P[~np.eye(N, dtype=bool)].mean() < 0.3


def test_standalone_gaussian_regression():
    from pyglm_b200.regression import SparseGaussianRegression, GaussianRegression
    rng = np.random.default_rng(0)
    N, B, T = 4, 2, 2000
    X = rng.standard_normal((T, N * B))
    w = np.zeros(N * B)
    w[:B] = [1.0, -0.5]
    y = X @ w + 0.3 + 0.2 * rng.standard_normal(T)
    np.random.seed(1)
    reg = SparseGaussianRegression(N, B, S_w=4.0, rho=0.5)
    for _ in range(15):
        reg.resample([(X, y)])
    assert reg.a[0] and np.allclose(reg.W[0], [1.0, -0.5], atol=0.05) and abs(reg.b[0] - 0.3) < 0.05
    ll = reg.log_likelihood((X, y))
    np.testing.assert_allclose(ll, O.gaussian_log_likelihood_terms(X, y, reg.a, reg.W, reg.b, reg.eta), rtol=1e-9)
    assert GaussianRegression(N, B).rho.min() == 1.0
