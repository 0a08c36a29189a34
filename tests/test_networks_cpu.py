"""Host network step (models.py:228-236, networks.py:76-190) on CPU: NIW posterior against the oracle, the layout of
the broadcast hyper-parameters, and which weights feed which Gaussian."""
import numpy as np

from oracle import pyglm_oracle as O
from pyglm_b200.networks import NIWDenseNetwork, NIWSparseNetwork


def _problem(N=12, B=2, seed=0):
    rng = np.random.default_rng(seed)
    return rng.random((N, N)) < 0.5, rng.standard_normal((N, N, B))


def test_niw_posterior_matches_oracle_and_uses_the_masked_weights():
    N, B = 12, 2
    A, W = _problem(N, B)
    np.random.seed(3)
    net = NIWSparseNetwork(N, B)
    eye = np.eye(N, dtype=bool)
    seen = {}
    for name, g in (("off", net._gaussian), ("self", net._self_gaussian)):
        orig = g._posterior
        g._posterior = (lambda data, _o=orig, _n=name: seen.__setitem__(_n, np.array(data)) or _o(data))
    net.resample((A, W))
    assert np.array_equal(seen["off"], W[~eye & A])          # networks.py:137-141, row-major order
    assert np.array_equal(seen["self"], W[eye & A])          # networks.py:144-145
    g = net._gaussian
    mu_n, sig_n, k_n, nu_n = O.niw_posterior(W[~eye & A], g.mu_0, g.sigma_0, g.kappa_0, g.nu_0)
    p = type(g)._posterior(g, W[~eye & A])
    np.testing.assert_allclose(p[0], mu_n, rtol=1e-13)
    np.testing.assert_allclose(p[1], sig_n, rtol=1e-13)
    assert p[2] == k_n and p[3] == nu_n


def test_broadcast_hyperparameters_layout():
    N, B = 7, 3
    np.random.seed(1)
    net = NIWSparseNetwork(N, B, rho=0.3, rho_self=0.9)
    S, M, R = net.sigma_W, net.mu_W, net.rho
    assert S.shape == (N, N, B, B) and M.shape == (N, N, B) and R.shape == (N, N)
    for i in range(N):
        for j in range(N):
            g = net._self_gaussian if i == j else net._gaussian
            assert np.array_equal(S[i, j], g.sigma) and np.array_equal(M[i, j], g.mu)
            assert R[i, j] == (0.9 if i == j else 0.3)
    dense = NIWDenseNetwork(N, B)
    assert np.array_equal(dense.sigma_W[2, 3], dense._gaussian.sigma) and np.all(dense.rho == 1.0)


def test_state_round_trip():
    N, B = 5, 2
    A, W = _problem(N, B, seed=2)
    np.random.seed(0)
    a, b = NIWSparseNetwork(N, B), NIWSparseNetwork(N, B)
    a.resample((A, W))
    b.set_state(a.get_state())
    assert np.array_equal(a.sigma_W, b.sigma_W) and np.array_equal(a.mu_W, b.mu_W)
