"""Learned adjacency priors (Beta-Bernoulli, stochastic block, latent distance).  The reference
snapshot has none of them running (networks.py:175,214,261), so there is no reference output to pin against: the
checks are the conjugate posteriors, brute-force conditionals on small graphs, posterior invariance of the moves
(prior-recovery test) and recovery of planted structure."""
import itertools

import numpy as np
from scipy import stats

from pyglm_b200.networks import (BetaBernoulli, NIWBetaBernoulliNetwork, NIWStochasticBlockNetwork,
                                 NIWLatentDistanceNetwork, elliptical_slice)


def _weights(N, B, seed=0):
    return np.random.default_rng(seed).standard_normal((N, N, B))


def test_beta_bernoulli_conjugate_posterior():
    np.random.seed(0)
    bb = BetaBernoulli(2.0, 3.0)
    bits = np.random.rand(50) < 0.3
    k = int(bits.sum())
    assert bb.posterior(bits) == (2.0 + k, 3.0 + 50 - k)
    draws = np.array([bb.resample(bits).rho for _ in range(4000)])
    assert stats.kstest(draws, stats.beta(2.0 + k, 3.0 + 50 - k).cdf).pvalue > 1e-3


def test_beta_bernoulli_network_uses_diagonal_and_offdiagonal_separately():
    N, B = 30, 2
    np.random.seed(1)
    A = np.random.rand(N, N) < 0.1
    A[np.diag_indices(N)] = True
    net = NIWBetaBernoulliNetwork(N, B)
    rho_off, rho_self = [], []
    for _ in range(300):
        net.resample((A, _weights(N, B)))
        R = net.rho
        assert R.shape == (N, N) and np.all(R[~np.eye(N, dtype=bool)] == R[0, 1]) and np.all(R.diagonal() == R[0, 0])
        rho_off.append(R[0, 1])
        rho_self.append(R[0, 0])
    k = int(A[~np.eye(N, dtype=bool)].sum())
    n = N * (N - 1)
    assert abs(np.mean(rho_off) - (1.0 + k) / (2.0 + n)) < 0.005
    assert abs(np.mean(rho_self) - (1.0 + N) / (2.0 + N)) < 0.01
    # keyword arguments reach the weight mixin, hyper-parameters keep the reference's layout
    net2 = NIWBetaBernoulliNetwork(N, B, a_0=5.0, is_diagonal_weight_special=False)
    assert net2._betabernoulli.a_0 == 5.0 and not net2.is_diagonal_weight_special
    assert net2.mu_W.shape == (N, N, B) and net2.sigma_W.shape == (N, N, B, B)
    shared = NIWBetaBernoulliNetwork(N, B, is_diagonal_conn_special=False)
    shared.resample((A, _weights(N, B)))
    assert np.all(shared.rho == shared.rho[0, 0])


def _sbm_log_joint(net, A, z):
    """log p(A off-diagonal, z | p, pi), from the definition."""
    N = net.N
    lj = np.log(net.pi)[z].sum()
    for i, j in itertools.product(range(N), range(N)):
        if i != j:
            pr = net.p[z[i], z[j]]
            lj += np.log(pr) if A[i, j] else np.log1p(-pr)
    return lj


def test_sbm_block_scores_equal_the_joint_conditional():
    N, B, C = 9, 1, 3
    np.random.seed(2)
    net = NIWStochasticBlockNetwork(N, B, C=C)
    A = np.random.rand(N, N) < 0.4
    for n in (0, 4, 8):
        s = net.block_scores(A, n)
        ref = np.empty(C)
        for c in range(C):
            z = net.z.copy()
            z[n] = c
            ref[c] = _sbm_log_joint(net, A, z)
        np.testing.assert_allclose(s - s[0], ref - ref[0], rtol=0, atol=1e-10)


def test_sbm_recovers_planted_blocks():
    N, B, C = 60, 1, 2
    rng = np.random.default_rng(3)
    z_true = np.repeat([0, 1], N // 2)
    p_true = np.array([[0.6, 0.05], [0.1, 0.5]])
    A = rng.random((N, N)) < p_true[np.ix_(z_true, z_true)]
    np.random.seed(3)
    net = NIWStochasticBlockNetwork(N, B, C=C)
    W = _weights(N, B)
    rho_mean = np.zeros((N, N))
    for it in range(60):
        net.resample((A, W))
        if it >= 30:
            rho_mean += net.rho / 30
    agree = np.mean(net.z == z_true)
    assert max(agree, 1.0 - agree) >= 0.95
    off = ~np.eye(N, dtype=bool)
    assert np.abs(rho_mean - p_true[np.ix_(z_true, z_true)])[off].max() < 0.1
    state = net.get_state()
    other = NIWStochasticBlockNetwork(N, B, C=C)
    other.set_state(state)
    assert np.array_equal(other.rho, net.rho) and np.array_equal(other.mu_W, net.mu_W)


def test_elliptical_slice_leaves_the_posterior_invariant():
    # Gaussian prior N(0, 2^2) x Gaussian likelihood N(1.5 | f, 1) -> posterior N(1.2, 0.8)
    np.random.seed(4)
    f, out = np.array(0.0), []
    for _ in range(6000):
        f, _ = elliptical_slice(f, lambda x: -0.5 * (1.5 - float(x)) ** 2, 2.0)
        out.append(float(f))
    out = np.array(out[200:])
    assert abs(out.mean() - 1.2) < 0.05 and abs(out.var() - 0.8) < 0.08


def test_latent_distance_likelihood_and_prior_recovery():
    N, B = 8, 1
    np.random.seed(5)
    net = NIWLatentDistanceNetwork(N, B, dim=2)
    A = np.random.rand(N, N) < 0.3
    R = net.rho
    ref = sum(np.log(R[i, j]) if A[i, j] else np.log1p(-R[i, j])
              for i in range(N) for j in range(N) if i != j)
    np.testing.assert_allclose(net.log_likelihood_adjacency(A), ref, rtol=1e-10)
    d01 = np.sum((net.L[0] - net.L[1]) ** 2)
    np.testing.assert_allclose(R[0, 1], 1.0 / (1.0 + np.exp(d01 - net.gamma)), rtol=1e-12)
    # the per-neuron score used by the slice move differs from the full likelihood by a constant in l_n
    links = A.astype(np.float64) + A.T
    keep, cand = net.L[3].copy(), np.random.randn(4, 2)
    full = []
    for l in cand:
        net.L[3] = l
        full.append(net.log_likelihood_adjacency(A))
    net.L[3] = keep
    part = [net.location_score(links, 3, l) for l in cand]
    np.testing.assert_allclose(np.diff(part), np.diff(full), rtol=0, atol=1e-10)

    # successive-conditional test: alternate A | (L, gamma) from the model with one resample of (L, gamma) | A; if
    # the moves leave the posterior invariant, (L, gamma) keep their prior marginals.
    N = 6
    net = NIWLatentDistanceNetwork(N, B, dim=2, sigma_l=1.0, mu_gamma=0.5, sigma_gamma=1.0)
    W = _weights(N, B)
    gam, l00 = [], []
    for it in range(1500):
        A = np.random.rand(N, N) < net.rho
        net.resample((A, W))
        gam.append(net.gamma)
        l00.append(net.L[0, 0])
    gam, l00 = np.array(gam[::5]), np.array(l00[::5])
    assert abs(gam.mean() - 0.5) < 0.25 and abs(gam.std() - 1.0) < 0.25
    assert abs(l00.mean()) < 0.25 and abs(l00.std() - 1.0) < 0.25


def test_latent_distance_recovers_planted_geometry():
    N, B = 50, 1
    rng = np.random.default_rng(6)
    L_true = rng.standard_normal((N, 2)) * 1.5
    d = ((L_true[:, None] - L_true[None]) ** 2).sum(-1)
    rho_true = 1.0 / (1.0 + np.exp(d - 1.0))
    A = rng.random((N, N)) < rho_true
    np.random.seed(6)
    net = NIWLatentDistanceNetwork(N, B, dim=2, sigma_l=1.5, mu_gamma=0.0, sigma_gamma=2.0)
    W = _weights(N, B)
    ll0 = net.log_likelihood_adjacency(A)
    rho_mean = np.zeros((N, N))
    for it in range(150):
        net.resample((A, W))
        if it >= 75:
            rho_mean += net.rho / 75
    off = ~np.eye(N, dtype=bool)
    assert net.log_likelihood_adjacency(A) > ll0 + 100
    assert np.corrcoef(rho_mean[off], rho_true[off])[0, 1] > 0.8
    other = NIWLatentDistanceNetwork(N, B, dim=2)
    other.set_state(net.get_state())
    assert np.array_equal(other.rho, net.rho)


def test_fixed_weight_combinations():
    from pyglm_b200 import networks
    N, B = 6, 2
    np.random.seed(7)
    A = np.random.rand(N, N) < 0.5
    for name, kw in (("FixedMeanBetaBernoulliNetwork", {}), ("FixedMeanStochasticBlockNetwork", dict(C=3)),
                     ("FixedMeanLatentDistanceNetwork", dict(dim=3))):
        net = getattr(networks, name)(N, B, mu=0.5, sigma=2.0, mu_self=-1.0, sigma_self=0.5, **kw)
        net.resample((A, _weights(N, B)))
        assert np.all(net.mu_W[0, 1] == 0.5) and np.all(net.mu_W[2, 2] == -1.0)
        assert np.array_equal(net.sigma_W[0, 1], 2.0 * np.eye(B)) and np.array_equal(net.sigma_W[3, 3], 0.5 * np.eye(B))
        assert net.rho.shape == (N, N) and np.all((net.rho > 0) & (net.rho < 1))
