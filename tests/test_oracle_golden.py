"""Pins the numpy oracle (oracle/pyglm_oracle.py) against fixtures produced by the reference's
own files (oracle/gen_golden.py -> tests/golden/).  CPU only."""
import numpy as np
import pytest

from oracle import pyglm_oracle as O

RTOL = 1e-12


def test_cosine_basis_matches_reference(golden):
    g = golden("basis.npz")
    for key in g.files:
        B, L = (int(s[1:]) for s in key.split("_"))
        np.testing.assert_allclose(O.cosine_basis(B, L), g[key], rtol=1e-14, atol=1e-16)


def test_basis_facts_survey_appendix_d():
    b1 = O.cosine_basis(1, 100) / 100
    assert b1.max() == pytest.approx(0.03799078716012692, rel=1e-13)
    assert b1[50, 0] == pytest.approx(0.00145808636170209, rel=1e-12)
    b3 = O.cosine_basis(3, 100) / 100
    assert list(b3.argmax(0)) == [0, 25, 50]
    np.testing.assert_allclose(b3.sum(0), 1.0, atol=1e-14)


@pytest.mark.parametrize("name", ["kat_small.npz", "kat_readme.npz"])
def test_kat_arrays(golden, name):
    g = golden(name)
    N, B = int(g["N"]), int(g["B"])
    Y = g["Y"].astype(float)
    X = O.convolve_with_basis(Y, g["basis"])
    np.testing.assert_allclose(X, g["X"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(O.convolve_direct(Y, g["basis"]), g["X"], rtol=0, atol=5e-16)
    a, W, b = g["a"], g["W"], g["b"]
    y = Y[:, N - 1]
    psi = O.activation(X, a, W, b)
    np.testing.assert_allclose(psi, g["psi"], rtol=RTOL)
    np.testing.assert_allclose(O.log_likelihood_terms(X, y, a, W, b), g["ll_terms"], rtol=RTOL)
    np.testing.assert_allclose(O.mean(X, a, W, b), g["mean"], rtol=RTOL)
    np.testing.assert_allclose(O.pg1_mean(psi), g["omega"], rtol=1e-12)
    J, h = O.lkhd_sufficient_statistics(O.flatten_X(X, N, B), g["omega"], O.kappa(y))
    np.testing.assert_allclose(J, g["J_lkhd"], rtol=RTOL)
    np.testing.assert_allclose(h, g["h_lkhd"], rtol=RTOL)
    hyper = dict(mu_w=O.expand_scalar(0.0, (N, B)), S_w=O.expand_cov(10.0, (N, B, B)),
                 mu_b=O.expand_scalar(-2.0, (1,)), S_b=O.expand_cov(1.0, (1, 1)))
    J0, h0 = O.prior_sufficient_statistics(**hyper)
    np.testing.assert_allclose(J0, g["J_prior"], rtol=RTOL)
    np.testing.assert_allclose(h0, g["h_prior"], rtol=RTOL)
    assert O.marginal_likelihood(J0, h0, J0 + J, h0 + h, a, B) == pytest.approx(float(g["ml_a"]), rel=1e-13)
    assert O.marginal_likelihood(J0, h0, J0 + J, h0 + h, np.ones(N, bool), B) == \
        pytest.approx(float(g["ml_ones"]), rel=1e-13)
    # a-scan with the recorded permutation / uniforms, then the W draw with the recorded normals
    trace = []
    a1 = O.collapsed_resample_a(J0, h0, J0 + J, h0 + h, g["scan_a0"], 0.5 * np.ones(N), B,
                                g["scan_perm"], g["scan_us"], trace=trace)
    assert np.array_equal(a1, g["scan_a"])
    np.testing.assert_allclose(np.array([[t[1], t[2]] for t in trace]), g["scan_lps"], rtol=1e-12)
    m = O._mask(a1, B)
    W1, b1 = O.resample_W(J0 + J, h0 + h, a1, B, g["draw_z"][m])
    np.testing.assert_allclose(W1, g["draw_W"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(b1, g["draw_b"], rtol=1e-11)


def test_kat_cfg2_scalars(golden):
    """SURVEY Appendix D, N=27 B=3 L=100 T=1e5 (BASELINE configs[1] shape): scalars only."""
    g = golden("kat_cfg2.npz")
    N, B, L, T = (int(g[k]) for k in "NBLT")
    Y = (np.random.default_rng(0).random((T, N)) < 0.05).astype(float)
    assert Y.sum() == float(g["sumY"]) == 134884
    X = O.convolve_with_basis(Y, O.cosine_basis(B, L) / L)
    assert X.sum() == pytest.approx(404513.48560710484, rel=1e-13)
    assert X.max() == pytest.approx(0.30812764546925336, rel=1e-13)
    np.testing.assert_allclose(X[g["X_rows"]], g["X_sample"], rtol=0, atol=1e-15)
    a, W, b = g["a"], g["W"], g["b"]
    y = Y[:, N - 1]
    psi = O.activation(X, a, W, b)
    assert psi.sum() == pytest.approx(-11712.261509123337, rel=1e-12)
    assert O.log_likelihood_terms(X, y, a, W, b).sum() == pytest.approx(-64989.20242030339, rel=1e-13)
    J, h = O.lkhd_sufficient_statistics(O.flatten_X(X, N, B), O.pg1_mean(psi), O.kappa(y))
    assert np.trace(J) == pytest.approx(31760.744213236714, rel=1e-12)
    assert np.linalg.norm(J) == pytest.approx(29882.190230115895, rel=1e-12)
    assert J[-1, -1] == pytest.approx(24848.887086551375, rel=1e-12)
    assert J[1, 0] == pytest.approx(80.01703136862069, rel=1e-12)
    assert h.sum() == pytest.approx(-226947.5115201395, rel=1e-12)
    assert h[0] == pytest.approx(-2261.687295282024, rel=1e-12)
    hyper = dict(mu_w=O.expand_scalar(0.0, (N, B)), S_w=O.expand_cov(10.0, (N, B, B)),
                 mu_b=O.expand_scalar(-2.0, (1,)), S_b=O.expand_cov(1.0, (1, 1)))
    J0, h0 = O.prior_sufficient_statistics(**hyper)
    assert O.marginal_likelihood(J0, h0, J0 + J, h0 + h, a, B) == pytest.approx(40521.302065428004, rel=1e-12)
    assert O.marginal_likelihood(J0, h0, J0 + J, h0 + h, np.ones(N, bool), B) == \
        pytest.approx(40514.35607382245, rel=1e-12)


def test_reference_test_suite_assertions(golden):
    """test/test_generate.py:24 (filter == explicit causal dot product) and :55 (lag law)."""
    g = golden("reference_tests.npz")
    Y = g["tm_Y"].astype(float)
    Xc = O.convolve_with_basis(Y, g["tm_basis"])
    assert np.allclose(g["tm_X_generate"], Xc)
    np.testing.assert_allclose(Xc, g["tm_X_conv"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(O.model_means(Xc, g["tm_A"], g["tm_W"], g["tm_b"]), g["tm_means"], rtol=1e-12)
    assert O.model_log_likelihood(Xc, Y, g["tm_A"], g["tm_W"], g["tm_b"]) == \
        pytest.approx(float(g["tm_ll"]), rel=1e-13)
    Y = g["tb_Y"].astype(float)
    X = O.convolve_with_basis(Y, np.eye(3))
    for n in range(2):
        for b in range(3):
            assert np.allclose(Y[:-(b + 1), n], X[(b + 1):, n, b])
    np.testing.assert_allclose(X, g["tb_X"], atol=1e-15)


def test_full_sweep_with_injected_randomness(golden):
    """One resample_model() of the reference (models.py:166-171) neuron by neuron."""
    g = golden("full_sweep.npz")
    N, B = int(g["N"]), int(g["B"])
    Y = g["Y"].astype(float)
    X = O.convolve_with_basis(Y, g["basis"])
    A, W, bias = g["A0"].copy(), g["W0"].copy(), g["b0"].copy()
    assert O.model_log_likelihood(X, Y, A, W, bias) == pytest.approx(float(g["ll0"]), rel=1e-13)
    np.testing.assert_allclose(O.model_means(X, A, W, bias), g["means0"], rtol=1e-12)
    hyper = dict(rho=g["rho"], mu_w=g["mu_w"], S_w=g["S_w"], mu_b=g["mu_b"], S_b=g["S_b"])
    for n in range(N):
        a, Wn, bn = O.resample_regression(X, Y[:, n], A[n], W[n], bias[n:n + 1], hyper,
                                          g["omega"][:, n], g["perm"][n], g["us"][n], g["z"][n])
        A[n], W[n], bias[n] = a, Wn, bn[0]
    assert np.array_equal(A, g["A1"])
    np.testing.assert_allclose(W, g["W1"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(bias, g["b1"], rtol=1e-10)
    assert O.model_log_likelihood(X, Y, A, W, bias) == pytest.approx(float(g["ll1"]), rel=1e-11)


def test_generate_matches_reference(golden):
    """oracle.generate against the reference's own generate() (models.py:98-151) on the uniforms it consumed."""
    g = golden("generate.npz")
    X, Y = O.generate(g["weights"], g["biases"], g["basis"], g["U"].shape[0], g["U"])
    assert np.array_equal(Y, g["Y"])
    np.testing.assert_allclose(X, g["X"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(X.reshape(X.shape[0], -1),
                               O.convolve_with_basis(Y, g["basis"]).reshape(X.shape[0], -1), atol=1e-14)


def test_gaussian_regression_matches_reference(golden):
    """SparseGaussianGLM (regression.py:380-446): sufficient statistics, log-likelihood, means, the reference's own
    resample() on recorded draws, and the (alpha, beta) it hands to sample_invgamma."""
    g = golden("gaussian.npz")
    N, B, T = int(g["N"]), int(g["B"]), int(g["T"])
    X, Y = g["X"].reshape(T, -1), g["Y"]
    hyper = dict(rho=g["rho"], mu_w=g["mu_w"], S_w=g["S_w"], mu_b=g["mu_b"], S_b=g["S_b"])
    ll0 = ll1 = 0.0
    for n in range(N):
        eta = g["eta0"][n]
        J, h = O.lkhd_sufficient_statistics(X, O.gaussian_omega(T, eta), O.gaussian_kappa(Y[:, n], eta))
        np.testing.assert_allclose(J, g["J"][n], rtol=1e-12)
        np.testing.assert_allclose(h, g["h"][n], rtol=1e-12)
        ll0 += O.gaussian_log_likelihood_terms(X, Y[:, n], g["A0"][n], g["W0"][n], g["b0"][n:n + 1], eta).sum()
        np.testing.assert_allclose(O.activation(X, g["A0"][n], g["W0"][n], g["b0"][n:n + 1]), g["means0"][:, n],
                                   rtol=1e-12, atol=1e-14)
        a, W, b = O.resample_regression(X, Y[:, n], g["A0"][n], g["W0"][n], g["b0"][n:n + 1], hyper,
                                        O.gaussian_omega(T, eta), g["perm"][n], g["us"][n], g["z"][n],
                                        kap=O.gaussian_kappa(Y[:, n], eta))
        assert np.array_equal(a, g["A1"][n])
        np.testing.assert_allclose(W, g["W1"][n], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(b[0], g["b1"][n], rtol=1e-9)
        alpha, beta = O.gaussian_eta_posterior(X, Y[:, n], g["A1"][n], g["W1"][n], g["b1"][n:n + 1],
                                               float(g["a_0"]), float(g["b_0"]))
        assert alpha == g["alpha"][n]
        np.testing.assert_allclose(beta, g["beta"][n], rtol=1e-12)
        ll1 += O.gaussian_log_likelihood_terms(X, Y[:, n], g["A1"][n], g["W1"][n], g["b1"][n:n + 1],
                                               g["eta1"][n]).sum()
    np.testing.assert_allclose(ll0, float(g["ll0"]), rtol=1e-12)
    np.testing.assert_allclose(ll1, float(g["ll1"]), rtol=1e-12)


def test_oracle_chain_matches_reference_chains(golden):
    """The numpy port's own Gibbs chain on the README configuration (BASELINE configs[0]) against six chains of the
    reference's unmodified sampler on the same data (tests/golden/chains_cfg1.npz): KS on the post-burn-in
    log-likelihood, P(A), E[a W], E[b], every bound derived from the spread between the reference chains."""
    from tests.chain_stats import load_case, summarize, assert_chain_matches_reference
    g, Y, (T, N, B, L) = load_case(golden, "chains_cfg1.npz")
    sweeps, burn = int(g["sweeps"]), int(g["burn"])
    m = O.OracleSparseBernoulliGLM(N, g["basis"], S_w=10.0, mu_b=-2.0, seed=21, pg_threads=1)
    m.add_data(Y)
    rec = ([], [], [], [])
    for _ in range(sweeps):
        m.resample_model()
        for r, v in zip(rec, (m.log_likelihood(), m.A, m.W, m.bias)):
            r.append(np.array(v))
    ll, PA, EW, Eb = summarize(*rec, burn=burn)
    assert_chain_matches_reference(g, ll, PA, EW, Eb)
