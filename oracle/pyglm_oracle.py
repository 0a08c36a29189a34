"""ORACLE -- test infrastructure, NOT product code.

CPU restatement (numpy/scipy, float64) of the reference's sparse-Bernoulli Gibbs hot path,
function by function, each citing the reference file:line it follows (paths relative to
/root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; nothing under pyglm_b200/ does.

Pinning: every deterministic function here is checked against fixtures produced by the
reference's OWN files run through oracle/ref_shim (oracle/gen_golden.py -> tests/golden/*.npz,
tests/test_oracle_golden.py), including the two assertions of the reference's test suite
(test/test_generate.py:24 and :55) and the SURVEY Appendix-D known-answer values.
The stochastic primitives (PG draws, Gaussian / categorical / NIW draws) have no golden vector
in the reference and their third-party implementations are absent: PARITY UNPINNED for the
draws themselves; their deterministic cores (Cholesky, solves, logsumexp) are pinned via
marginal_likelihood / resample_W with injected randomness.

All random inputs are injectable (omega, perm, u, z) so that the CUDA path can be compared on
identical draws.
"""
import ctypes
import os

import numpy as np
import scipy.linalg
import scipy.signal as sig
from scipy.linalg import block_diag
from scipy.linalg.lapack import dpotrs
from scipy.special import logsumexp

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- PG sampler (C)
_pg_lib = None


def pg_lib():
    """ctypes handle on oracle/libpg_oracle.so (built by oracle/Makefile / __graft_entry__.build)."""
    global _pg_lib
    if _pg_lib is None:
        path = os.path.join(_HERE, "libpg_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/libpg_oracle.so missing: run `make -C oracle`")
        lib = ctypes.CDLL(path)
        lib.pg1_drawv.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint32,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.pg_omp_max_threads.restype = ctypes.c_int
        lib.philox_stream_unif.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64,
                                           ctypes.c_int, ctypes.c_void_p]
        lib.philox4x32_10.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _pg_lib = lib
    return _pg_lib


def pg1_draw(psi, seed, call_id=0, rng_kind=0, nthreads=0):
    """omega_i ~ PG(1, psi_i).  Stands for pgdrawvpar(ppgs, ones, psi, out), regression.py:501-508.
    rng_kind 0: per-element Philox stream (the stream the CUDA kernel uses); 1: per-thread RNGs."""
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    out = np.empty_like(psi)
    pg_lib().pg1_drawv(psi.ctypes.data, psi.size, int(seed), int(call_id), int(rng_kind),
                       int(nthreads), out.ctypes.data)
    return out


def philox_uniforms(seed, call_id, elem, count):
    out = np.empty(count)
    pg_lib().philox_stream_unif(int(seed), int(call_id), int(elem), int(count), out.ctypes.data)
    return out


def pg1_mean(psi):
    """E[PG(1, psi)] = tanh(psi/2) / (2 psi), 1/4 at 0."""
    psi = np.asarray(psi, dtype=np.float64)
    safe = np.where(np.abs(psi) < 1e-6, 1.0, psi)
    return np.where(np.abs(psi) < 1e-6, 0.25 - psi ** 2 / 48.0, np.tanh(safe / 2.0) / (2.0 * safe))


def pg1_var(psi):
    """Var[PG(1, psi)] = (sinh psi - psi) / (4 psi^3 cosh^2(psi/2)), 1/24 at 0."""
    psi = np.asarray(psi, dtype=np.float64)
    safe = np.where(np.abs(psi) < 1e-3, 1.0, psi)
    v = (np.sinh(safe) - safe) / (4.0 * safe ** 3 * np.cosh(safe / 2.0) ** 2)
    return np.where(np.abs(psi) < 1e-3, 1.0 / 24.0, v)


def jstar_cdf(x, z, form="auto", terms=200):
    """P(J*(1, z) <= x), the law Devroye's sampler draws from (PG(1, psi) = J*(1, |psi|/2) / 4; Polson, Scott &
    Windle 2013, eqs. (12)-(14)).  Independent of the sampler: the density is cosh(z) exp(-x z^2/2) sum_n (-1)^n a_n(x)
    and each piece of the alternating series integrates in closed form.
      left  (fast for small x): a_n(x) = 2 c/sqrt(2 pi x^3) exp(-c^2/(2x)), c = 2n+1, i.e. twice the first-passage
            density of level c; with the tilt exp(-z^2 x/2) the integral is exp(-cz) P(tau_c <= x) for a Brownian
            motion with drift z:   2 cosh(z) sum (-1)^n [ e^{-cz} Phi((zx-c)/sqrt x) + e^{cz} Phi(-(zx+c)/sqrt x) ]
      right (fast for large x): a_n(x) = pi(n+1/2) exp(-(n+1/2)^2 pi^2 x/2); with lam_n = z^2/2 + (n+1/2)^2 pi^2/2
            and sum (-1)^n pi(n+1/2)/lam_n = 1/cosh(z):   1 - cosh(z) sum (-1)^n pi(n+1/2) exp(-lam_n x)/lam_n
    form="auto" switches at the sampler's truncation point 0.64; tests check that the two forms agree there."""
    from scipy.special import log_ndtr
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    z = abs(float(z))
    n = np.arange(terms, dtype=np.float64)[:, None]
    sgn = np.where(n % 2 == 0, 1.0, -1.0)
    xs = np.maximum(x[None, :], 1e-300)
    out = np.empty_like(x)
    left = (x <= 0.64) if form == "auto" else np.full(x.shape, form == "left")
    if left.any():
        c = 2.0 * n + 1.0
        xl = xs[:, left]
        lcz = np.log(np.cosh(z)) if z < 300 else z - np.log(2.0)
        t1 = np.exp(np.minimum(lcz - c * z + log_ndtr((z * xl - c) / np.sqrt(xl)), 700.0))
        t2 = np.exp(np.minimum(lcz + c * z + log_ndtr(-(z * xl + c) / np.sqrt(xl)), 700.0))
        out[left] = 2.0 * np.sum(sgn * (t1 + t2), axis=0)
    if (~left).any():
        lam = 0.5 * z * z + 0.5 * (n + 0.5) ** 2 * np.pi ** 2
        xr = xs[:, ~left]
        lcz = np.log(np.cosh(z)) if z < 300 else z - np.log(2.0)
        # cosh(z) exp(-lam x) stays finite: lam x >= z^2 x / 2 and x > 0.64 only matters while z is moderate
        out[~left] = 1.0 - np.sum(sgn * np.pi * (n + 0.5) * np.exp(lcz - lam * xr) / lam, axis=0)
    return np.clip(np.where(x <= 0, 0.0, out), 0.0, 1.0)


def pg1_cdf(w, psi):
    """P(PG(1, psi) <= w) for scalar psi (regression.py:496-508 draws omega_t from this law)."""
    return jstar_cdf(4.0 * np.asarray(w, dtype=np.float64), 0.5 * abs(float(psi)))


# --------------------------------------------------------------------------- (1) filter build
def cosine_basis(B, L=100, orth=False, norm=True, n_eye=0, a=1.0 / 120, b=0.5):
    """Raised-cosine bumps on a log-time axis.  Follows pyglm/utils/basis.py:61-106."""
    n_cos = B - n_eye
    assert n_cos >= 0 and n_eye >= 0
    basis = np.zeros((L, B))
    basis[:n_eye, :n_eye] = np.eye(n_eye)
    u_ir = np.log(a * np.arange(L) + b)                                   # :82-83
    ctrs = u_ir[np.floor(np.linspace(n_eye, (L / 2.0), n_cos)).astype(int)]  # :84
    if len(ctrs) == 1:
        w = ctrs / 2                                                      # :85-86
    else:
        w = (ctrs[-1] - ctrs[0]) / (n_cos - 1)                            # :88
    for i in range(n_cos):                                                # :91-93
        arg = np.maximum(-np.pi, np.minimum(np.pi, (u_ir - ctrs[i]) * np.pi / w / 2.0))
        basis[:, n_eye + i] = (np.cos(arg) + 1) / 2.0
    if orth:
        basis = scipy.linalg.orth(basis)
    elif norm:
        if np.any(basis < 0):
            raise Exception("We can only normalize nonnegative impulse responses!")
        basis = basis / np.tile(np.sum(basis, axis=0), [L, 1]) / (1.0 / L)  # :105
    return basis


def convolve_with_basis(S, basis):
    """X[t,n,b] = sum_{l=1..L} basis[l-1,b] S[t-l,n].  Follows pyglm/utils/basis.py:5-34
    (zero row prepended :18, per-b FFT convolution truncated to T :24-27, clip :30-32)."""
    T, N = S.shape
    R, B = basis.shape
    basis = np.vstack((np.zeros((1, B)), basis))
    F = np.empty((T, N, B))
    for b in range(B):
        F[:, :, b] = sig.fftconvolve(S, np.reshape(basis[:, b], (R + 1, 1)), 'full')[:T, :]
    if np.amin(basis) >= 0 and np.amin(S) >= 0:
        np.clip(F, 0, np.inf, out=F)
    return F


def convolve_direct(S, basis):
    """The same filter as an explicit causal sum (what test/test_generate.py:24 pins the FFT
    route against: generate() builds X[t] = Y[t-L:t].T.dot(flipud(basis)), models.py:121,140)."""
    T, N = S.shape
    L, B = basis.shape
    X = np.zeros((T, N, B))
    for l in range(1, L + 1):
        if l >= T:
            break
        X[l:] += S[:T - l, :, None] * basis[l - 1][None, None, :]
    return X


# --------------------------------------------------------------------------- host helpers
def expand_scalar(x, shp):
    """pyglm/utils/utils.py:7-12."""
    if np.isscalar(x):
        x = x * np.ones(shp)
    else:
        assert x.shape == shp
    return x


def expand_cov(c, shp):
    """pyglm/utils/utils.py:15-27."""
    d = shp[-1]
    if np.isscalar(c):
        c = c * np.eye(d)
        tshp = np.array(shp)
        tshp[-2:] = 1
        c = np.tile(c, tshp)
    else:
        assert c.shape == shp
    return c


def logistic(x):
    """pyglm/utils/utils.py:3-4."""
    return 1. / (1 + np.exp(-x))


# --------------------------------------------------------------------------- (5) psi, LL, mean
def flatten_X(X, N, B):
    """regression.py:173-180: (T,N,B) -> (T, N*B), n-major / b-minor."""
    if X.ndim == 2:
        assert X.shape[1] == N * B
        return X
    return np.reshape(X, (-1, N * B))


def activation(X, a, W, b):
    """psi = X . vec(a o W) + b.  regression.py:195-201."""
    N, B = W.shape
    X = flatten_X(X, N, B)
    w = np.reshape(a[:, None] * W, (N * B,))
    return X.dot(w) + np.ravel(b)[0]


def log_likelihood_terms(X, y, a, W, b):
    """Per-bin Bernoulli log-likelihood y psi - log(1+e^psi).  regression.py:491-494 with
    a_func=y (:515), b_func=1 (:518), c_func=1 (:521)."""
    psi = activation(X, a, W, b)
    return np.log(1.0) + y * psi - 1.0 * np.log1p(np.exp(psi))


def model_log_likelihood(X, Y, A, W, bias):
    """models.py:82-96: sum over neurons and bins."""
    ll = 0
    for n in range(Y.shape[1]):
        ll += log_likelihood_terms(X, Y[:, n], A[n], W[n], bias[n:n + 1]).sum()
    return ll


def mean(X, a, W, b):
    """regression.py:524-526."""
    return logistic(activation(X, a, W, b))


def model_means(X, A, W, bias):
    """models.py:153-163 for one dataset -> (T, N)."""
    return np.column_stack([mean(X, A[n], W[n], bias[n:n + 1]) for n in range(A.shape[0])])


def generate(weights, biases, basis, T, U):
    """models.py:98-151 with regression.py:528-541, the draw `npr.rand(N) < p` replaced by the given uniforms
    U (T, N) so that a device simulation can be replayed step for step.  weights (N, N, B), biases (N,).
    Returns X (T, N, B), Y (T, N)."""
    N = weights.shape[0]
    L, B = basis.shape
    flipped = np.flipud(basis)                                   # models.py:117-122
    Wm = weights.reshape((N, N * B))                             # models.py:124
    Y = np.zeros((T + L, N))
    X = np.zeros((T + L, N, B))
    for t in range(L, T + L):
        X[t] = Y[t - L:t].T.dot(flipped)                         # models.py:140
        psi = Wm.dot(X[t].reshape((N * B,))) + biases            # models.py:143
        Y[t] = U[t - L] < logistic(psi)                          # regression.py:538-539
    return X[L:], Y[L:]


# ---------------------------------------------------------------- Gaussian observations (regression.py:380-446)
def gaussian_omega(T, eta):
    """regression.py:419-421: omega_t = 1 / eta for every bin."""
    return np.ones(T) / eta


def gaussian_kappa(y, eta):
    """regression.py:423-424."""
    return y / eta


def gaussian_log_likelihood_terms(X, y, a, W, b, eta):
    """regression.py:400-404: -1/2 log(2 pi eta) - 1/2 (y - mean)^2 / eta per bin (eta is a variance)."""
    return -0.5 * np.log(2 * np.pi * eta) - 0.5 * (y - activation(X, a, W, b)) ** 2 / eta


def gaussian_eta_posterior(X, y, a, W, b, a_0, b_0):
    """(alpha, beta) handed to sample_invgamma by _resample_eta (regression.py:432-445).  As in the reference the
    residual sum of squares enters beta without the factor 1/2."""
    T = X.shape[0]
    return a_0 + T / 2.0, b_0 + np.sum((y - activation(X, a, W, b)) ** 2)


def generate_gaussian(weights, biases, etas, basis, T, Z):
    """models.py:98-151 with SparseGaussianRegression.rvs (regression.py:406-417): y_t = psi_t + sqrt(eta) z_t.  The
    reference draws with regressions[0] only (models.py:146), i.e. with etas[0] for every neuron; Z (T, N) are the
    standard normals.  Returns X (T, N, B), Y (T, N)."""
    N = weights.shape[0]
    L, B = basis.shape
    flipped = np.flipud(basis)
    Wm = weights.reshape((N, N * B))
    Y = np.zeros((T + L, N))
    X = np.zeros((T + L, N, B))
    for t in range(L, T + L):
        X[t] = Y[t - L:t].T.dot(flipped)
        psi = Wm.dot(X[t].reshape((N * B,))) + biases
        Y[t] = psi + np.sqrt(etas[0]) * Z[t - L]
    return X[L:], Y[L:]


def kappa(y):
    """kappa = a_func(y) - b_func(y)/2 = y - 1/2.  regression.py:510-511."""
    return y - 0.5


# --------------------------------------------------------------------------- (3) sufficient stats
def lkhd_sufficient_statistics(X, omega, kap):
    """J = [X,1]^T diag(omega) [X,1], h = [X,1]^T kappa for one dataset.  regression.py:225-262."""
    T, NB = X.shape
    J = np.zeros((NB + 1, NB + 1))
    h = np.zeros(NB + 1)
    XO = X * omega[:, None]                     # :251
    J[:NB, :NB] += XO.T.dot(X)                  # :252
    Xsum = XO.sum(0)                            # :253
    J[:NB, -1] += Xsum
    J[-1, :NB] += Xsum
    J[-1, -1] += omega.sum()                    # :256
    h[:NB] += kap.T.dot(X)                      # :259
    h[-1] += kap.sum()                          # :260
    return J, h


def natural_params(mu_w, S_w, mu_b, S_b):
    """regression.py:138-151."""
    N, B = mu_w.shape
    J_w = np.zeros((N, B, B))
    h_w = np.zeros((N, B))
    for n in range(N):
        J_w[n] = np.linalg.inv(S_w[n])
        h_w[n] = J_w[n].dot(mu_w[n])
    J_b = np.linalg.inv(S_b)
    h_b = J_b.dot(mu_b)
    return J_w, h_w, J_b, h_b


def prior_sufficient_statistics(mu_w, S_w, mu_b, S_b):
    """regression.py:210-223."""
    J_w, h_w, J_b, h_b = natural_params(mu_w, S_w, mu_b, S_b)
    J_prior = block_diag(*J_w, J_b)
    h_prior = np.concatenate((h_w.ravel(), h_b.ravel()))
    return J_prior, h_prior


# --------------------------------------------------------------------------- (4) spike and slab
def _mask(a, B):
    return np.concatenate((np.repeat(a, B), [1])).astype(bool)


def marginal_likelihood(J_prior, h_prior, J_post, h_post, a, B):
    """regression.py:343-378 (Cholesky form)."""
    m = _mask(a, B)
    J0 = J_prior[np.ix_(m, m)]
    h0 = h_prior[m]
    Jp = J_post[np.ix_(m, m)]
    hp = h_post[m]
    L0 = np.linalg.cholesky(J0)
    Lp = np.linalg.cholesky(Jp)
    ml = 0
    ml -= np.sum(np.log(np.diag(Lp)))
    ml += np.sum(np.log(np.diag(L0)))
    ml += 0.5 * hp.T.dot(dpotrs(Lp, hp, lower=True)[0])
    ml -= 0.5 * h0.T.dot(dpotrs(L0, h0, lower=True)[0])
    return ml


def sample_discrete_from_log_u(lps, u):
    """pybasicbayes.util.stats.sample_discrete_from_log for a 1-D vector with the uniform
    injected (regression.py:315).  Third-party restatement, SURVEY Appendix B.2."""
    cum = np.exp(lps - logsumexp(lps)).cumsum()
    return int(np.sum(u * cum[-1] > cum))


def collapsed_resample_a(J_prior, h_prior, J_post, h_post, a, rho, B, perm, us, trace=None):
    """regression.py:282-320 with the permutation (:286) and the per-step uniforms (:315)
    injected.  Returns the new a (copy).  trace, if a list, receives (n, lp0, lp1, v)."""
    a = np.array(a, dtype=bool).copy()
    ml_prev = marginal_likelihood(J_prior, h_prior, J_post, h_post, a, B)
    for step, n in enumerate(perm):
        lps = np.zeros(2)
        v_prev = int(a[n])
        lps[v_prev] += ml_prev
        lps[v_prev] += v_prev * np.log(rho[n]) + (1 - v_prev) * np.log(1 - rho[n])
        v_new = 1 - v_prev
        a[n] = v_new
        ml_new = marginal_likelihood(J_prior, h_prior, J_post, h_post, a, B)
        lps[v_new] += ml_new
        lps[v_new] += v_new * np.log(rho[n]) + (1 - v_new) * np.log(1 - rho[n])
        v_smpl = sample_discrete_from_log_u(lps, us[step])
        a[n] = v_smpl
        if trace is not None:
            trace.append((int(n), lps[0], lps[1], int(v_smpl)))
        if v_smpl != v_prev:
            ml_prev = ml_new
    return a


def sample_gaussian_info(J, h, z):
    """pybasicbayes.util.stats.sample_gaussian(J=, h=) with z injected (regression.py:334).
    Third-party restatement, SURVEY Appendix B.2."""
    L = np.linalg.cholesky(J)
    x = scipy.linalg.solve_triangular(L, z, lower=True, trans='T')
    return x + scipy.linalg.cho_solve((L, True), h)


def resample_W(J_post, h_post, a, B, z):
    """regression.py:323-340.  z has one entry per ACTIVE coordinate (|a| B + 1).
    Returns (W (N,B), b (1,))."""
    N = len(a)
    m = _mask(a, B)
    Jp = J_post[np.ix_(m, m)]
    hp = h_post[m]
    Wv = sample_gaussian_info(Jp, hp, z)
    W = np.zeros((N, B))
    W[np.asarray(a, dtype=bool), :] = Wv[:-1].reshape((-1, B))
    return W, np.reshape(Wv[-1], (1,))


def deterministic_sparsity(rho):
    """regression.py:153-155."""
    return np.all((rho < 1e-6) | (rho > 1 - 1e-6))


def resample_regression(X, y, a, W, b, hyper, omega, perm, us, z_full, kap=None):
    """One regression.resample (regression.py:265-280) for ONE dataset with all randomness
    injected: omega (T,), perm (N,), us (N,), z_full (N*B+1,).  z_full is indexed by
    COORDINATE: the standard normal for active coordinate d is z_full[d] (the CUDA kernel keys
    its normals by coordinate the same way), i.e. sample_gaussian receives z_full[mask].
    hyper = dict(rho, mu_w, S_w, mu_b, S_b) in expanded shapes."""
    N, B = W.shape
    Xf = flatten_X(X, N, B)
    J_prior, h_prior = prior_sufficient_statistics(hyper['mu_w'], hyper['S_w'], hyper['mu_b'], hyper['S_b'])
    J_l, h_l = lkhd_sufficient_statistics(Xf, omega, kappa(y) if kap is None else kap)   # kap: Gaussian y / eta
    J_post = J_prior + J_l
    h_post = h_prior + h_l
    rho = hyper['rho']
    if deterministic_sparsity(rho):
        a_new = np.round(rho).astype(bool)
    else:
        a_new = collapsed_resample_a(J_prior, h_prior, J_post, h_post, a, rho, B, perm, us)
    m = _mask(a_new, B)
    W_new, b_new = resample_W(J_post, h_post, a_new, B, z_full[m])
    return a_new, W_new, b_new


# --------------------------------------------------------------------------- NIW network (host step)
def sample_invwishart(S, nu, rng):
    """pybasicbayes.util.stats.sample_invwishart; third-party restatement (Appendix B.3)."""
    n = S.shape[0]
    chol = np.linalg.cholesky(S)
    if (nu <= 81 + n) and (nu == np.round(nu)):
        x = rng.standard_normal((int(nu), n))
    else:
        x = np.diag(np.sqrt(np.atleast_1d(rng.chisquare(nu - np.arange(n)))))
        x[np.triu_indices_from(x, 1)] = rng.standard_normal(n * (n - 1) // 2)
    R = np.linalg.qr(x, 'r')
    T = scipy.linalg.solve_triangular(R.T, chol.T, lower=True).T
    return np.dot(T, T.T)


def niw_posterior(data, mu_0, sigma_0, kappa_0, nu_0):
    """NIW posterior hyper-parameters for rows of `data` (n,B); n == 0 -> prior.  Appendix B.3."""
    D = len(mu_0)
    data = np.asarray(data).reshape((-1, D))
    n = data.shape[0]
    if n == 0:
        return mu_0, sigma_0, kappa_0, nu_0
    xbar = data.mean(0)
    c = data - xbar
    sumsq = c.T.dot(c)
    mu_n = kappa_0 / (kappa_0 + n) * mu_0 + n / (kappa_0 + n) * xbar
    sigma_n = sigma_0 + sumsq + kappa_0 * n / (kappa_0 + n) * np.outer(xbar - mu_0, xbar - mu_0)
    return mu_n, sigma_n, kappa_0 + n, nu_0 + n


def niw_resample(data, mu_0, sigma_0, kappa_0, nu_0, rng):
    mu_n, sigma_n, kappa_n, nu_n = niw_posterior(data, mu_0, sigma_0, kappa_0, nu_0)
    sigma = sample_invwishart(sigma_n, nu_n, rng)
    mu = rng.multivariate_normal(mu_n, sigma / kappa_n)
    return mu, sigma


# --------------------------------------------------------------------------- whole-model port
class OracleSparseBernoulliGLM(object):
    """Plain numpy port of SparseBernoulliGLM + NIWSparseNetwork following the reference's own
    control flow (models.py:166-171,224-236; regression.py:265-280; networks.py:132-149), used for
    (a) long CPU chains in the distributional tests and (b) the timed CPU baseline in bench.py.
    Dense dgemm Gram with the XO temporary and 2 Choleskys per flip, exactly like the reference.
    Randomness: numpy Generator for everything but PG, which uses oracle/pg_devroye.c."""

    def __init__(self, N, basis, rho=0.5, mu_w=0.0, S_w=1.0, mu_b=0.0, S_b=1.0, seed=0,
                 pg_threads=0):
        self.N = N
        self.basis = basis
        self.B = B = basis.shape[1]
        self.rng = np.random.default_rng(seed)
        self.seed = seed
        self.pg_calls = 0
        self.pg_threads = pg_threads
        self.hyper = []
        for _ in range(N):
            self.hyper.append(dict(rho=expand_scalar(rho, (N,)), mu_w=expand_scalar(mu_w, (N, B)),
                                   S_w=expand_cov(S_w, (N, B, B)), mu_b=expand_scalar(mu_b, (1,)),
                                   S_b=expand_cov(S_b, (1, 1))))
        # prior draw of the state (regression.py:87-92)
        self.A = np.zeros((N, N), dtype=bool)
        self.W = np.zeros((N, N, B))
        self.bias = np.zeros(N)
        for n in range(N):
            h = self.hyper[n]
            self.A[n] = self.rng.random(N) < h['rho']
            for m in range(N):
                self.W[n, m] = self.A[n, m] * self.rng.multivariate_normal(h['mu_w'][m], h['S_w'][m])
            self.bias[n] = self.rng.multivariate_normal(h['mu_b'], h['S_b'])[0]
        # NIW network state (networks.py:81-94): off-diagonal and self priors
        self.niw = dict(mu_0=np.zeros(B), sigma_0=np.eye(B), kappa_0=1.0, nu_0=3.0)
        self.net_rho = 0.5 * np.ones((N, N))
        self.data_list = []
        # wall-clock seconds spent in the T-proportional part (psi, PG, Gram) and in the rest (priors, a-scan, W draw)
        # of resample_regression: lets the timed CPU baseline extrapolate from a time sub-sample (bench.py)
        self.t_aug = 0.0
        self.t_scan = 0.0

    def add_data(self, Y, X=None):
        if X is None:
            X = convolve_with_basis(Y, self.basis)
        self.data_list.append((X, Y))

    def log_likelihood(self):
        return sum(model_log_likelihood(X, Y, self.A, self.W, self.bias) for X, Y in self.data_list)

    def resample_regression(self, n):
        """regression.py:265-280 for postsynaptic neuron n."""
        import time as _time
        N, B = self.N, self.B
        h = self.hyper[n]
        _t0 = _time.perf_counter()
        J_prior, h_prior = prior_sufficient_statistics(h['mu_w'], h['S_w'], h['mu_b'], h['S_b'])
        J_l = np.zeros_like(J_prior)
        h_l = np.zeros_like(h_prior)
        _t1 = _time.perf_counter()
        for X, Y in self.data_list:
            Xf = flatten_X(X, N, B)
            y = Y[:, n]
            psi = activation(Xf, self.A[n], self.W[n], self.bias[n:n + 1])
            self.pg_calls += 1
            omega = pg1_draw(psi, self.seed, self.pg_calls, rng_kind=1, nthreads=self.pg_threads)
            Jd, hd = lkhd_sufficient_statistics(Xf, omega, kappa(y))
            J_l += Jd
            h_l += hd
        _t2 = _time.perf_counter()
        J_post = J_prior + J_l
        h_post = h_prior + h_l
        if deterministic_sparsity(h['rho']):
            a = np.round(h['rho']).astype(bool)
        else:
            perm = self.rng.permutation(N)
            us = self.rng.random(N)
            a = collapsed_resample_a(J_prior, h_prior, J_post, h_post, self.A[n], h['rho'], B, perm, us)
        m = _mask(a, B)
        z = self.rng.standard_normal(int(m.sum()))
        Wn, bn = resample_W(J_post, h_post, a, B, z)
        self.A[n], self.W[n], self.bias[n] = a, Wn, bn[0]
        self.t_aug += _t2 - _t1
        self.t_scan += (_t1 - _t0) + (_time.perf_counter() - _t2)

    def resample_network(self):
        """models.py:228-236 + networks.py:132-149 (NIWSparseNetwork, diagonal special)."""
        N, B = self.N, self.B
        off = ~np.eye(N, dtype=bool)
        p = self.niw
        mu_o, sig_o = niw_resample(self.W[off & self.A], p['mu_0'], p['sigma_0'], p['kappa_0'],
                                   max(p['nu_0'], B + 2.), self.rng)
        mu_s, sig_s = niw_resample(self.W[np.eye(N, dtype=bool) & self.A], p['mu_0'], p['sigma_0'],
                                   p['kappa_0'], p['nu_0'], self.rng)
        for n in range(N):
            mu_w = np.tile(mu_o, (N, 1))
            S_w = np.tile(sig_o, (N, 1, 1))
            mu_w[n] = mu_s
            S_w[n] = sig_s
            self.hyper[n]['mu_w'] = mu_w
            self.hyper[n]['S_w'] = S_w
            self.hyper[n]['rho'] = self.net_rho[n]

    def resample_model(self, neurons=None):
        for n in (range(self.N) if neurons is None else neurons):
            self.resample_regression(n)
        self.resample_network()


# --------------------------------------------------------------------------- integer-digit Gram (checker)
# Restatement of the arithmetic of pyglm_b200/csrc/gram_tc.cu (OUR tensor-core formulation of
# regression.py:251-256, not a reference function): used by the tests to check the digit planes and the int64
# sums of the tcgen05 kernel bit for bit.  The FP64 value it approximates is lkhd_sufficient_statistics above.
def tc_bound(cmax):
    """The value that maps to 2^(8S) in a row's fixed-point representation: 1.02 x its largest entry (1 for an all-zero
    row).  Not rounded to a power of two (gram_tc.cu: tc_bound)."""
    return cmax * 1.02 if cmax > 0 else 1.0


def tc_scale(cmax, S):
    return 2.0 ** (8 * S) / tc_bound(cmax)


def tc_digits_int(v, S):
    """S radix-256 digits of the integer array v >= 0: digit 0 (most significant) unsigned, the others signed."""
    v = np.asarray(v, dtype=np.int64).copy()
    out = []
    for _ in range(S - 1):
        lo = ((v + 128) & 255) - 128
        out.append(lo)
        v = (v - lo) >> 8
    out.append(v)
    return out[::-1]


def tc_digits(scaled, S):
    """S radix-256 digits of rint(scaled) (the omega operand)."""
    return tc_digits_int(np.rint(scaled).astype(np.int64), S)


def tc_dither(p, t):
    """Index-keyed dither in (-1/2, 1/2) added before rounding X~ to its fixed-point grid (gram_tc.cu: tc_dither)."""
    with np.errstate(over="ignore"):
        h = np.asarray(p, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) \
            + np.asarray(t, dtype=np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
        h ^= h >> np.uint64(33)
        h *= np.uint64(0xFF51AFD7ED558CCD)
        h ^= h >> np.uint64(33)
        h *= np.uint64(0xC4CEB9FE1A85EC53)
        h ^= h >> np.uint64(33)
    return ((h >> np.uint64(40)).astype(np.float64) + 0.5) * 2.0 ** -24 - 0.5


def tc_rword(t):
    """Per-bin dither word of the product rounding (gram_tc.cu: tc_rword = lowbias32 of the global bin index)."""
    t = np.asarray(t, dtype=np.uint64)
    m = np.uint64(0xFFFFFFFF)
    x = ((t & m) ^ (((t >> np.uint64(32)) * np.uint64(0x9E3779B9)) & m)) & m
    x = ((x ^ (x >> np.uint64(16))) * np.uint64(0x7FEB352D)) & m
    x = ((x ^ (x >> np.uint64(15))) * np.uint64(0x846CA68B)) & m
    return x ^ (x >> np.uint64(16))


def tc_xq(Xt, i, bx, S, t_off=0):
    """Fixed-point column i of the design: rint(x 2^(8S) / bx_i + dither(i, global bin)) as uint64."""
    T = Xt.shape[0]
    t = np.arange(T, dtype=np.uint64) + np.uint64(t_off)
    v = np.rint(Xt[:, i] * (2.0 ** (8 * S) / bx[i]) + tc_dither(np.full(T, i, dtype=np.uint64), t))
    return v.astype(np.int64).astype(np.uint64)


def tc_z_fix(Xt, i, j, ex, S, t_off=0):
    """Z_fix[t] = (xq_i xq_j + rnd(t)) >> 8S as gram_tc.cu forms it (integer arithmetic; pair index i(i+1)/2 + j)."""
    T = Xt.shape[0]
    rw = tc_rword(np.arange(T, dtype=np.uint64) + np.uint64(t_off))
    xi, xj = tc_xq(Xt, i, ex, S, t_off), tc_xq(Xt, j, ex, S, t_off)
    if S == 4:
        return ((xi * xj + rw) >> np.uint64(32)).astype(np.int64)           # < 2^64: exact in uint64
    assert S == 5
    add = (rw << np.uint64(8)) | (rw >> np.uint64(24))
    return np.array([(int(a) * int(b) + int(c)) >> 40 for a, b, c in zip(xi, xj, add)], dtype=np.int64)


def tc_z_digits(Xt, i, j, ex, S, t_off=0):
    """Digits of Z[:, (i,j)] = Xt[:, i] Xt[:, j] as gram_tc.cu stores / builds them."""
    return tc_digits_int(tc_z_fix(Xt, i, j, ex, S, t_off), S)


def tc_gram_reference(Xt, om, S, t_off=0, ex=None, eo=None):
    """Xt (T, D) = [X, 1] >= 0, om (T, n) > 0.  Returns (Jint (n, D(D+1)/2) int64 exact digit sums,
    J (n, D, D) float64 lower triangle) as gram_tc.cu computes them.  ex / eo: scale exponents when they come from a
    longer recording than Xt (time slabs): the rows' bounds (tc_bound of the column maxima).  The digit dot products run as float64 matrix products: every partial sum
    is an integer below 2^53, hence exact."""
    T, D = Xt.shape
    n = om.shape[1]
    assert T * 4 * 255 * 255 < 2 ** 53
    ex = [tc_bound(c) for c in Xt.max(0)] if ex is None else ex
    eo = [tc_bound(c) for c in om.max(0)] if eo is None else eo
    od = [np.stack([tc_digits(om[:, c] * (2.0 ** (8 * S) / eo[c]), S)[b] for c in range(n)], axis=1).astype(np.float64)
          for b in range(S)]                                              # od[b]: (T, n)
    xq = np.stack([tc_xq(Xt, i, ex, S, t_off) for i in range(D)], axis=1)  # (T, D) uint64
    rw = tc_rword(np.arange(T, dtype=np.uint64) + np.uint64(t_off))
    Jint = np.zeros((n, D * (D + 1) // 2), dtype=np.int64)
    J = np.zeros((n, D, D))
    for i in range(D):
        if S == 4:
            z = ((xq[:, i:i + 1] * xq[:, :i + 1] + rw[:, None]) >> np.uint64(32)).astype(np.int64)   # (T, i+1)
        else:
            z = np.stack([tc_z_fix(Xt, i, j, ex, S, t_off) for j in range(i + 1)], axis=1)
        zd = [d.astype(np.float64) for d in tc_digits_int(z, S)]          # zd[a]: (T, i+1)
        tot = np.zeros((i + 1, n), dtype=np.int64)
        for a in range(S):
            for b in range(S - a):
                tot += (zd[a].T @ od[b]).astype(np.int64) << (8 * (S - 1 - a - b))
        p0 = i * (i + 1) // 2
        Jint[:, p0:p0 + i + 1] = tot.T
        for c in range(n):
            J[c, i, :i + 1] = tot[:, c].astype(np.float64) * (((ex[i] * np.array(ex[:i + 1])) * eo[c]) * 2.0 ** (-8 * S - 8))
    return Jint, J
