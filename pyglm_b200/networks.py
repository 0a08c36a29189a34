"""Network priors: hierarchical priors over the weights / adjacency of the N regressions.

API-compatible with pyglm/networks.py (`rho` (N,N), `mu_W` (N,N,B), `sigma_W` (N,N,B,B),
`resample((A, W))`, the four class combinations).  Per BASELINE.json's north star the hyper-parameter
update stays a small HOST step: B x B matrices once per sweep.  The NIW-conjugate Gaussian that the
reference takes from pybasicbayes (networks.py:9,89,94) is restated here.
"""
import numpy as np
import scipy.linalg

from .utils.utils import expand_scalar, expand_cov


class NIWGaussian(object):
    """Gaussian with a normal-inverse-Wishart prior: .mu, .sigma, .resample(data).
    Stands for pybasicbayes.distributions.Gaussian as used at networks.py:89,94,141,145,149."""

    def __init__(self, mu_0, sigma_0, kappa_0, nu_0, rng=None):
        self.mu_0 = np.asarray(mu_0, dtype=np.float64)
        self.sigma_0 = np.asarray(sigma_0, dtype=np.float64)
        self.kappa_0, self.nu_0 = float(kappa_0), float(nu_0)
        self._rng = rng
        self.mu, self.sigma = None, None
        self.resample()          # constructing with hyper-parameters only draws (mu, sigma) from the prior

    def _randn(self, *shape):
        return np.random.randn(*shape) if self._rng is None else self._rng.standard_normal(shape)

    def _posterior(self, data):
        D = self.mu_0.shape[0]
        data = np.asarray(data, dtype=np.float64).reshape((-1, D))
        n = data.shape[0]
        if n == 0:
            return self.mu_0, self.sigma_0, self.kappa_0, self.nu_0
        xbar = data.mean(axis=0)
        dev = data - xbar
        kappa_n = self.kappa_0 + n
        mu_n = (self.kappa_0 * self.mu_0 + n * xbar) / kappa_n
        d0 = xbar - self.mu_0
        sigma_n = self.sigma_0 + dev.T.dot(dev) + (self.kappa_0 * n / kappa_n) * np.outer(d0, d0)
        return mu_n, sigma_n, kappa_n, self.nu_0 + n

    def _sample_invwishart(self, S, nu):
        d = S.shape[0]
        chol = np.linalg.cholesky(S)
        if nu <= 81 + d and nu == round(nu):
            x = self._randn(int(nu), d)
        else:
            chi = np.random.chisquare(nu - np.arange(d)) if self._rng is None else self._rng.chisquare(nu - np.arange(d))
            x = np.diag(np.sqrt(np.atleast_1d(chi)))
            x[np.triu_indices_from(x, 1)] = self._randn(d * (d - 1) // 2)
        R = np.linalg.qr(x, "r")
        T = scipy.linalg.solve_triangular(R.T, chol.T, lower=True).T
        return T.dot(T.T)

    def resample(self, data=()):
        mu_n, sigma_n, kappa_n, nu_n = self._posterior(data)
        assert nu_n > sigma_n.shape[0] - 1 and kappa_n > 0
        self.sigma = self._sample_invwishart(sigma_n, nu_n)
        L = np.linalg.cholesky(self.sigma / kappa_n)
        self.mu = mu_n + L.dot(self._randn(L.shape[0]))
        return self

    def get_params(self):
        return dict(mu=self.mu.copy(), sigma=self.sigma.copy())

    def set_params(self, mu, sigma):
        self.mu, self.sigma = np.array(mu), np.array(sigma)


class _NetworkModel(object):
    """Base: stores N (nodes) and B (weight dimension); checks resample's input (networks.py:13-42)."""

    def __init__(self, N, B, **kwargs):
        self.N, self.B = N, B

    def resample(self, data=[]):
        assert isinstance(data, tuple)
        A, W = data
        assert A.shape == (self.N, self.N) and A.dtype == bool
        assert W.shape == (self.N, self.N, self.B)

    def log_likelihood(self, x):
        return 0

    def rvs(self, size=[]):
        return None

    # state exchanged between ranks in multi-GPU runs (rank 0 resamples, the others receive)
    def get_state(self):
        return {}

    def set_state(self, state):
        pass


class _IndependentGaussianMixin(_NetworkModel):
    """Every weight is Gaussian with a shared NIW prior; self-connections get their own
    (networks.py:76-149)."""

    def __init__(self, N, B, mu_0=0.0, sigma_0=1.0, kappa_0=1.0, nu_0=3.0,
                 is_diagonal_weight_special=True, **kwargs):
        super(_IndependentGaussianMixin, self).__init__(N, B)
        mu_0 = expand_scalar(mu_0, (B,))
        sigma_0 = expand_cov(sigma_0, (B, B))
        self._gaussian = NIWGaussian(mu_0, sigma_0, kappa_0, max(nu_0, B + 2.))
        self.is_diagonal_weight_special = is_diagonal_weight_special
        if is_diagonal_weight_special:
            self._self_gaussian = NIWGaussian(mu_0, sigma_0, kappa_0, nu_0)

    @property
    def mu_W(self):
        N, B = self.N, self.B
        mu = np.repeat(np.reshape(self._gaussian.mu, (1, B)), N * N, axis=0)
        if self.is_diagonal_weight_special:
            mu[::N + 1] = self._self_gaussian.mu
        return mu.reshape(N, N, B)

    @property
    def sigma_W(self):
        N, B = self.N, self.B
        sigma = np.repeat(np.reshape(self._gaussian.sigma, (1, B * B)), N * N, axis=0)
        if self.is_diagonal_weight_special:
            sigma[::N + 1] = np.reshape(self._self_gaussian.sigma, (B * B,))
        return sigma.reshape(N, N, B, B)

    def resample(self, data=[]):
        super(_IndependentGaussianMixin, self).resample(data)
        A, W = data
        if self.is_diagonal_weight_special:
            # W[~eye & A] and W[eye & A] (networks.py:137-145), row-major order kept
            N = self.N
            off = A.copy()
            off[np.arange(N), np.arange(N)] = False
            self._gaussian.resample(W.reshape(N * N, self.B)[np.flatnonzero(off)])
            self._self_gaussian.resample(W[np.arange(N), np.arange(N)][A.diagonal()])
        else:
            self._gaussian.resample(W[A])

    def get_state(self):
        s = super(_IndependentGaussianMixin, self).get_state()
        s["gaussian"] = self._gaussian.get_params()
        if self.is_diagonal_weight_special:
            s["self_gaussian"] = self._self_gaussian.get_params()
        return s

    def set_state(self, state):
        super(_IndependentGaussianMixin, self).set_state(state)
        self._gaussian.set_params(**state["gaussian"])
        if self.is_diagonal_weight_special:
            self._self_gaussian.set_params(**state["self_gaussian"])


class _FixedWeightsMixin(_NetworkModel):
    """Fixed Gaussian prior on every weight (networks.py:151-173).  The reference builds `_sigma` from `mu`
    (networks.py:158); here sigma is used, which is what the signature promises."""

    def __init__(self, N, B, mu=0.0, sigma=1.0, mu_self=None, sigma_self=None, **kwargs):
        super(_FixedWeightsMixin, self).__init__(N, B)
        self._mu = np.array(expand_scalar(mu, (N, N, B)), dtype=np.float64)
        self._sigma = np.array(expand_cov(sigma, (N, N, B, B)), dtype=np.float64)
        if (mu_self is not None) and (sigma_self is not None):
            self._mu[np.arange(N), np.arange(N), :] = expand_scalar(mu_self, (N, B))
            self._sigma[np.arange(N), np.arange(N), :] = expand_cov(sigma_self, (N, B, B))

    @property
    def mu_W(self):
        return self._mu

    @property
    def sigma_W(self):
        return self._sigma

    def resample(self, data=[]):
        super(_FixedWeightsMixin, self).resample(data)


class _FixedAdjacencyMixin(_NetworkModel):
    """Fixed connection probability (networks.py:178-190).  Like the reference, extra keyword arguments are
    NOT forwarded to the weight mixin: only rho / rho_self take effect."""

    def __init__(self, N, B, rho=0.5, rho_self=None, **kwargs):
        super(_FixedAdjacencyMixin, self).__init__(N, B)
        self._rho = np.array(expand_scalar(rho, (N, N)), dtype=np.float64)
        if rho_self is not None:
            self._rho[np.diag_indices(N)] = rho_self

    @property
    def rho(self):
        return self._rho

    def resample(self, data=[]):
        super(_FixedAdjacencyMixin, self).resample(data)


class _DenseAdjacencyMixin(_NetworkModel):
    """Fully connected: rho = 1 (networks.py:194-204)."""

    def __init__(self, N, B, **kwargs):
        super(_DenseAdjacencyMixin, self).__init__(N, B)
        self._rho = np.ones((N, N))

    @property
    def rho(self):
        return self._rho

    def resample(self, data=[]):
        super(_DenseAdjacencyMixin, self).resample(data)


class _IndependentBernoulliMixin(_NetworkModel):
    """Beta-Bernoulli connection probability: not implemented in the reference either (networks.py:214)."""

    def __init__(self, N, B, a_0=1.0, b_0=1.0, is_diagonal_conn_special=True, **kwargs):
        super(_IndependentBernoulliMixin, self).__init__(N, B)
        raise NotImplementedError("TODO: Implement the BetaBernoulli class")


class FixedMeanDenseNetwork(_DenseAdjacencyMixin, _FixedWeightsMixin):
    pass


class FixedMeanSparseNetwork(_FixedAdjacencyMixin, _FixedWeightsMixin):
    pass


class NIWDenseNetwork(_DenseAdjacencyMixin, _IndependentGaussianMixin):
    pass


class NIWSparseNetwork(_FixedAdjacencyMixin, _IndependentGaussianMixin):
    pass
