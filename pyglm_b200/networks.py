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
        # on the transposed copy (D contiguous rows of n numbers): reductions along the long axis of an (n, D) array
        # with D = 1..4 are several times slower, and this runs between two sweeps on the host (18 000 rows at cfg3)
        dt = np.ascontiguousarray(data.T)
        xbar = dt.sum(axis=1) / n
        dt -= xbar[:, None]
        kappa_n = self.kappa_0 + n
        mu_n = (self.kappa_0 * self.mu_0 + n * xbar) / kappa_n
        d0 = xbar - self.mu_0
        sigma_n = self.sigma_0 + dt.dot(dt.T) + (self.kappa_0 * n / kappa_n) * np.outer(d0, d0)
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

    # Hooks through which a structured WEIGHT prior takes part in the moves of the latent variables it shares with a
    # structured ADJACENCY prior (block labels z, locations L): the extra log-likelihood of the weights, 0 here.
    def _weight_block_scores(self, n):
        return 0.0

    def _weight_location_score(self, n, l):
        return 0.0

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


class BetaBernoulli(object):
    """Bernoulli probability with a conjugate Beta(a_0, b_0) prior: .rho, .resample(bits).  The class the reference
    names but never defines (networks.py:214-218): rho | bits ~ Beta(a_0 + #ones, b_0 + #zeros)."""

    def __init__(self, a_0=1.0, b_0=1.0):
        assert np.isscalar(a_0) and np.isscalar(b_0) and a_0 > 0 and b_0 > 0
        self.a_0, self.b_0 = float(a_0), float(b_0)
        self.rho = None
        self.resample()

    def posterior(self, bits=()):
        bits = np.asarray(bits, dtype=bool)
        k = int(bits.sum())
        return self.a_0 + k, self.b_0 + bits.size - k

    def resample(self, bits=()):
        a_n, b_n = self.posterior(bits)
        # keep rho strictly inside (0, 1): the scan takes log(rho) and log(1 - rho) (regression.py:302-303)
        self.rho = float(np.clip(np.random.beta(a_n, b_n), 1e-12, 1.0 - 1e-12))
        return self


def _offdiag(N):
    mask = np.ones((N, N), dtype=bool)
    mask[np.diag_indices(N)] = False
    return mask


def _log1pexp(x):
    return np.logaddexp(0.0, x)


def elliptical_slice(f, log_lkhd, sigma, cur=None):
    """One elliptical-slice move (Murray, Adams & MacKay 2010) of f ~ N(0, sigma^2 I) under log_lkhd: leaves the
    posterior invariant, always moves, no step size.  Draws from the global numpy stream."""
    nu = sigma * np.random.randn(*np.shape(f))
    cur = log_lkhd(f) if cur is None else cur
    log_y = cur + np.log(np.random.rand())
    theta = 2.0 * np.pi * np.random.rand()
    lo, hi = theta - 2.0 * np.pi, theta
    while True:
        prop = f * np.cos(theta) + nu * np.sin(theta)
        ll = log_lkhd(prop)
        if ll > log_y:
            return prop, ll
        if theta < 0.0:
            lo = theta
        else:
            hi = theta
        if hi - lo < 1e-12:                  # bracket collapsed onto the current point
            return f, cur
        theta = lo + (hi - lo) * np.random.rand()


class _IndependentBernoulliMixin(_NetworkModel):
    """Beta-Bernoulli connection probability, shared by all off-diagonal pairs; self-connections get their own when
    `is_diagonal_conn_special`.  The reference raises NotImplementedError before its body (networks.py:214); this
    is that body (networks.py:216-259) with the missing BetaBernoulli supplied.  SURVEY 8f rank 3 -- no reference
    output exists to pin against: parity unpinned, validated against the conjugate posterior."""

    def __init__(self, N, B, a_0=1.0, b_0=1.0, is_diagonal_conn_special=True, **kwargs):
        super(_IndependentBernoulliMixin, self).__init__(N, B, **kwargs)
        assert np.isscalar(a_0)
        assert np.isscalar(b_0)
        self._betabernoulli = BetaBernoulli(a_0, b_0)
        self.is_diagonal_conn_special = is_diagonal_conn_special
        if is_diagonal_conn_special:
            self._self_betabernoulli = BetaBernoulli(a_0, b_0)

    @property
    def rho(self):
        N = self.N
        rho = self._betabernoulli.rho * np.ones((N, N))
        if self.is_diagonal_conn_special:
            rho[np.diag_indices(N)] = self._self_betabernoulli.rho
        return rho

    def resample(self, data=[]):
        super(_IndependentBernoulliMixin, self).resample(data)
        A, W = data
        if self.is_diagonal_conn_special:
            self._betabernoulli.resample(A[_offdiag(self.N)])
            self._self_betabernoulli.resample(A.diagonal())
        else:
            self._betabernoulli.resample(A)

    def get_state(self):
        s = super(_IndependentBernoulliMixin, self).get_state()
        s["rho_off"] = self._betabernoulli.rho
        if self.is_diagonal_conn_special:
            s["rho_self"] = self._self_betabernoulli.rho
        return s

    def set_state(self, state):
        super(_IndependentBernoulliMixin, self).set_state(state)
        self._betabernoulli.rho = state["rho_off"]
        if self.is_diagonal_conn_special:
            self._self_betabernoulli.rho = state["rho_self"]


class _StochasticBlockAdjacencyMixin(_NetworkModel):
    """Stochastic block model over the adjacency (the "block models" TODO at networks.py:175,261; Linderman, Adams &
    Pillow 2016, the paper README.md:26-28 cites): z_n ~ Cat(pi), pi ~ Dir(alpha), p[c, c'] ~ Beta(a_0, b_0),
    rho[n, n'] = p[z_n, z_n'] for the connection n' -> n (row = postsynaptic, as `adjacency` is laid out,
    models.py:58-60).  Self-connections take a separate Beta-Bernoulli probability.  One resample = one Gibbs pass:
    z_n in turn given the rest (each class scored from the row and column of n), then p and pi from their conjugate
    posteriors.  Host step, O(N^2 C) per sweep.  Not in the reference snapshot: parity unpinned."""

    def __init__(self, N, B, C=2, alpha=1.0, a_0=1.0, b_0=1.0, z=None, **kwargs):
        super(_StochasticBlockAdjacencyMixin, self).__init__(N, B, **kwargs)
        assert C >= 1 and a_0 > 0 and b_0 > 0
        self.C, self.a_0, self.b_0 = int(C), float(a_0), float(b_0)
        self.alpha = np.array(expand_scalar(alpha, (self.C,)), dtype=np.float64)
        self.pi = np.random.dirichlet(self.alpha)
        self.z = np.random.choice(self.C, size=N, p=self.pi) if z is None else np.array(z, dtype=np.int64)
        assert self.z.shape == (N,) and self.z.min() >= 0 and self.z.max() < self.C
        self.p = np.clip(np.random.beta(self.a_0, self.b_0, size=(self.C, self.C)), 1e-12, 1 - 1e-12)
        self._self_betabernoulli = BetaBernoulli(a_0, b_0)

    @property
    def rho(self):
        rho = self.p[np.ix_(self.z, self.z)]
        rho[np.diag_indices(self.N)] = self._self_betabernoulli.rho
        return rho

    def block_scores(self, A, n):
        """log p(z_n = c | z_-n, A, p, pi) up to a constant, for every c."""
        Z = np.eye(self.C)[self.z]
        Z[n] = 0.0
        both = np.stack([A[n], A[:, n]]).astype(np.float64)
        ones = both.dot(Z)                                   # (2, C): row / column links of n into each class
        zeros = Z.sum(0)[None, :] - ones
        lp, lq = np.log(self.p), np.log1p(-self.p)
        return (np.log(self.pi) + lp.dot(ones[0]) + lq.dot(zeros[0])
                + ones[1].dot(lp) + zeros[1].dot(lq))

    def resample(self, data=[]):
        super(_StochasticBlockAdjacencyMixin, self).resample(data)
        A, W = data
        N, C = self.N, self.C
        for n in np.random.permutation(N):
            s = self.block_scores(A, n) + self._weight_block_scores(n)
            pr = np.exp(s - s.max())
            self.z[n] = np.searchsorted(np.cumsum(pr), np.random.rand() * pr.sum())
        Z = np.eye(C)[self.z]
        off = _offdiag(N)
        ones = Z.T.dot((A & off).astype(np.float64)).dot(Z)
        pairs = Z.T.dot(off.astype(np.float64)).dot(Z)
        self.p = np.clip(np.random.beta(self.a_0 + ones, self.b_0 + pairs - ones), 1e-12, 1 - 1e-12)
        self.pi = np.maximum(np.random.dirichlet(self.alpha + Z.sum(0)), 1e-300)
        self._self_betabernoulli.resample(A.diagonal())

    def get_state(self):
        s = super(_StochasticBlockAdjacencyMixin, self).get_state()
        s["sbm"] = dict(z=self.z.copy(), p=self.p.copy(), pi=self.pi.copy(), rho_self=self._self_betabernoulli.rho)
        return s

    def set_state(self, state):
        super(_StochasticBlockAdjacencyMixin, self).set_state(state)
        s = state["sbm"]
        self.z, self.p, self.pi = np.array(s["z"]), np.array(s["p"]), np.array(s["pi"])
        self._self_betabernoulli.rho = s["rho_self"]


class _LatentDistanceAdjacencyMixin(_NetworkModel):
    """Latent distance model over the adjacency (the "distance models" TODO at networks.py:261; Linderman, Adams &
    Pillow 2016): every neuron has a location l_n in R^dim, l_n ~ N(0, sigma_l^2 I), and
    rho[n, n'] = logistic(gamma - |l_n - l_n'|^2), gamma ~ N(mu_gamma, sigma_gamma^2).  Self-connections take a
    separate Beta-Bernoulli probability.  One resample = an elliptical-slice move of each l_n given the rest (its
    row and column of A), then one of gamma -- both priors are Gaussian, so the moves are exact and need no tuning.
    Host step, O(N^2 dim) per sweep.  Not in the reference snapshot: parity unpinned."""

    def __init__(self, N, B, dim=2, sigma_l=1.0, mu_gamma=0.0, sigma_gamma=1.0, a_0=1.0, b_0=1.0, L=None, **kwargs):
        super(_LatentDistanceAdjacencyMixin, self).__init__(N, B, **kwargs)
        self.dim, self.sigma_l = int(dim), float(sigma_l)
        self.mu_gamma, self.sigma_gamma = float(mu_gamma), float(sigma_gamma)
        self.L = self.sigma_l * np.random.randn(N, self.dim) if L is None else np.array(L, dtype=np.float64)
        assert self.L.shape == (N, self.dim)
        self.gamma = self.mu_gamma + self.sigma_gamma * np.random.randn()
        self._self_betabernoulli = BetaBernoulli(a_0, b_0)

    def logits(self):
        sq = (self.L * self.L).sum(1)
        return self.gamma - np.maximum(sq[:, None] + sq[None, :] - 2.0 * self.L.dot(self.L.T), 0.0)

    @property
    def rho(self):
        rho = np.clip(1.0 / (1.0 + np.exp(-self.logits())), 1e-12, 1 - 1e-12)
        rho[np.diag_indices(self.N)] = self._self_betabernoulli.rho
        return rho

    def log_likelihood_adjacency(self, A):
        """log p(A off-diagonal | L, gamma)."""
        off = _offdiag(self.N)
        x = self.logits()[off]
        return float(np.sum(A[off] * x - _log1pexp(x)))

    def location_score(self, links, n, l):
        """log p(row and column n of A | l_n = l, the rest) up to a constant; links = A + A.T."""
        diff = self.L - l
        x = self.gamma - np.einsum("md,md->m", diff, diff)
        x[n] = self.gamma                                    # the m = n term is constant in l_n (distance 0)
        return float(links[n].dot(x) - 2.0 * _log1pexp(x).sum())

    def resample(self, data=[]):
        super(_LatentDistanceAdjacencyMixin, self).resample(data)
        A, W = data
        N = self.N
        links = A.astype(np.float64) + A.T                  # links[n, m]: how many of n<-m, m<-n are present
        for n in np.random.permutation(N):
            self.L[n], _ = elliptical_slice(
                self.L[n], lambda l, n=n: self.location_score(links, n, l) + self._weight_location_score(n, l),
                self.sigma_l)
        off = _offdiag(N)
        negd = (self.logits() - self.gamma)[off]
        a_off = A[off].astype(np.float64)

        def llg(g0):
            x = g0 + self.mu_gamma + negd
            return float(np.sum(a_off * x - _log1pexp(x)))
        g0, _ = elliptical_slice(np.array(self.gamma - self.mu_gamma), llg, self.sigma_gamma)
        self.gamma = float(g0) + self.mu_gamma
        self._self_betabernoulli.resample(A.diagonal())

    def get_state(self):
        s = super(_LatentDistanceAdjacencyMixin, self).get_state()
        s["distance"] = dict(L=self.L.copy(), gamma=self.gamma, rho_self=self._self_betabernoulli.rho)
        return s

    def set_state(self, state):
        super(_LatentDistanceAdjacencyMixin, self).set_state(state)
        s = state["distance"]
        self.L, self.gamma = np.array(s["L"]), float(s["gamma"])
        self._self_betabernoulli.rho = s["rho_self"]


def _gauss_logpdf(x, mu, sigma):
    """log N(x; mu, sigma) for stacks: x, mu (..., B), sigma (..., B, B) -> (...)."""
    d = x - mu
    sol = np.linalg.solve(sigma, d[..., None])[..., 0]
    _, logdet = np.linalg.slogdet(sigma)
    return -0.5 * ((d * sol).sum(-1) + logdet + x.shape[-1] * np.log(2.0 * np.pi))


class _StochasticBlockWeightsMixin(_NetworkModel):
    """Block-dependent weights (the weight half of the "stochastic block models" TODO at networks.py:175; the SBM row
    of Linderman, Adams & Pillow 2016, README.md:26-28): W[n, n'] | a = 1 ~ N(mu[z_n, z_n'], Sigma[z_n, z_n']) for the
    connection n' -> n, every block pair with its own NIW-conjugate Gaussian (NIWGaussian, as the shared prior of
    networks.py:76-149), self-connections with theirs.  Put FIRST among the bases: combined with
    _StochasticBlockAdjacencyMixin the two share the labels z, and z_n is then resampled from the adjacency AND the
    weights of row / column n (the hook _weight_block_scores); over any other adjacency prior the mixin keeps its own
    labels, z_n ~ Cat(pi), pi ~ Dir(alpha).  Host step, O(N^2 C B^2) per sweep.  Not in the reference snapshot:
    parity unpinned, validated against brute-force conditionals and by recovery of planted structure."""

    def __init__(self, N, B, C=2, alpha=1.0, z=None, mu_0=0.0, sigma_0=1.0, kappa_0=1.0, nu_0=3.0, **kwargs):
        super(_StochasticBlockWeightsMixin, self).__init__(N, B, C=C, alpha=alpha, z=z, **kwargs)
        self._shares_z = isinstance(self, _StochasticBlockAdjacencyMixin)
        if not self._shares_z:
            self.C = int(C)
            self.alpha = np.array(expand_scalar(alpha, (self.C,)), dtype=np.float64)
            self.pi = np.random.dirichlet(self.alpha)
            self.z = np.random.choice(self.C, size=N, p=self.pi) if z is None else np.array(z, dtype=np.int64)
            assert self.z.shape == (N,) and self.z.min() >= 0 and self.z.max() < self.C
        mu_0 = expand_scalar(mu_0, (B,))
        sigma_0 = expand_cov(sigma_0, (B, B))
        nu = max(nu_0, B + 2.)
        self._block_gaussians = [[NIWGaussian(mu_0, sigma_0, kappa_0, nu) for _ in range(self.C)]
                                 for _ in range(self.C)]
        self._self_gaussian = NIWGaussian(mu_0, sigma_0, kappa_0, nu_0)
        self._AW = None

    @property
    def block_mu(self):
        return np.array([[g.mu for g in row] for row in self._block_gaussians])            # (C, C, B)

    @property
    def block_sigma(self):
        return np.array([[g.sigma for g in row] for row in self._block_gaussians])         # (C, C, B, B)

    @property
    def mu_W(self):
        mu = self.block_mu[np.ix_(self.z, self.z)]
        mu[np.diag_indices(self.N)] = self._self_gaussian.mu
        return mu

    @property
    def sigma_W(self):
        sigma = self.block_sigma[np.ix_(self.z, self.z)]
        sigma[np.diag_indices(self.N)] = self._self_gaussian.sigma
        return sigma

    def _weight_block_scores(self, n):
        """log p(present weights of row and column n | z_n = c, the rest) for every c."""
        A, W = self._AW
        mu, sigma = self.block_mu, self.block_sigma
        out = np.zeros(self.C)
        row = np.flatnonzero(A[n])
        row = row[row != n]
        col = np.flatnonzero(A[:, n])
        col = col[col != n]
        for c in range(self.C):
            if row.size:
                out[c] += _gauss_logpdf(W[n, row], mu[c, self.z[row]], sigma[c, self.z[row]]).sum()
            if col.size:
                out[c] += _gauss_logpdf(W[col, n], mu[self.z[col], c], sigma[self.z[col], c]).sum()
        return out

    def resample(self, data=[]):
        A, W = data
        self._AW = (A, W)                     # the shared-label move of the adjacency mixin reads it through the hook
        # block parameters given the labels, then the labels (own, or -- further down the chain -- the shared ones)
        off = A & _offdiag(self.N)
        for c in range(self.C):
            for c2 in range(self.C):
                m = off & (self.z[:, None] == c) & (self.z[None, :] == c2)
                self._block_gaussians[c][c2].resample(W[m])
        self._self_gaussian.resample(W[np.arange(self.N), np.arange(self.N)][A.diagonal()])
        if not self._shares_z:
            for n in np.random.permutation(self.N):
                s = np.log(self.pi) + self._weight_block_scores(n)
                pr = np.exp(s - s.max())
                self.z[n] = np.searchsorted(np.cumsum(pr), np.random.rand() * pr.sum())
            counts = np.bincount(self.z, minlength=self.C)
            self.pi = np.maximum(np.random.dirichlet(self.alpha + counts), 1e-300)
        super(_StochasticBlockWeightsMixin, self).resample(data)

    def get_state(self):
        s = super(_StochasticBlockWeightsMixin, self).get_state()
        s["block_weights"] = dict(mu=self.block_mu, sigma=self.block_sigma, self_gaussian=self._self_gaussian.get_params(),
                                  z=self.z.copy(), pi=np.array(self.pi))
        return s

    def set_state(self, state):
        super(_StochasticBlockWeightsMixin, self).set_state(state)
        s = state["block_weights"]
        for c in range(self.C):
            for c2 in range(self.C):
                self._block_gaussians[c][c2].set_params(s["mu"][c, c2], s["sigma"][c, c2])
        self._self_gaussian.set_params(**s["self_gaussian"])
        if not self._shares_z:
            self.z, self.pi = np.array(s["z"]), np.array(s["pi"])


class _LatentDistanceWeightsMixin(_NetworkModel):
    """Distance-dependent weights (the weight half of the "distance models" TODO at networks.py:261; the latent
    distance row of Linderman, Adams & Pillow 2016, whose mean weight falls off as -|l_n - l_n'|^2 + mu_0):
    W[n, n'] | a = 1 ~ N(m + beta |l_n - l_n'|^2, Sigma), with the B x 2 coefficient matrix Theta = [m, beta] and Sigma
    under the conjugate matrix-normal inverse-Wishart prior Theta | Sigma ~ MN(M_0, Sigma, V_0), Sigma ~ IW(S_0, nu_0)
    (M_0 = [mu_0, beta_0], beta_0 = -1 by default: closer neurons, stronger weights).  Self-connections keep their own
    NIW Gaussian.  Put FIRST among the bases: combined with _LatentDistanceAdjacencyMixin the two share the locations
    L, and each elliptical-slice move of l_n then sees the adjacency AND the weights of row / column n (the hook
    _weight_location_score); over any other adjacency prior the mixin keeps its own locations, l_n ~ N(0, sigma_l^2 I).
    Host step, O(N^2 (dim + B^2)) per sweep.  Not in the reference snapshot: parity unpinned."""

    def __init__(self, N, B, dim=2, sigma_l=1.0, L=None, mu_0=0.0, beta_0=-1.0, v_0=1.0, sigma_0=1.0, nu_0=3.0,
                 kappa_0=1.0, **kwargs):
        super(_LatentDistanceWeightsMixin, self).__init__(N, B, dim=dim, sigma_l=sigma_l, L=L, **kwargs)
        self._shares_L = isinstance(self, _LatentDistanceAdjacencyMixin)
        if not self._shares_L:
            self.dim, self.sigma_l = int(dim), float(sigma_l)
            self.L = self.sigma_l * np.random.randn(N, self.dim) if L is None else np.array(L, dtype=np.float64)
            assert self.L.shape == (N, self.dim)
        self.M_0 = np.stack([expand_scalar(mu_0, (B,)), expand_scalar(beta_0, (B,))], axis=1).astype(np.float64)
        self.V_0 = np.array(expand_cov(v_0, (2, 2)), dtype=np.float64)
        self.S_0 = np.array(expand_cov(sigma_0, (B, B)), dtype=np.float64)
        self.nu_0 = float(max(nu_0, B + 2.))
        self._iw = NIWGaussian(np.zeros(B), self.S_0, 1.0, self.nu_0)          # for its inverse-Wishart sampler
        self._self_gaussian = NIWGaussian(expand_scalar(mu_0, (B,)), self.S_0, kappa_0, nu_0)
        self.theta, self.sigma = None, None
        self._resample_regression(np.zeros((0, 2)), np.zeros((0, B)))           # a draw from the prior
        self._AW = None

    def sq_distances(self, L=None):
        L = self.L if L is None else L
        sq = (L * L).sum(1)
        return np.maximum(sq[:, None] + sq[None, :] - 2.0 * L.dot(L.T), 0.0)

    def regression_posterior(self, X, Y):
        """MNIW posterior of (Theta, Sigma) from design rows X (n, 2) = [1, d^2] and weights Y (n, B)."""
        V0i = np.linalg.inv(self.V_0)
        Vni = V0i + X.T.dot(X)
        Vn = np.linalg.inv(Vni)
        Mn = (self.M_0.dot(V0i) + Y.T.dot(X)).dot(Vn)
        Sn = self.S_0 + Y.T.dot(Y) + self.M_0.dot(V0i).dot(self.M_0.T) - Mn.dot(Vni).dot(Mn.T)
        return Mn, Vn, 0.5 * (Sn + Sn.T), self.nu_0 + X.shape[0]

    def _resample_regression(self, X, Y):
        Mn, Vn, Sn, nun = self.regression_posterior(X, Y)
        self.sigma = self._iw._sample_invwishart(Sn, nun)
        Z = np.random.randn(*Mn.shape)
        self.theta = Mn + np.linalg.cholesky(self.sigma).dot(Z).dot(np.linalg.cholesky(Vn).T)

    @property
    def mu_W(self):
        d2 = self.sq_distances()
        mu = self.theta[:, 0][None, None, :] + d2[:, :, None] * self.theta[:, 1][None, None, :]
        mu[np.diag_indices(self.N)] = self._self_gaussian.mu
        return mu

    @property
    def sigma_W(self):
        N, B = self.N, self.B
        sigma = np.repeat(np.reshape(self.sigma, (1, B * B)), N * N, axis=0)
        sigma[::N + 1] = np.reshape(self._self_gaussian.sigma, (B * B,))
        return sigma.reshape(N, N, B, B)

    def _weight_location_score(self, n, l):
        """log p(present weights of row and column n | l_n = l, the rest) up to a constant."""
        A, W = self._AW
        diff = self.L - l
        d2 = np.einsum("md,md->m", diff, diff)
        tot = 0.0
        for idx, Wsel in ((A[n], W[n]), (A[:, n], W[:, n])):
            m = np.flatnonzero(idx)
            m = m[m != n]
            if m.size:
                mean = self.theta[:, 0][None, :] + d2[m, None] * self.theta[:, 1][None, :]
                tot += _gauss_logpdf(Wsel[m], mean, self.sigma).sum()
        return float(tot)

    def resample(self, data=[]):
        A, W = data
        self._AW = (A, W)
        off = A & _offdiag(self.N)
        d2 = self.sq_distances()[off]
        self._resample_regression(np.stack([np.ones_like(d2), d2], axis=1), W[off])
        self._self_gaussian.resample(W[np.arange(self.N), np.arange(self.N)][A.diagonal()])
        if not self._shares_L:
            for n in np.random.permutation(self.N):
                self.L[n], _ = elliptical_slice(self.L[n], lambda l, n=n: self._weight_location_score(n, l), self.sigma_l)
        super(_LatentDistanceWeightsMixin, self).resample(data)

    def get_state(self):
        s = super(_LatentDistanceWeightsMixin, self).get_state()
        s["distance_weights"] = dict(theta=self.theta.copy(), sigma=self.sigma.copy(), L=self.L.copy(),
                                     self_gaussian=self._self_gaussian.get_params())
        return s

    def set_state(self, state):
        super(_LatentDistanceWeightsMixin, self).set_state(state)
        s = state["distance_weights"]
        self.theta, self.sigma = np.array(s["theta"]), np.array(s["sigma"])
        self._self_gaussian.set_params(**s["self_gaussian"])
        if not self._shares_L:
            self.L = np.array(s["L"])


class FixedMeanDenseNetwork(_DenseAdjacencyMixin, _FixedWeightsMixin):
    pass


class FixedMeanSparseNetwork(_FixedAdjacencyMixin, _FixedWeightsMixin):
    pass


class NIWDenseNetwork(_DenseAdjacencyMixin, _IndependentGaussianMixin):
    pass


class NIWSparseNetwork(_FixedAdjacencyMixin, _IndependentGaussianMixin):
    pass


# learned adjacency priors with the NIW weight prior.  Unlike the reference's fixed-adjacency
# combinations these forward their keyword arguments to the weight mixin.
class NIWBetaBernoulliNetwork(_IndependentBernoulliMixin, _IndependentGaussianMixin):
    pass


class NIWStochasticBlockNetwork(_StochasticBlockAdjacencyMixin, _IndependentGaussianMixin):
    pass


class NIWLatentDistanceNetwork(_LatentDistanceAdjacencyMixin, _IndependentGaussianMixin):
    pass


# the same adjacency priors over a FIXED Gaussian weight prior (mu, sigma, mu_self, sigma_self as in
# FixedMeanSparseNetwork): with the slab held fixed, absent connections are told apart by the marginal likelihood
# alone, which is what lets the block / distance structure of the graph be learned from short recordings
class FixedMeanBetaBernoulliNetwork(_IndependentBernoulliMixin, _FixedWeightsMixin):
    pass


class FixedMeanStochasticBlockNetwork(_StochasticBlockAdjacencyMixin, _FixedWeightsMixin):
    pass


class FixedMeanLatentDistanceNetwork(_LatentDistanceAdjacencyMixin, _FixedWeightsMixin):
    pass


# The paper's full structured models: weights AND adjacency depend on the same latent variables (block labels /
# locations).  The weight mixin comes first so that its constructor hands C / z (dim / L) on to the adjacency mixin and
# its resample() runs before the shared latent variables move.
class StochasticBlockNetwork(_StochasticBlockWeightsMixin, _StochasticBlockAdjacencyMixin):
    pass


class LatentDistanceNetwork(_LatentDistanceWeightsMixin, _LatentDistanceAdjacencyMixin):
    pass


# structured weights over a fixed / dense adjacency prior: the weights alone inform the labels / locations
class BlockWeightsSparseNetwork(_StochasticBlockWeightsMixin, _FixedAdjacencyMixin):
    pass


class BlockWeightsDenseNetwork(_StochasticBlockWeightsMixin, _DenseAdjacencyMixin):
    pass


class DistanceWeightsSparseNetwork(_LatentDistanceWeightsMixin, _FixedAdjacencyMixin):
    pass


class DistanceWeightsDenseNetwork(_LatentDistanceWeightsMixin, _DenseAdjacencyMixin):
    pass
