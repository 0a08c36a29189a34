"""Per-postsynaptic-neuron sparse regressions: the host-side mirror of pyglm/regression.py.

A regression object owns the HOST copy of its state (a, W, b) and hyper-parameters (rho, mu_w, S_w, mu_b,
S_b) with the reference's attribute names, shapes and setter semantics (regression.py:73-136), so user code
such as `model.regressions[n].a[n] = True` (examples/synthetic.py:30-32) keeps working.  All arithmetic on
data (activation, mean, log_likelihood, omega, resample) runs on the GPU through pyglm_b200.kernels; inside a
model the N regressions are updated together by GibbsEngine.sweep and these per-object methods are only used
for standalone regressions (examples/bernoulli_regression.py).
"""
import numpy as np
import numpy.random as npr
import torch

from .kernels import pad_ldn
from .priors import prior_arrays
from .utils.utils import logistic, expand_scalar, expand_cov


class _SparseScalarRegressionBase(object):
    """y_t = sum_n a_n (w_n . x_{t,n}) + b + noise with a spike-and-slab prior (regression.py:40-378)."""

    def __init__(self, N, B, rho=0.5, mu_w=0.0, S_w=1.0, mu_b=0.0, S_b=1.0):
        self.N, self.B = N, B
        self.rho = rho
        self.mu_w = mu_w
        self.mu_b = mu_b
        self.S_w = S_w
        self.S_b = S_b
        # initial state: a draw from the prior (regression.py:87-92)
        self._a = npr.rand(N) < self.rho
        self._W = np.zeros((N, B))
        for n in range(N):
            self._W[n] = self._a[n] * npr.multivariate_normal(self.mu_w[n], self.S_w[n])
        self._b = np.atleast_1d(npr.multivariate_normal(self.mu_b, self.S_b)).astype(np.float64)
        # device RNG stream of this regression when used standalone
        # (drawn from numpy's global state only, like the reference's PG seeds, regression.py:476: reproducible
        # under np.random.seed)
        self._seed = int(npr.randint(2 ** 31 - 1))
        self._calls = 0

    # ---- state: plain writable numpy arrays
    @property
    def a(self):
        return self._a

    @a.setter
    def a(self, value):
        value = np.asarray(value)
        assert value.shape == (self.N,)
        self._a = value.astype(bool)

    @property
    def W(self):
        return self._W

    @W.setter
    def W(self, value):
        value = np.asarray(value, dtype=np.float64)
        assert value.shape == (self.N, self.B)
        self._W = np.array(value)

    @property
    def b(self):
        return self._b

    @b.setter
    def b(self, value):
        self._b = np.array(value, dtype=np.float64).reshape((1,))

    # ---- hyper-parameters (regression.py:95-136)
    @property
    def rho(self):
        return self._rho

    @rho.setter
    def rho(self, value):
        self._rho = expand_scalar(value, (self.N,))

    @property
    def mu_w(self):
        return self._mu_w

    @mu_w.setter
    def mu_w(self, value):
        self._mu_w = expand_scalar(value, (self.N, self.B))

    @property
    def mu_b(self):
        return self._mu_b

    @mu_b.setter
    def mu_b(self, value):
        self._mu_b = expand_scalar(value, (1,))

    @property
    def S_w(self):
        return self._S_w

    @S_w.setter
    def S_w(self, value):
        self._S_w = expand_cov(value, (self.N, self.B, self.B))

    @property
    def S_b(self):
        return self._S_b

    @S_b.setter
    def S_b(self, value):
        assert np.isscalar(value)
        self._S_b = expand_cov(value, (1, 1))

    @property
    def natural_params(self):
        """(J_w, h_w, J_b, h_b) as regression.py:138-151."""
        J_w = np.linalg.inv(self.S_w)
        h_w = np.einsum("nbc,nc->nb", J_w, self.mu_w)
        J_b = np.linalg.inv(self.S_b)
        return J_w, h_w, J_b, J_b.dot(self.mu_b)

    @property
    def deterministic_sparsity(self):
        return bool(np.all((self.rho < 1e-6) | (self.rho > 1 - 1e-6)))

    # ---- data handling (regression.py:173-193)
    def _flatten_X(self, X):
        if X.ndim == 2:
            assert X.shape[1] == self.N * self.B
        elif X.ndim == 3:
            X = np.reshape(X, (-1, self.N * self.B))
        else:
            raise Exception
        return X

    def extract_data(self, data):
        assert isinstance(data, tuple) and len(data) == 2
        X, y = data
        T = X.shape[0]
        assert y.shape == (T, 1) or y.shape == (T,)
        return self._flatten_X(np.asarray(X)), y

    # ---- device helpers for standalone use
    def _kernels(self):
        from .engine import default_kernels
        return default_kernels()

    def _device_design(self, X):
        K = self._kernels()
        X = np.ascontiguousarray(self._flatten_X(np.asarray(X, dtype=np.float64)))
        return K, K.pack_design(K.to_device(X))

    def _device_Wt(self, K, ldx):
        NB = self.N * self.B
        Wt = np.zeros((ldx, pad_ldn(1)))
        Wt[:NB, 0] = (self.a[:, None] * self.W).reshape(NB)
        Wt[NB, 0] = self.b[0]
        return K.to_device(Wt)

    def _device_activation(self, X):
        K, Xp = self._device_design(X)
        psi = K.activation(Xp, self._device_Wt(K, Xp.shape[1]), self.N * self.B + 1, 1)
        return K, Xp, psi

    def activation(self, X):
        """psi = X . vec(a o W) + b (regression.py:195-201)."""
        return self._device_activation(X)[2][:, 0].cpu().numpy()

    def mean(self, X):
        raise NotImplementedError

    def omega(self, X, y):
        raise NotImplementedError

    def kappa(self, X, y):
        raise NotImplementedError

    # ---- Gibbs update of one standalone regression (regression.py:265-280)
    def resample(self, datas):
        N, B = self.N, self.B
        D = N * B + 1
        K = self._kernels()
        J = h = None
        for data in datas:
            assert isinstance(data, tuple)
            X, y = self.extract_data(data)
            y = np.asarray(y, dtype=np.float64).reshape(-1)
            K, Xp = self._device_design(X)
            T = Xp.shape[0]
            wcol = K.zeros(T, pad_ldn(1))
            wcol[:, 0] = K.to_device(self.omega_device(K, Xp, y))
            Jd = K.weighted_gram(Xp, wcol, D, 1)
            wcol[:, 0] = K.to_device(self.kappa(X, y))
            hd = K.xt_kappa(Xp, wcol, D, 1)
            J = Jd if J is None else J + Jd
            h = hd if h is None else h + hd
        if J is None:                       # no data: posterior = prior
            J = K.zeros(1, Xp_ld(D), Xp_ld(D))
            h = K.zeros(1, Xp_ld(D))
        pr = prior_arrays(self.rho[None], self.mu_w[None], self.S_w[None], self.mu_b, self.S_b.reshape(-1))
        do_scan = pr.pop("do_scan")
        a0 = np.array(self.a if do_scan[0] else np.round(self.rho), dtype=np.uint8)[None]
        self._calls += 1
        perm, us, z = K.scan_randomness(N, B, 1, 0, self._seed, 2 * self._calls + 1)
        a_dev = K.to_device(a0)
        W, bias, _, _, status = K.spike_slab_update(
            N, B, J, h, {k: K.to_device(v) for k, v in pr.items()}, perm, us, z,
            K.to_device(do_scan.astype(np.uint8)), a_dev)
        if int(status.cpu()[0]) != 0:
            raise FloatingPointError("spike-and-slab update lost positive definiteness")
        self._a = a_dev.cpu().numpy()[0].astype(bool)
        self._W = W.cpu().numpy()[0]
        self._b = bias.cpu().numpy().reshape((1,))


def Xp_ld(D):
    from .kernels import pad_ldx
    return pad_ldx(D)


def sample_invgamma(alpha, beta):
    """pybasicbayes.util.stats.sample_invgamma as called at regression.py:398, 445: 1 / Gamma(alpha, scale = 1/beta)."""
    return 1.0 / npr.gamma(alpha, 1.0 / np.asarray(beta, dtype=np.float64))


class SparseGaussianRegression(_SparseScalarRegressionBase):
    """Sparse regression with Gaussian observations of variance eta (regression.py:380-446).  The Gibbs update is
    the Bernoulli one with omega = 1/eta and kappa = y/eta, so the sufficient statistics are X~^T X~ / eta and
    X~^T y / eta: the same Gram and spike-and-slab kernels, no augmentation."""

    def __init__(self, N, B, a_0=2.0, b_0=2.0, eta=None, **kwargs):
        super(SparseGaussianRegression, self).__init__(N, B, **kwargs)
        assert np.isscalar(a_0) and a_0 > 0
        assert np.isscalar(b_0) and a_0 > 0                  # sic (regression.py:392)
        self.a_0, self.b_0 = a_0, b_0
        if eta is not None:
            assert np.isscalar(eta) and eta > 0
            self.eta = eta
        else:
            self.eta = float(sample_invgamma(self.a_0, self.b_0))

    def log_likelihood(self, x):
        X, y = self.extract_data(x)
        return -0.5 * np.log(2 * np.pi * self.eta) - 0.5 * (y - self.mean(X)) ** 2 / self.eta

    def rvs(self, size=[], X=None, psi=None):
        if psi is None:
            if X is None:
                assert isinstance(size, int)
                X = npr.randn(size, self.N * self.B)
            psi = self.mean(self._flatten_X(X))
        return psi + np.sqrt(self.eta) * npr.randn(*psi.shape)

    def omega_device(self, K, Xp, y):
        return np.ones(Xp.shape[0]) / self.eta

    def omega(self, X, y):
        return 1. / self.eta * np.ones(X.shape[0])

    def kappa(self, X, y):
        return y / self.eta

    def mean(self, X):
        return self.activation(X)

    def resample(self, datas):
        super(SparseGaussianRegression, self).resample(datas)
        self._resample_eta(datas)

    def _resample_eta(self, datas):
        alpha, beta = self.a_0, self.b_0
        for data in datas:
            X, y = self.extract_data(data)
            alpha += X.shape[0] / 2.0
            beta += np.sum((y - self.mean(X)) ** 2)          # no factor 1/2, as regression.py:443
        self.eta = float(sample_invgamma(alpha, beta))


class GaussianRegression(SparseGaussianRegression):
    """Dense weights: rho = 1 (regression.py:448-456)."""

    def __init__(self, N, B, **kwargs):
        kwargs["rho"] = np.ones(N)
        super(GaussianRegression, self).__init__(N, B, **kwargs)


class _SparsePGRegressionBase(_SparseScalarRegressionBase):
    """Count observations through Polya-gamma augmentation (regression.py:459-511)."""

    def a_func(self, y):
        raise NotImplementedError

    def b_func(self, y):
        raise NotImplementedError

    def c_func(self, y):
        raise NotImplementedError

    def log_likelihood(self, x):
        """Per-bin log-likelihood (regression.py:491-494); log(1+e^psi) is evaluated overflow-free."""
        X, y = self.extract_data(x)
        psi = self.activation(X)
        softplus = np.maximum(psi, 0.0) + np.log1p(np.exp(-np.abs(psi)))
        return np.log(self.c_func(y)) + self.a_func(y) * psi - self.b_func(y) * softplus

    def omega_device(self, K, Xp, y):
        """omega ~ PG(b(y), psi) on the device; only b == 1 (Bernoulli) is implemented, as in the reference."""
        assert np.all(self.b_func(y) == 1.0), "only PG(1, psi) is implemented"
        psi = K.activation(Xp, self._device_Wt(K, Xp.shape[1]), self.N * self.B + 1, 1)
        om = K.zeros(psi.shape[0], psi.shape[1])
        self._calls += 1
        K.pg_draw(psi, 1, om, self._seed, 2 * self._calls, 0, 0, 1)
        return om[:, 0].cpu().numpy()

    def omega(self, X, y):
        """The Polya-gamma precisions for (X, y) (regression.py:496-508)."""
        K, Xp = self._device_design(X)
        return self.omega_device(K, Xp, np.asarray(y, dtype=np.float64).reshape(-1)).reshape(np.shape(y))

    def kappa(self, X, y):
        return self.a_func(y) - self.b_func(y) / 2.0


class SparseBernoulliRegression(_SparsePGRegressionBase):
    """Bernoulli observations with a logistic link (regression.py:514-541)."""

    def a_func(self, data):
        return data

    def b_func(self, data):
        return np.ones_like(data, dtype=float)

    def c_func(self, data):
        return 1.0

    def mean(self, X):
        return logistic(self.activation(X))

    def rvs(self, X=None, size=[], psi=None):
        if psi is None:
            if X is None:
                assert isinstance(size, int)
                X = npr.randn(size, self.N * self.B)
            p = self.mean(self._flatten_X(X))
        else:
            p = logistic(psi)
        return npr.rand(*p.shape) < p


class BernoulliRegression(SparseBernoulliRegression):
    """Dense weights: rho = 1 (regression.py:544-552)."""

    def __init__(self, N, B, **kwargs):
        kwargs["rho"] = np.ones(N)
        super(BernoulliRegression, self).__init__(N, B, **kwargs)
