"""ORACLE shim: the NIW-conjugate Gaussian used by pyglm/networks.py:89,94,141,145,149.
Restated from pybasicbayes.distributions.Gaussian."""
import numpy as np
import scipy.linalg
from scipy import stats


def sample_invwishart(S, nu):
    n = S.shape[0]
    chol = np.linalg.cholesky(S)
    if (nu <= 81 + n) and (nu == np.round(nu)):
        x = np.random.randn(int(nu), n)
    else:
        x = np.diag(np.sqrt(np.atleast_1d(stats.chi2.rvs(nu - np.arange(n)))))
        x[np.triu_indices_from(x, 1)] = np.random.randn(n * (n - 1) // 2)
    R = np.linalg.qr(x, 'r')
    T = scipy.linalg.solve_triangular(R.T, chol.T, lower=True).T
    return np.dot(T, T.T)


def sample_niw(mu, lmbda, kappa, nu):
    lmbda = sample_invwishart(lmbda, nu)
    mu = np.random.multivariate_normal(mu, lmbda / kappa)
    return mu, lmbda


class Gaussian(object):
    def __init__(self, mu=None, sigma=None, mu_0=None, sigma_0=None, kappa_0=None, nu_0=None):
        self.mu, self.sigma = mu, sigma
        self.mu_0, self.sigma_0, self.kappa_0, self.nu_0 = mu_0, sigma_0, kappa_0, nu_0
        if mu is None and sigma is None and all(v is not None for v in (mu_0, sigma_0, kappa_0, nu_0)):
            self.resample()

    @staticmethod
    def _get_statistics(data, D):
        data = np.asarray(data).reshape((-1, D))
        n = data.shape[0]
        if n > 0:
            xbar = data.mean(0)
            centered = data - xbar
            sumsq = centered.T.dot(centered)
        else:
            xbar, sumsq = None, None
        return n, xbar, sumsq

    def _posterior_hypparams(self, n, xbar, sumsq):
        mu_0, sigma_0, kappa_0, nu_0 = self.mu_0, self.sigma_0, self.kappa_0, self.nu_0
        if n > 0:
            mu_n = self.kappa_0 / (self.kappa_0 + n) * self.mu_0 + n / (self.kappa_0 + n) * xbar
            kappa_n = self.kappa_0 + n
            nu_n = self.nu_0 + n
            sigma_n = self.sigma_0 + sumsq + \
                self.kappa_0 * n / (self.kappa_0 + n) * np.outer(xbar - self.mu_0, xbar - self.mu_0)
            return mu_n, sigma_n, kappa_n, nu_n
        return mu_0, sigma_0, kappa_0, nu_0

    def resample(self, data=[]):
        D = len(self.mu_0)
        self.mu, self.sigma = sample_niw(*self._posterior_hypparams(*self._get_statistics(data, D)))
        return self
