"""ORACLE shim: the three pybasicbayes.util.stats functions used at
pyglm/regression.py:315 (sample_discrete_from_log), :334 (sample_gaussian) and
:398/:445 (sample_invgamma).  Restated from the package's published behaviour."""
import numpy as np
import scipy.linalg
from scipy.special import logsumexp


def sample_gaussian(mu=None, Sigma=None, J=None, h=None):
    """x ~ N(J^{-1} h, J^{-1}) drawn in information form: L = chol(J) (lower);
    x = L^{-T} z + J^{-1} h with z ~ N(0, I)."""
    if J is None:
        mu = np.zeros(Sigma.shape[0]) if mu is None else mu
        return mu + np.linalg.cholesky(Sigma).dot(np.random.randn(Sigma.shape[0]))
    L = np.linalg.cholesky(J)
    z = np.random.randn(J.shape[0])
    x = scipy.linalg.solve_triangular(L, z, lower=True, trans='T')
    if h is not None:
        x = x + scipy.linalg.cho_solve((L, True), h)
    return x


def sample_discrete_from_log(p_log, axis=0, dtype=np.int32):
    """Inverse-CDF draw from unnormalised log-probabilities along `axis`."""
    lognorms = logsumexp(p_log, axis=axis)
    cumvals = np.exp(p_log - np.expand_dims(lognorms, axis)).cumsum(axis)
    thesize = np.array(p_log.shape)
    thesize[axis] = 1
    last = np.take(cumvals, [-1], axis=axis)
    randvals = np.random.random(size=thesize) * np.reshape(last, thesize)
    return np.sum(randvals > cumvals, axis=axis, dtype=dtype)


def sample_invgamma(alpha, beta):
    return 1.0 / np.random.gamma(alpha, 1.0 / beta)
