"""Small host helpers with the semantics of pyglm/utils/utils.py (they define which constructor inputs
are accepted: scalars are broadcast, arrays must already have the full shape)."""
import numpy as np


def logistic(x):
    """1 / (1 + e^-x)  (pyglm/utils/utils.py:3-4)."""
    return 1.0 / (1.0 + np.exp(-np.asarray(x, dtype=np.float64)))


def expand_scalar(x, shp):
    """Scalar -> constant array of shape shp; an array must already have that shape and is passed through
    uncopied (pyglm/utils/utils.py:7-12, including its aliasing of array inputs)."""
    if np.isscalar(x):
        return np.full(shp, x, dtype=np.float64)
    assert x.shape == tuple(shp), "expected shape %s, got %s" % (tuple(shp), x.shape)
    return x


def expand_cov(c, shp):
    """Scalar c -> c * I tiled to shape (..., d, d); arrays are checked and passed through
    (pyglm/utils/utils.py:15-27)."""
    shp = tuple(shp)
    assert len(shp) >= 2 and shp[-2] == shp[-1]
    if np.isscalar(c):
        out = np.zeros(shp, dtype=np.float64)
        idx = np.arange(shp[-1])
        out[..., idx, idx] = c
        return out
    assert c.shape == shp, "expected shape %s, got %s" % (shp, c.shape)
    return c
