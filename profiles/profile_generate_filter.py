"""Workload for ncu: one add_data (filter kernel) and one generate() at cfg3 size."""
import sys
import numpy as np
sys.path.insert(0, ".")
from pyglm_b200.models import SparseBernoulliGLM
from pyglm_b200.utils.basis import cosine_basis
N, B, L, T = 200, 2, 100, 100000
np.random.seed(0)
m = SparseBernoulliGLM(N, basis=cosine_basis(B=B, L=L) / L, regression_kwargs=dict(S_w=10.0, mu_b=-2.), seed=1)
for n, r in enumerate(m.regressions):
    r.a[:] = False
    r.W[:] = 0.0
    r.a[n] = True
    r.W[n, :] = -2.0
X, Y = m.generate(T=T, keep=False)
m.add_data(Y, host_X=False)
print("rate", Y.mean())
