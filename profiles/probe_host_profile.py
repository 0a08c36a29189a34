"""cProfile of the host side of overlapped sweeps through resample_model() (where the host's time between two sweeps
goes: it must stay below the ~6 ms the pre-launched augmentation keeps the device busy)."""
import sys, cProfile, pstats
import numpy as np
sys.path.insert(0, ".")
from bench import synthetic_spikes
from pyglm_b200.models import SparseBernoulliGLM
from pyglm_b200.utils.basis import cosine_basis
N, T, B, L = 200, 100000, 2, 100
np.random.seed(0)
m = SparseBernoulliGLM(N, basis=cosine_basis(B=B, L=L) / L, regression_kwargs=dict(S_w=10.0, mu_b=-2.), seed=1234)
m.add_data(synthetic_spikes(T, N), host_X=False)
for _ in range(5):
    m.resample_model()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    m.resample_model()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
