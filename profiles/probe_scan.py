"""Probe: per-sweep density of the adjacency and device time of the spike-and-slab kernel at cfg3."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyglm_b200.models import SparseBernoulliGLM
from pyglm_b200.utils.basis import cosine_basis
N, B, T = 200, 2, 100000
np.random.seed(0)
basis = cosine_basis(B=B, L=100) / 100
Y = (np.random.default_rng(0).random((T, N)) < 0.05).astype(np.float64)
m = SparseBernoulliGLM(N, basis=basis, regression_kwargs=dict(S_w=10.0, mu_b=-2.), seed=1234)
m.add_data(Y, host_X=False)
K = m.engine.K
orig = K.spike_slab_update
def timed(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = orig(*a, **k); e1.record(); torch.cuda.synchronize()
    timed.ms = e0.elapsed_time(e1); return out
K.spike_slab_update = timed
for it in range(6):
    dens0 = m.adjacency.mean()
    t0 = time.perf_counter(); m.resample_regressions(); torch.cuda.synchronize(); t1 = time.perf_counter()
    t2 = time.perf_counter(); m.resample_network(); t3 = time.perf_counter()
    print("sweep %d: density before %.3f after %.3f  spike_slab %.2f ms  regressions %.1f ms  network(host) %.1f ms  max row %d"
          % (it, dens0, m.adjacency.mean(), timed.ms, (t1 - t0) * 1e3, (t3 - t2) * 1e3, m.adjacency.sum(1).max()))
