"""Timeline of the overlapped single-GPU sweep (engine._sweep_overlapped): every phase's (start, end) relative to one
reference event, for two steady-state sweeps, read from the per-phase CUDA events of both side streams.
    python profiles/probe_overlap_timeline.py [N] [T] [warm-up sweeps]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from bench import synthetic_spikes
from pyglm_b200.models import SparseBernoulliGLM
from pyglm_b200.utils.basis import cosine_basis

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
B, L = 2, 100
np.random.seed(0)
m = SparseBernoulliGLM(N, basis=cosine_basis(B=B, L=L) / L, regression_kwargs=dict(S_w=10.0, mu_b=-2.), seed=1234)
m.add_data(synthetic_spikes(T, N), host_X=False)
eng = m.engine
for _ in range(4):
    m.resample_model()
ds = m._device_datasets()[0]
A, W, b = m._host_state()
hyp = m._stacked_hypers()
WARM = int(sys.argv[3]) if len(sys.argv) > 3 else 3
for _ in range(WARM):
    A, W, b = eng.sweep([ds], A, W, b, hyp)
torch.cuda.synchronize()
ref = torch.cuda.Event(enable_timing=True)
ref.record()
eng.profile = {}
for _ in range(3):
    A, W, b = eng.sweep([ds], A, W, b, hyp)
torch.cuda.synchronize()
rows = []
for name, pairs in eng.profile.items():
    for a, e in pairs:
        rows.append((ref.elapsed_time(a), ref.elapsed_time(e), name))
rows.sort()
for s, e, name in rows:
    if name in ("weighted_gram", "idle_before_scan"):
        continue
    print("%8.3f -> %8.3f  (%6.3f ms)  %s" % (s, e, e - s, name))
