"""cProfile of the host side of resample_model() for one rank of a `world`-GPU job (see probe_rank_share.py).

    python profiles/probe_host_profile.py [--world 8] [--shard time] [--steps 10]
"""
import argparse
import cProfile
import io
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "profiles")
from bench import CONFIGS, synthetic_spikes  # noqa: E402
from pyglm_b200.models import SparseBernoulliGLM  # noqa: E402
from pyglm_b200.utils.basis import cosine_basis  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg3")
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--shard", default="time")
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
sys.argv = sys.argv[:1]
from probe_rank_share import FakeComm  # noqa: E402  (runs nothing: guarded below)

cfg = CONFIGS[a.config]
N, B, L, T = cfg["N"], cfg["B"], cfg["L"], cfg["T"]
np.random.seed(0)
model = SparseBernoulliGLM(N, basis=cosine_basis(B=B, L=L) / L, regression_kwargs=dict(S_w=10.0, mu_b=-2.),
                           seed=1234, comm=FakeComm(a.world, 0) if a.world > 1 else None, shard=a.shard)
model.add_data(synthetic_spikes(T, N), host_X=False)
for _ in range(3):
    model.resample_model()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(a.steps):
    model.resample_model()
torch.cuda.synchronize()
pr.disable()
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(45)
print(out.getvalue())
