"""Per-phase device times of the sweep at a bench configuration (CUDA events inside GibbsEngine.sweep).

    python profiles/probe_sweep_phases.py [--config cfg3] [--steps 5] [--gram auto]
"""
import argparse
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import CONFIGS, synthetic_spikes  # noqa: E402
from pyglm_b200.models import SparseBernoulliGLM  # noqa: E402
from pyglm_b200.utils.basis import cosine_basis  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg3")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--gram", default="auto")
ap.add_argument("--tag", default="")
a = ap.parse_args()
cfg = CONFIGS[a.config]
N, B, L, T = cfg["N"], cfg["B"], cfg["L"], cfg["T"]
np.random.seed(0)
model = SparseBernoulliGLM(N, basis=cosine_basis(B=B, L=L) / L, regression_kwargs=dict(S_w=10.0, mu_b=-2.), seed=1234,
                           gram=a.gram)
model.add_data(synthetic_spikes(T, N), host_X=False)
for _ in range(a.warmup):
    model.resample_model()
eng = model.engine
eng.profile = {}
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    model.resample_model()
e1.record()
torch.cuda.synchronize()
out = dict(tag=a.tag, config=a.config, ms_per_sweep=e0.elapsed_time(e1) / a.steps, phases=eng.phase_ms(),
           density=float(model.adjacency.mean()))
print(json.dumps(out))
