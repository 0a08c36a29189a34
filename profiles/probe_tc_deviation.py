"""Deviation of the tcgen05 integer-digit Gram from the FP64 kernel on the chain's own omega, sweep by sweep.

    python profiles/probe_tc_deviation.py [--config cfg3] [--digits 4]
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
ap.add_argument("--digits", type=int, default=4)
ap.add_argument("--sweeps", type=int, default=6)
a = ap.parse_args()
cfg = CONFIGS[a.config]
N, B, L, T = cfg["N"], cfg["B"], cfg["L"], cfg["T"]
np.random.seed(0)
model = SparseBernoulliGLM(N, basis=cosine_basis(B=B, L=L) / L, regression_kwargs=dict(S_w=10.0, mu_b=-2.), seed=1234,
                           gram="fp64")
model.add_data(synthetic_spikes(T, N), host_X=False)
eng = model.engine
K = eng.K
ds = model._device_datasets()[0]
D = eng.D
plan = K.gram_tc_plan(ds.Xp, D, N, a.digits)
tril = torch.tril(torch.ones(D, D, dtype=torch.bool, device=K.device))
for sweep in range(a.sweeps):
    model.resample_model()
    omega = [v for k, v in ds.buffers.items() if k[0] == "omega"][0]
    Jr = K.weighted_gram(ds.Xp, omega, D, N)[:, :D, :D]
    Jt = plan.gram(omega)[:, :D, :D]
    diff = (Jt - Jr).abs()
    rel = torch.where(tril & (Jr != 0), diff / Jr.abs(), torch.zeros_like(Jr))
    dg = torch.diagonal(Jr, dim1=1, dim2=2)
    nrm = torch.where(tril, diff / torch.sqrt(dg[:, :, None] * dg[:, None, :]), torch.zeros_like(Jr))
    om = omega[:, :N]
    print(json.dumps(dict(sweep=sweep + 1, digits=a.digits, max_rel=float(rel.max()), p999_rel=float(torch.quantile(rel[:8][:, tril].flatten(), 0.999)),
                          max_normwise=float(nrm.max()), omega_mean=float(om.mean()), omega_max=float(om.max()),
                          omega_mean_over_colmax=float((om.mean(0) / om.max(0).values).min()), density=float(model.adjacency.mean()))))
    del Jr, Jt, diff, rel, nrm
