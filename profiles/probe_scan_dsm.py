"""Probe: the two spike-and-slab kernels on the SAME call at the benchmark shape (cfg3: N=200, B=2, D=401).

A cfg3 chain is run for a few sweeps so that the state has a realistic density; the arguments of one
pyglm_spike_slab_update call are captured and replayed with the one-CTA-per-neuron kernel (PYGLM_SS_VARIANT=1) and
with the cluster kernel (88: 8 CTAs per neuron, P in distributed shared memory) for the first n_loc neurons,
n_loc = 200 / 100 / 50 / 25 (the scan blocks of 1 / 2 / 4 / 8 ranks).  Prints ms per call (CUDA events, median of 5)
and the agreement of the two kernels (adjacency, max |dW|).    python profiles/probe_scan_dsm.py [sweeps]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pyglm_b200.models import SparseBernoulliGLM  # noqa: E402
from pyglm_b200.utils.basis import cosine_basis  # noqa: E402

N, B, T = 200, 2, 100000
sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
np.random.seed(0)
basis = cosine_basis(B=B, L=100) / 100
Y = (np.random.default_rng(0).random((T, N)) < 0.05).astype(np.float64)
m = SparseBernoulliGLM(N, basis=basis, regression_kwargs=dict(S_w=10.0, mu_b=-2.), seed=1234)
m.add_data(Y, host_X=False)
K = m.engine.K
orig = K.spike_slab_update
captured = {}


def capture(*a, **k):
    captured["args"] = [x.clone() if torch.is_tensor(x) else x for x in a]
    captured["prior"] = {kk: v.clone() for kk, v in a[4].items()}
    return orig(*a, **k)


os.environ["PYGLM_SS_VARIANT"] = "1"
for it in range(sweeps):
    if it == sweeps - 1:
        K.spike_slab_update = capture
    m.resample_model()
K.spike_slab_update = orig
torch.cuda.synchronize()
args = captured["args"]
Nn, Bb, J, h, _, perm, us, z, do_scan, a0 = args[:10]
prior = captured["prior"]
print("density of the captured state: %.3f" % float(a0.float().mean()), flush=True)


def run(variant, n_loc, debug=False):
    os.environ["PYGLM_SS_VARIANT"] = variant
    pr = {k: v[:n_loc].contiguous() for k, v in prior.items()}
    times = []
    out = None
    for rep in range(6):
        a = a0[:n_loc].clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        W, b, lo, ml, st = orig(Nn, Bb, J[:n_loc], h[:n_loc], pr, perm[:n_loc], us[:n_loc], z[:n_loc], do_scan[:n_loc], a,
                                want_logodds=True, want_ml=True)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            times.append(e0.elapsed_time(e1))
        out = (a.cpu().numpy(), W.cpu().numpy(), b.cpu().numpy(), lo.cpu().numpy(), ml.cpu().numpy(), int(st.abs().sum()))
    return float(np.median(times)), out


res = []
for n_loc in [int(x) for x in os.environ.get("PROBE_NLOC", "200,100,50,25").split(",")]:
    t_old, o_old = run("1", n_loc)
    row = dict(n_loc=n_loc, one_cta_ms=t_old)
    for v in os.environ.get("PROBE_VARIANTS", "84,88").split(","):
        t_new, o_new = run(v, n_loc)
        row[("cluster%s_ms" % v[1]) if len(v) > 1 else ("variant%s_ms" % v)] = t_new
        row["adjacency_equal"] = bool(np.array_equal(o_old[0], o_new[0]))
        row["max_abs_dW"] = float(np.max(np.abs(o_old[1] - o_new[1])))
        row["max_abs_dlogodds"] = float(np.max(np.abs(o_old[3] - o_new[3])))
        row["max_rel_dml"] = float(np.max(np.abs(o_old[4] - o_new[4]) / np.abs(o_old[4])))
        row["status"] = [o_old[5], o_new[5]]
    print(json.dumps(row), flush=True)
    res.append(row)
