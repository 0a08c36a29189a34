"""A/B of the overlapped and the one-block single-GPU sweep in ONE process, alternating: ms per sweep over blocks of 16
sweeps through the engine call and through resample_model(), with and without the periodic spot checks of the Gram.
    python profiles/probe_sweep_times.py [blocks]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from bench import synthetic_spikes
from pyglm_b200.models import SparseBernoulliGLM
from pyglm_b200.utils.basis import cosine_basis

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 3
N, T, B, L = 200, 100000, 2, 100
np.random.seed(0)
m = SparseBernoulliGLM(N, basis=cosine_basis(B=B, L=L) / L, regression_kwargs=dict(S_w=10.0, mu_b=-2.), seed=1234)
m.add_data(synthetic_spikes(T, N), host_X=False)
eng = m.engine
ds = m._device_datasets()[0]


def block(mode, api, n=16):
    eng.overlap = mode == "overlap"
    if api == "model":
        for _ in range(3):
            m.resample_model()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ts = []
        for _ in range(n):
            t1 = time.perf_counter()
            m.resample_model()
            ts.append((time.perf_counter() - t1) * 1e3)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / n, ts
    A, W, b = m._host_state()
    hyp = m._stacked_hypers()
    for _ in range(3):
        A, W, b = eng.sweep([ds], A, W, b, hyp)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ts = []
    for _ in range(n):
        t1 = time.perf_counter()
        A, W, b = eng.sweep([ds], A, W, b, hyp)
        ts.append((time.perf_counter() - t1) * 1e3)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3 / n
    for k in range(N):                       # hand the engine's state back to the model
        m.regressions[k].a, m.regressions[k].W, m.regressions[k].b = A[k], W[k], b[k:k + 1]
    return dt, ts


for recheck in (16, 0):
    eng.TC_RECHECK_EVERY = recheck
    for rep in range(blocks):
        for api in ("engine", "model"):
            for mode in ("overlap", "one-block"):
                dt, ts = block(mode, api)
                print("recheck=%-2d %-6s %-9s %.2f ms/sweep | %s" % (recheck, api, mode, dt, " ".join("%.1f" % t for t in ts)))
