"""Duration of the tcgen05 Gram MMA kernel at cfg3's D = 401, n = 200 as a function of the slab length T
(time-sharded runs hand each rank T/world bins): kernel time should scale with T.

    python profiles/probe_tc_tsweep.py [--Ts 100000,50000,...]
"""
import argparse
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pyglm_b200.kernels import CudaKernels, pad_ldn  # noqa: E402
from pyglm_b200.utils.basis import cosine_basis  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--Ts", default="100000,50000,49984,50112,25000,12500,12480")
ap.add_argument("--N", type=int, default=200)
ap.add_argument("--B", type=int, default=2)
ap.add_argument("--n", type=int, default=200)
a = ap.parse_args()
K = CudaKernels(torch.device("cuda", 0))
N, B, n = a.N, a.B, a.n
D = N * B + 1
rng = np.random.default_rng(0)
basis = K.to_device(cosine_basis(B=B, L=100) / 100)
for T in [int(t) for t in a.Ts.split(",")]:
    Y = (rng.random((T, N)) < 0.05).astype(np.float64)
    Xp = K.filter_spikes(K.to_device(Y), basis, True)
    om = K.zeros(T, pad_ldn(n))
    om[:, :n] = torch.from_numpy(0.02 + 0.4 * rng.random((T, n)) ** 3).to(K.device)
    plan = K.gram_tc_plan(Xp, D, n, 4)
    plan.slice_omega(om)
    for _ in range(2):
        plan.mma()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        plan.mma()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    g = plan.geom
    print(json.dumps(dict(T=T, Tpad=g["Tpad"], n_chunks=g["n_chunks"], bpc=g["blocks_per_chunk"], ms=ms,
                          ms_per_1e5_bins=ms * 1e5 / T)), flush=True)
    del plan, Xp, om
    torch.cuda.empty_cache()
