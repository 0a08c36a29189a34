"""GPU probe (round 2): the tcgen05 Gram with resident Z digit planes against the streaming kernel that builds the Z
tiles in shared memory, at the shapes that matter -- cfg3 (N=200, B=2, T=1e5), one rank's slab of cfg4 (N=100, B=3,
T=1.25e6) and one rank's neuron block of cfg5 (N=1000, B=1, 125 neurons, T=1e6) -- plus the FP64 DMMA kernel on a time
prefix for scale.  Prints one JSON line per shape.

    python profiles/probe_gram_stream.py [--shapes cfg3,cfg4r,cfg5r] [--reps 5]
"""
import argparse
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pyglm_b200.kernels import CudaKernels, pad_ldn, gram_tc_bytes  # noqa: E402
from pyglm_b200.utils.basis import cosine_basis  # noqa: E402

SHAPES = {
    "cfg3": dict(N=200, B=2, T=100000, n=200),
    "cfg3n25": dict(N=200, B=2, T=100000, n=25),        # one of 8 neuron-sharded ranks
    "cfg4r": dict(N=100, B=3, T=1250000, n=100),
    "cfg5r": dict(N=1000, B=1, T=1000000, n=125),
    "cfg2": dict(N=27, B=3, T=100000, n=27),
}


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="cfg3,cfg4r,cfg5r")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    K = CudaKernels()
    for name in args.shapes.split(","):
        sh = SHAPES[name]
        N, B, T, n = sh["N"], sh["B"], sh["T"], sh["n"]
        D = N * B + 1
        L = 100
        rng = np.random.default_rng(0)
        basis = K.to_device(cosine_basis(B=B, L=L) / L)
        # spikes generated on the device in slices (cfg5: 8 GB of float64)
        g = torch.Generator(device=K.device)
        g.manual_seed(0)
        Y = (torch.rand(T, N, generator=g, device=K.device, dtype=torch.float32) < 0.05).to(torch.float64)
        Xp = K.filter_spikes(Y, basis, True)
        del Y
        om = K.zeros(T, pad_ldn(n))
        om[:, :n] = 0.02 + 0.4 * torch.rand(T, n, generator=g, device=K.device, dtype=torch.float64) ** 3
        flop = float(n) * T * D * (D + 1)
        out = dict(shape=name, N=N, B=B, T=T, n_local=n, D=D, algorithmic_flop=flop)
        free = torch.cuda.mem_get_info()[0]
        plans = {}
        if gram_tc_bytes(D, n, T, 4) < 0.85 * free:
            plans["resident"] = K.gram_tc_plan(Xp, D, n, 4)
        plans["stream"] = K.gram_tc_plan(Xp, D, n, 4, stream=True)
        Jints = {}
        for mode, plan in plans.items():
            plan.slice_omega(om)
            ms = timed(plan.mma, args.reps)
            Jints[mode] = plan.Jint.clone()
            out[mode + "_ms"] = ms
            out[mode + "_int8_tops"] = flop * 10 / (ms * 1e-3) / 1e12
            out[mode + "_fp64_equiv_tflops"] = flop / (ms * 1e-3) / 1e12
        if len(Jints) == 2:
            out["stream_equals_resident"] = bool(torch.equal(Jints["resident"], Jints["stream"]))
        # FP64 kernel on a prefix (and the deviation of the streaming result from it on that prefix)
        Tp = min(T, 100000)
        pp = K.gram_tc_plan(Xp[:Tp].contiguous(), D, min(n, 8), 4, stream=True)
        omp = om[:Tp, :64].contiguous()
        J_tc = pp.gram(omp)
        ms64 = timed(lambda: K.weighted_gram(Xp[:Tp], omp, D, min(n, 8)), 2)
        J64 = K.weighted_gram(Xp[:Tp], omp, D, min(n, 8))
        tril = torch.tril(torch.ones(D, D, dtype=torch.bool, device=K.device))
        a, b = J_tc[:, :D, :D], J64[:, :D, :D]
        dev = torch.where(tril & (b != 0), (a - b).abs() / b.abs(), torch.zeros_like(b))
        out["max_rel_dev_vs_fp64_prefix"] = float(dev.max())
        out["fp64_kernel_tflops_prefix"] = float(min(n, 8)) * Tp * D * (D + 1) / (ms64 * 1e-3) / 1e12
        print(json.dumps(out), flush=True)
        del plans, Jints, Xp, om, pp
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
