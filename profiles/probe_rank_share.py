"""Per-rank sweep time of an N-GPU neuron-sharded run, measured on ONE GPU: rank r of `world` does exactly the work
it would do in the real job (its block of postsynaptic neurons); the all-gather is replaced by a local copy of the
same size, so what is measured is the compute + host part of a rank's sweep (the NCCL exchange is 0.64 MB at cfg3).
Used to find what limits strong scaling without spending 8 GPUs on it.

    python profiles/probe_rank_share.py [--config cfg3] [--worlds 1,2,4,8] [--steps 5]
"""
import argparse
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import CONFIGS, synthetic_spikes  # noqa: E402
from pyglm_b200.distributed import Comm  # noqa: E402
from pyglm_b200.models import SparseBernoulliGLM  # noqa: E402
from pyglm_b200.utils.basis import cosine_basis  # noqa: E402


class FakeComm(Comm):
    def __init__(self, world, rank):
        self.enabled, self.group, self.world, self.rank = False, None, world, rank

    def all_gather_rows(self, local):
        return local.repeat((self.world,) + (1,) * (local.dim() - 1))

    def reduce_scatter_rows(self, full):
        n_max = full.shape[0] // self.world
        return full[self.rank * n_max:(self.rank + 1) * n_max].clone()

    def all_reduce_sum(self, t):
        return t

    def all_reduce_max(self, t):
        return t

    def all_reduce_min(self, t):
        return t

    def broadcast_object(self, obj, src=0):
        return obj

    def barrier(self):
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--worlds", default="1,2,4,8")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--gram", default="auto")
    ap.add_argument("--shard", default="neuron")
    ap.add_argument("--pipeline", type=int, default=1)
    ap.add_argument("--network", default=None, help="class in pyglm_b200.networks, e.g. NIWLatentDistanceNetwork")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    N, B, L, T = cfg["N"], cfg["B"], cfg["L"], cfg["T"]
    if T * N > 4e8:        # cfg5: 8 GB of float64 spikes -- draw them in slabs
        Y = np.concatenate([synthetic_spikes(T // 10, N, seed=s) for s in range(10)])
    else:
        Y = synthetic_spikes(T, N)
    for world in [int(w) for w in a.worlds.split(",")]:
        np.random.seed(0)
        comm = FakeComm(world, 0) if world > 1 else None
        net = None
        if a.network:
            from pyglm_b200 import networks
            net = getattr(networks, a.network)(N, B)
        model = SparseBernoulliGLM(N, basis=cosine_basis(B=B, L=L) / L, regression_kwargs=dict(S_w=10.0, mu_b=-2.),
                                   seed=1234, gram=a.gram, comm=comm, shard=a.shard, network=net)
        model.add_data(Y, host_X=False)
        model.engine.pipeline = bool(a.pipeline)
        if world > 1 and a.shard == "time":
            # one GPU cannot form the reduce-scattered totals the first-sweep accuracy check compares (it would judge
            # the slab's partial sums, which is stricter than the real job): accept the 4-digit plan as the job does
            model.engine.TC_ACCEPT = 1e-7
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
        ph = eng.phase_ms()
        ms = e0.elapsed_time(e1) / a.steps
        import time
        t0 = time.perf_counter()
        model.resample_network()
        host_net_ms = (time.perf_counter() - t0) * 1e3
        print(json.dumps(dict(world=world, shard=a.shard, network=a.network or "NIWSparseNetwork", host_network_step_ms=host_net_ms, pipeline=a.pipeline, n_loc=eng.scan_hi - eng.scan_lo, ms_per_sweep_e2e=ms, phases=ph,
                              other_ms=ms - sum(v for k, v in ph.items() if not k.startswith("gram_")),
                              density=float(model.adjacency.mean()))), flush=True)
        del model, eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
