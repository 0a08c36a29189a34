"""world_size-2 gloo tests of the multi-GPU host logic on CPU: the sharded engines must reproduce the
single-process sweep.  The compute backend here is the oracle stand-in (tests/oracle_kernels.py); the CUDA path
of the same logic is covered by tests/test_multigpu.py on the GPU box."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    from oracle import pyglm_oracle as O
    N, B, L, T = 5, 2, 10, 600
    basis = O.cosine_basis(B, L) / L
    Y = (np.random.default_rng(3).random((T, N)) < 0.1).astype(np.float64)
    return N, B, basis, Y


def _run_chain(comm, shard, n_sweeps=2, prior=None):
    """Build the model on this rank and run a short chain; identical host RNG seeding on every rank -- except for the
    network prior when one with latent state is asked for: it is built from a rank-dependent seed, and the model
    must hand out rank 0's state."""
    from pyglm_b200 import networks
    from pyglm_b200.engine import GibbsEngine
    from pyglm_b200.models import SparseBernoulliGLM
    from tests.oracle_kernels import OracleKernels
    N, B, basis, Y = _problem()
    np.random.seed(100 + comm.rank)
    net = None if prior is None else getattr(networks, prior)(N, B)
    np.random.seed(0)
    m = SparseBernoulliGLM(N, basis=basis, network=net, regression_kwargs=dict(S_w=10.0, mu_b=-2.0, rho=0.4), seed=77)
    m._engine = GibbsEngine(N, B, kernels=OracleKernels(), seed=77, comm=comm, shard=shard)
    m.add_data(Y, host_X=False)
    lls = []
    for _ in range(n_sweeps):
        m.resample_model()
        lls.append(m.log_likelihood())
    return m.adjacency, m.weights, m.biases, np.array(lls), m.means[0] if shard == "neuron" else None


def _worker(rank, world, port, shard, out, prior=None, n_sweeps=2):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pyglm_b200.distributed import Comm
        A, W, b, lls, mu = _run_chain(Comm(), shard, n_sweeps=n_sweeps, prior=prior)
        if rank == 0:
            np.savez(out, A=A, W=W, b=b, lls=lls, mu=mu if mu is not None else np.zeros(0))
        # every rank must end with the same state
        t = torch.from_numpy(np.concatenate([A.ravel().astype(float), W.ravel(), b]))
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(t, ref)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shard", ["neuron", "time"])
def test_two_rank_sweep_matches_single_process(tmp_path, shard):
    from pyglm_b200.distributed import Comm
    A0, W0, b0, lls0, mu0 = _run_chain(Comm(), "neuron")
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), shard, out), nprocs=2, join=True)
    g = np.load(out)
    assert np.array_equal(g["A"], A0)
    tol = dict(rtol=0, atol=0) if shard == "neuron" else dict(rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(g["W"], W0, **tol)
    np.testing.assert_allclose(g["b"], b0, **tol)
    np.testing.assert_allclose(g["lls"], lls0, rtol=1e-12)
    if shard == "neuron":
        np.testing.assert_allclose(g["mu"], mu0, rtol=1e-12)


@pytest.mark.parametrize("prior", ["NIWStochasticBlockNetwork", "NIWLatentDistanceNetwork",
                                   # the full structured models: block / distance dependent weights as well
                                   "StochasticBlockNetwork", "LatentDistanceNetwork"])
def test_two_rank_chain_with_stateful_network_prior(tmp_path, prior):
    """Block labels / latent locations persist from sweep to sweep: all ranks start from rank 0's network state and
    draw the host step from rank 0's numpy stream, so the sharded chain is the single-process chain."""
    from pyglm_b200.distributed import Comm
    A0, W0, b0, lls0, _ = _run_chain(Comm(), "neuron", n_sweeps=4, prior=prior)
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), "neuron", out, prior, 4), nprocs=2, join=True)
    g = np.load(out)
    assert np.array_equal(g["A"], A0)
    np.testing.assert_allclose(g["W"], W0, rtol=0, atol=0)
    np.testing.assert_allclose(g["lls"], lls0, rtol=1e-12)


def test_partitions():
    from pyglm_b200.distributed import block_partition, time_partition
    for N, world in [(200, 8), (27, 4), (5, 2), (3, 8), (1000, 8)]:
        blocks = [block_partition(N, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and max(b[1] for b in blocks) == N
        assert all(b[1] - b[0] <= b[2] for b in blocks)
        covered = sorted(i for b in blocks for i in range(b[0], b[1]))
        assert covered == list(range(N))
    for T, world in [(100000, 8), (7, 4)]:
        slabs = [time_partition(T, world, r) for r in range(world)]
        assert sorted(i for s in slabs for i in range(*s)) == list(range(T))
