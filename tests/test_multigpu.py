"""Two-GPU (NCCL) runs of the CUDA path: neuron-sharded and time-sharded sweeps must reproduce the single-GPU
chain (the Philox streams are keyed by global (time, neuron) indices, so the draws do not depend on sharding).
Skipped unless at least two GPUs are visible."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _chain(comm, shard, n_sweeps=3, gram="auto", prior=None, N=10, T=5000):
    from pyglm_b200 import networks
    from pyglm_b200.models import SparseBernoulliGLM
    from pyglm_b200.utils.basis import cosine_basis
    B, L = 2, 20
    basis = cosine_basis(B, L) / L
    Y = (np.random.default_rng(3).random((T, N)) < 0.08).astype(np.float64)
    # a prior with latent state is built from a rank-dependent seed: the model has to hand out rank 0's
    np.random.seed(0 if prior is None else 100 + comm.rank)
    net = None if prior is None else getattr(networks, prior)(N, B)
    np.random.seed(0)
    m = SparseBernoulliGLM(N, basis=basis, network=net, regression_kwargs=dict(S_w=10.0, mu_b=-2.0), seed=77,
                           comm=comm, shard=shard, gram=gram)
    m.add_data(Y, host_X=False)
    lls = []
    for _ in range(n_sweeps):
        m.resample_model()
        lls.append(m.log_likelihood())
    return m.adjacency, m.weights, m.biases, np.array(lls)


def _worker(rank, world, port, shard, out, gram="auto", prior=None, N=10, T=5000, peer="1"):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["PYGLM_PEER_EXCHANGE"] = peer
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from pyglm_b200.distributed import Comm
        A, W, b, lls = _chain(Comm(), shard, gram=gram, prior=prior, N=N, T=T)
        if rank == 0:
            np.savez(out, A=A, W=W, b=b, lls=lls)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shard", ["neuron", "time"])
def test_two_gpu_chain_matches_single_gpu(tmp_path, shard):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from pyglm_b200.distributed import Comm
    A0, W0, b0, lls0 = _chain(Comm(), "neuron")
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), shard, out), nprocs=2, join=True)
    g = np.load(out)
    assert np.array_equal(g["A"], A0)
    np.testing.assert_allclose(g["W"], W0, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(g["b"], b0, rtol=1e-8)
    np.testing.assert_allclose(g["lls"], lls0, rtol=1e-9)


def test_two_gpu_time_sharded_tensor_core_gram(tmp_path):
    """Time-sharded psi / PG / tcgen05 Gram with the exact int64 reduce-scatter of the integer partial sums: J is the
    single-GPU J bit for bit (same digits, same integer sums, same draws); h = X~^T kappa is an FP64 all-reduce, so
    the chain agrees with the 1-GPU tensor-core chain to round-off, with identical adjacency."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from pyglm_b200.distributed import Comm
    A0, W0, b0, lls0 = _chain(Comm(), "neuron", gram="tc")
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), "time", out, "tc"), nprocs=2, join=True)
    g = np.load(out)
    assert np.array_equal(g["A"], A0)
    np.testing.assert_allclose(g["W"], W0, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(g["b"], b0, rtol=1e-9)
    np.testing.assert_allclose(g["lls"], lls0, rtol=1e-11)


@pytest.mark.parametrize("prior", ["NIWStochasticBlockNetwork", "NIWLatentDistanceNetwork"])
def test_two_gpu_chain_with_stateful_network_prior(tmp_path, prior):
    """Adjacency priors that carry latent state (block labels, locations): every rank draws the host step itself from
    rank 0's numpy stream, starting from rank 0's network state, so the 2-GPU chain is the single-GPU chain."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from pyglm_b200.distributed import Comm
    A0, W0, b0, lls0 = _chain(Comm(), "neuron", prior=prior)
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), "neuron", out, "auto", prior), nprocs=2, join=True)
    g = np.load(out)
    assert np.array_equal(g["A"], A0)
    np.testing.assert_allclose(g["W"], W0, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(g["lls"], lls0, rtol=1e-9)


@pytest.mark.parametrize("peer", ["1", "0"], ids=["peer-memory", "nccl"])
def test_two_gpu_time_sharded_cluster_scan_and_both_exchange_paths(tmp_path, peer):
    """N = 40, B = 2 (D = 81): large enough for the CLUSTER scan kernel (csrc/spike_slab_dsm.cu) on both the single-GPU
    run and the 20-neuron scan blocks of the two ranks, with the tensor-core Gram time-sharded.  Run once with the
    exchanges as our own kernels over peer-mapped memory (reduce-scatter fused into the finalize pass, state rows pushed
    over NVLink) and once with the NCCL collectives: both must reproduce the single-GPU chain."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from pyglm_b200.distributed import Comm
    N, T = 40, 20000
    A0, W0, b0, lls0 = _chain(Comm(), "neuron", gram="tc", N=N, T=T)
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), "time", out, "tc", None, N, T, peer), nprocs=2, join=True)
    g = np.load(out)
    assert np.array_equal(g["A"], A0)
    np.testing.assert_allclose(g["W"], W0, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(g["b"], b0, rtol=1e-8)
    np.testing.assert_allclose(g["lls"], lls0, rtol=1e-10)
