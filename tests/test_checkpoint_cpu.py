"""Checkpoint / resume of a chain (optional row): state_dict() holds the sampler state, the hyper-parameters,
the network's latent state, the engine's Philox seed + sweep counter and numpy's RNG state; a chain continued from it
-- in the same model or in a freshly built one -- is the uninterrupted chain bit for bit.  Host logic on CPU with the
oracle stand-in kernels (tests/oracle_kernels.py); the CUDA path of the same calls is in tests/test_model_gpu.py."""
import pickle

import numpy as np
import pytest


def _model(prior, seed_np=0):
    from oracle import pyglm_oracle as O
    from pyglm_b200 import networks
    from pyglm_b200.distributed import Comm
    from pyglm_b200.engine import GibbsEngine
    from pyglm_b200.models import SparseBernoulliGLM
    from tests.oracle_kernels import OracleKernels
    N, B, L, T = 5, 2, 10, 500
    basis = O.cosine_basis(B, L) / L
    Y = (np.random.default_rng(3).random((T, N)) < 0.1).astype(np.float64)
    np.random.seed(seed_np)
    net = None if prior is None else getattr(networks, prior)(N, B)
    m = SparseBernoulliGLM(N, basis=basis, network=net, regression_kwargs=dict(S_w=10.0, mu_b=-2.0, rho=0.4), seed=5)
    m._engine = GibbsEngine(N, B, kernels=OracleKernels(), seed=5, comm=Comm(), shard="neuron")
    m.add_data(Y, host_X=False)
    return m


def _run(m, n):
    out = []
    for _ in range(n):
        m.resample_model()
        out.append((m.adjacency.copy(), m.weights.copy(), m.biases.copy(), m.network.rho.copy(), m.network.mu_W.copy()))
    return out


@pytest.mark.parametrize("prior", [None, "NIWStochasticBlockNetwork", "LatentDistanceNetwork"])
def test_resumed_chain_is_the_uninterrupted_chain(prior):
    m = _model(prior)
    _run(m, 2)
    sd = pickle.loads(pickle.dumps(m.state_dict()))           # survives a round trip through a file
    ref = _run(m, 3)
    # same model, wound back
    m.load_state_dict(sd)
    again = _run(m, 3)
    # a freshly built model with different seeds everywhere
    other = _model(prior, seed_np=123)
    other.engine.seed = 999
    other.load_state_dict(sd)
    fresh = _run(other, 3)
    for got in (again, fresh):
        for a, b in zip(ref, got):
            for x, y in zip(a, b):
                assert np.array_equal(x, y)


def test_state_dict_rejects_a_different_model():
    m = _model(None)
    sd = m.state_dict()
    sd["N"] = 6
    with pytest.raises(AssertionError):
        m.load_state_dict(sd)
