"""CPU-side checks of the C-ABI library and the host logic (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from pyglm_b200.build import build_library
    from pyglm_b200 import cabi
    build_library()
    return cabi.load()


def test_every_header_symbol_is_exported_and_bound(lib):
    from pyglm_b200 import cabi
    header = open(os.path.join(ROOT, "include", "pyglm_b200.h")).read()
    declared = set(re.findall(r"\b(pyglm_[a-z0-9_]+)\s*\(", header))
    declared.discard("pyglm_stream_t")
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name
    assert declared == set(cabi.SIGNATURES), "ctypes table and header disagree: %s" % (declared ^ set(cabi.SIGNATURES))
    assert lib.pyglm_abi_version() == 1


def test_gram_tile_enumeration_covers_lower_triangle(lib):
    for D in (1, 5, 8, 9, 33, 82, 401):
        n = lib.pyglm_gram_tiles(D, 0, None, 0)
        buf = np.zeros((n, 2), dtype=np.int32)
        assert lib.pyglm_gram_tiles(D, 0, buf.ctypes.data_as(ctypes.c_void_p), n) == n
        cover = np.zeros((D + 40, D + 40), dtype=bool)
        for i0, j0 in buf:
            assert i0 % 8 == 0 and j0 % 32 == 0 and j0 <= i0 + 7
            mcount = min(4, (i0 + 7 - j0) // 8 + 1)
            cover[i0:i0 + 8, j0:j0 + 8 * mcount] = True
        ii, jj = np.tril_indices(D)
        assert cover[ii, jj].all()
        # mode 1: only the row block of the bias row
        n1 = lib.pyglm_gram_tiles(D, 1, None, 0)
        b1 = np.zeros((n1, 2), dtype=np.int32)
        lib.pyglm_gram_tiles(D, 1, b1.ctypes.data_as(ctypes.c_void_p), n1)
        assert set(b1[:, 0]) == {((D - 1) // 8) * 8}
    assert lib.pyglm_gram_slabs(351, 200, 100000) == 1
    assert lib.pyglm_gram_slabs(21, 27, 100000) > 1


def test_argument_errors_are_reported_without_a_gpu(lib):
    rc = lib.pyglm_filter_spikes(None, None, 10, 2, 5, 1, 0, None, 32, None)
    assert rc != 0 and b"null pointer" in lib.pyglm_last_error()
    rc = lib.pyglm_pg_draw(None, 8, 10, 4, None, 8, 1, 1, 0, 0, 4, None)
    assert rc != 0 and b"null pointer" in lib.pyglm_last_error()


def test_no_cpu_fallback():
    """Without a CUDA device the compute backend must refuse to construct (this container has no GPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pyglm_b200 import cabi
    from pyglm_b200.kernels import CudaKernels
    with pytest.raises(cabi.PyglmCudaError):
        CudaKernels()
    from pyglm_b200.utils.basis import convolve_with_basis
    with pytest.raises(RuntimeError):
        convolve_with_basis(np.zeros((10, 2)), np.eye(3))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pyglm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_host_api_without_gpu(golden):
    """Construction, hyper-parameter expansion, state views, cosine_basis and the prior terms are host logic."""
    from pyglm_b200.models import SparseBernoulliGLM
    from pyglm_b200.priors import prior_arrays
    from pyglm_b200.utils.basis import cosine_basis
    from oracle import pyglm_oracle as O
    g = golden("basis.npz")
    for key in g.files:
        B, L = (int(s[1:]) for s in key.split("_"))
        np.testing.assert_allclose(cosine_basis(B, L), g[key], rtol=1e-14, atol=1e-16)
    np.random.seed(0)
    N, B = 4, 2
    m = SparseBernoulliGLM(N, basis=cosine_basis(B, 20) / 20, regression_kwargs=dict(S_w=10.0, mu_b=-2.))
    r = m.regressions[0]
    assert r.rho.shape == (N,) and r.mu_w.shape == (N, B) and r.S_w.shape == (N, B, B) and r.S_b.shape == (1, 1)
    assert m.weights.shape == (N, N, B) and m.adjacency.dtype == bool and m.biases.shape == (N,)
    assert np.all(m.weights[~m.adjacency] == 0)
    with pytest.raises(AssertionError):
        r.S_b = np.eye(1)                      # S_b must be a scalar (regression.py:135)
    # SURVEY 5's keyword extras: precision= is an alias of gram=, device= picks the engine's GPU (nothing is created here)
    m2 = SparseBernoulliGLM(N, basis=cosine_basis(B, 20) / 20, precision="int8", device=0, seed=3)
    assert m2._gram == "tc" and m2._device == 0 and m2._engine is None
    assert SparseBernoulliGLM(N, basis=cosine_basis(B, 20) / 20, precision="fp64")._gram == "fp64"
    with pytest.raises(AssertionError):
        SparseBernoulliGLM(N, basis=cosine_basis(B, 20) / 20, precision="fp64", gram="tc")
    # prior terms against the oracle's dense prior statistics
    rng = np.random.default_rng(1)
    S_w = np.stack([[np.eye(B) * rng.uniform(0.5, 3) + 0.1 for _ in range(N)] for _ in range(3)])
    mu_w = rng.standard_normal((3, N, B))
    rho = rng.uniform(0.1, 0.9, (3, N))
    rho[2] = [0, 1, 1, 0]
    mub, Sb = [0.5, -2, 0], [1.0, 2.0, 0.5]
    pr = prior_arrays(rho, mu_w, S_w, np.array(mub), np.array(Sb))
    assert list(pr["do_scan"]) == [True, True, False]
    for j in range(3):
        J0, h0 = O.prior_sufficient_statistics(mu_w[j], S_w[j], np.array([mub[j]]), np.array([[Sb[j]]]))
        for mth in range(N):
            blk = slice(mth * B, (mth + 1) * B)
            np.testing.assert_allclose(pr["J0w"][j, mth], J0[blk, blk], rtol=1e-13)
            np.testing.assert_allclose(pr["h0w"][j, mth], h0[blk], rtol=1e-13)
            c = 0.5 * np.linalg.slogdet(J0[blk, blk])[1] - 0.5 * h0[blk] @ np.linalg.solve(J0[blk, blk], h0[blk])
            assert pr["cprior"][j, mth] == pytest.approx(c, rel=1e-12, abs=1e-14)
        assert pr["J0b"][j] == pytest.approx(J0[-1, -1]) and pr["h0b"][j] == pytest.approx(h0[-1])


@pytest.mark.parametrize("B", [1, 2, 3])
def test_prior_terms_closed_forms_equal_the_lapack_route(B):
    """priors.prior_arrays takes closed forms on contiguous planes for B <= 2 (it runs on the host between two sweeps)
    and numpy's batched inv / slogdet beyond: both must give regression.py:138-151, 210-223 for non-isotropic blocks."""
    from pyglm_b200.priors import prior_arrays
    rng = np.random.default_rng(B)
    n, N = 4, 7
    M = rng.standard_normal((n, N, B, B))
    S_w = M @ M.transpose(0, 1, 3, 2) + 0.3 * np.eye(B)
    mu_w = rng.standard_normal((n, N, B))
    rho = rng.uniform(0.05, 0.95, (n, N))
    rho[1] = 1.0                                             # a deterministic row: no scan (regression.py:153-155)
    pr = prior_arrays(rho, mu_w, S_w, rng.standard_normal(n), rng.uniform(0.5, 2.0, n))
    J = np.linalg.inv(S_w)
    h = np.einsum("nmbc,nmc->nmb", J, mu_w)
    sign, logdet = np.linalg.slogdet(J)
    np.testing.assert_allclose(pr["J0w"], J, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(pr["h0w"], h, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(pr["cprior"], 0.5 * logdet - 0.5 * np.einsum("nmb,nmb->nm", mu_w, h), rtol=1e-12, atol=1e-13)
    assert list(pr["do_scan"]) == [True, False, True, True]
    assert pr["J0w"].flags.c_contiguous and pr["h0w"].flags.c_contiguous
    bad = S_w.copy()
    bad[2, 3] = -np.eye(B)
    with pytest.raises(ValueError):
        prior_arrays(rho, mu_w, bad, np.zeros(n), np.ones(n))
