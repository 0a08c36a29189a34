"""Basis construction (host) and the filtered-spike-train build (device).

cosine_basis mirrors pyglm/utils/basis.py:61-106 and stays on the host (an L x B table built once).
convolve_with_basis mirrors pyglm/utils/basis.py:5-34 but runs the causal-convolution CUDA kernel
(csrc/filter.cu) instead of per-column FFTs.
"""
import numpy as np
import scipy.linalg


def cosine_basis(B, L=100, orth=False, norm=True, n_eye=0, a=1.0 / 120, b=0.5):
    """L x B table of raised-cosine bumps whose centres are spread linearly in log time.

    The first n_eye columns are unit impulses at lags 1..n_eye; the remaining B - n_eye columns are
    (cos(clip((u - c_i) pi / (2 w), -pi, pi)) + 1) / 2 with u = log(a t + b).  With norm=True every column is
    rescaled to sum to L (callers divide by L, cf. examples/synthetic.py:24); with orth=True the columns are
    orthonormalised instead."""
    n_cos = B - n_eye
    assert n_cos >= 0 and n_eye >= 0
    table = np.zeros((L, B))
    table[np.arange(n_eye), np.arange(n_eye)] = 1.0
    u = np.log(a * np.arange(L) + b)
    centre_bins = np.floor(np.linspace(n_eye, L / 2.0, n_cos)).astype(int)
    centres = u[centre_bins]
    width = centres / 2 if n_cos == 1 else (centres[-1] - centres[0]) / (n_cos - 1)
    for i in range(n_cos):
        phase = np.clip((u - centres[i]) * np.pi / width / 2.0, -np.pi, np.pi)
        table[:, n_eye + i] = 0.5 * (np.cos(phase) + 1.0)
    if orth:
        return scipy.linalg.orth(table)
    if norm:
        if np.any(table < 0):
            raise Exception("We can only normalize nonnegative impulse responses!")
        table = table / table.sum(axis=0, keepdims=True) * L
    return table


def interpolate_basis(basis, dt, dt_max, norm=True, allow_instantaneous=False):
    """Resample an (L, B) basis defined on [0, dt_max] at bin width dt (pyglm/utils/basis.py:36-58; host, unused by
    the sampler).  norm: every column integrates to one; a zero row is prepended unless allow_instantaneous."""
    L, B = basis.shape
    t_new = np.arange(0.0, dt_max, step=dt)
    t_old = np.linspace(0.0, dt_max, L)
    out = np.column_stack([np.interp(t_new, t_old, basis[:, b]) for b in range(B)])
    if norm:
        out = out / (dt * out.sum(axis=0))
    return out if allow_instantaneous else np.vstack((np.zeros((1, B)), out))


def convolve_with_basis(S, basis):
    """X[t, n, b] = sum_{l=1..L} basis[l-1, b] * S[t-l, n]  ->  (T, N, B) float64 host array.

    Strictly causal (basis row 0 is lag 1), zero history before t = 0, clipped at zero when both inputs are
    non-negative, exactly as pyglm/utils/basis.py:5-34.  Runs on the current CUDA device."""
    from ..engine import default_kernels
    S = np.ascontiguousarray(S, dtype=np.float64)
    basis = np.ascontiguousarray(basis, dtype=np.float64)
    assert S.ndim == 2 and basis.ndim == 2
    T, N = S.shape
    B = basis.shape[1]
    K = default_kernels()
    clip = bool(np.amin(basis) >= 0 and np.amin(S) >= 0) if S.size else False
    if T == 0:
        return np.empty((0, N, B))
    Xp = K.filter_spikes(K.to_device(S), K.to_device(basis), clip)
    return K.unpack_design(Xp, N * B).cpu().numpy().reshape(T, N, B)
