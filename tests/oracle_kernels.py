"""TEST-ONLY stand-in for pyglm_b200.kernels.CudaKernels that evaluates every kernel with the CPU oracle on
CPU torch tensors.  It exists so the host-side sharding logic (partitioning, reduce-scatter of Gram partials,
all-gather of (a, W, b), the rank-0 network broadcast) can be exercised with world_size-2 gloo process groups
on a machine without a GPU.  It is never imported by the product package."""
import numpy as np
import torch

from oracle import pyglm_oracle as O
from pyglm_b200.kernels import pad_ldx


class OracleKernels(object):
    name = "oracle-cpu (tests only)"

    def __init__(self):
        self.device = torch.device("cpu")
        self.launches = 0

    def empty(self, *shape, dtype=torch.float64):
        return torch.zeros(*shape, dtype=dtype)

    zeros = empty

    def to_device(self, arr, dtype=None):
        t = torch.from_numpy(np.ascontiguousarray(arr)).clone()
        return t if dtype is None else t.to(dtype)

    # ---- (1)
    def filter_spikes(self, S, basis, clip):
        S, basis = S.numpy(), basis.numpy()
        T, N = S.shape
        B = basis.shape[1]
        X = O.convolve_direct(S, basis).reshape(T, N * B)
        if clip:
            X = np.maximum(X, 0)
        return self.pack_design(torch.from_numpy(X))

    def pack_design(self, X):
        T, NB = X.shape
        Xp = torch.zeros(T, pad_ldx(NB + 1), dtype=torch.float64)
        Xp[:, :NB] = X
        Xp[:, NB] = 1.0
        return Xp

    def unpack_design(self, Xp, NB):
        return Xp[:, :NB].clone()

    # ---- (5)
    def activation(self, Xp, Wt, D, n, out=None):
        psi = torch.zeros(Xp.shape[0], Wt.shape[1], dtype=torch.float64) if out is None else out
        psi[:, :n] = Xp[:, :D] @ Wt[:D, :n]
        return psi

    def loglik(self, Xp, Wt, D, n, Y, y_col0):
        psi = (Xp[:, :D] @ Wt[:D, :n]).numpy()
        y = Y[:, y_col0:y_col0 + n].numpy()
        return torch.tensor([float((y * psi - np.logaddexp(0, psi)).sum())], dtype=torch.float64)

    def means(self, Xp, Wt, D, n):
        return torch.sigmoid(Xp[:, :D] @ Wt[:D, :n])

    # ---- (2): per-element Philox streams keyed by the GLOBAL (t, n) index -> independent of the sharding
    def pg_draw(self, psi, n_valid, omega, seed, call_id, t_off, n_off, n_total):
        T = psi.shape[0]
        full = np.zeros((t_off + T) * n_total)
        idx = ((t_off + np.arange(T))[:, None] * n_total + n_off + np.arange(n_valid)[None, :]).ravel()
        full[idx] = psi[:, :n_valid].numpy().ravel()
        omega[:, :n_valid] = torch.from_numpy(O.pg1_draw(full, seed, call_id, rng_kind=0)[idx].reshape(T, n_valid))
        return omega

    # ---- (3)
    def weighted_gram(self, Xp, Om, D, n_valid, J=None, nslabs=None):
        ldx = Xp.shape[1]
        if J is None:
            J = torch.zeros(n_valid, ldx, ldx, dtype=torch.float64)
        X = Xp[:, :D]
        for j in range(n_valid):
            J[j, :D, :D] = torch.tril((X * Om[:, j:j + 1]).T @ X)
        return J

    def xt_kappa(self, Xp, kappa, D, n_valid):
        h = torch.zeros(n_valid, Xp.shape[1], dtype=torch.float64)
        h[:, :D] = (Xp[:, :D].T @ kappa[:, :n_valid]).T
        return h

    # ---- (4)
    def scan_randomness(self, N, B, n_loc, n_off, seed, call_id):
        D = N * B + 1
        perm = np.zeros((n_loc, N), dtype=np.int32)
        us, z = np.zeros((n_loc, N)), np.zeros((n_loc, D))
        for j in range(n_loc):
            rng = np.random.default_rng([seed, call_id, n_off + j])
            perm[j], us[j], z[j] = rng.permutation(N), rng.random(N), rng.standard_normal(D)
        return torch.from_numpy(perm), torch.from_numpy(us), torch.from_numpy(z)

    def spike_slab_update(self, N, B, J, h, prior, perm, us, z, do_scan, a, P_ws=None, want_logodds=False,
                          want_ml=False):
        n_loc = a.shape[0]
        D = N * B + 1
        W = torch.zeros(n_loc, N, B, dtype=torch.float64)
        bias = torch.zeros(n_loc, dtype=torch.float64)
        for j in range(n_loc):
            Jl = J[j, :D, :D].numpy()
            Jl = np.tril(Jl) + np.tril(Jl, -1).T
            J0 = np.zeros((D, D))
            for m in range(N):
                J0[m * B:(m + 1) * B, m * B:(m + 1) * B] = prior["J0w"][j, m].numpy()
            J0[-1, -1] = float(prior["J0b"][j])
            h0 = np.concatenate([prior["h0w"][j].numpy().ravel(), [float(prior["h0b"][j])]])
            aj = a[j].numpy().astype(bool)
            if int(do_scan[j]):
                lr = prior["logit_rho"][j].numpy()
                rho = 1.0 / (1.0 + np.exp(-lr))
                aj = O.collapsed_resample_a(J0, h0, J0 + Jl, h0 + h[j, :D].numpy(), aj, rho, B, perm[j].numpy(),
                                            us[j].numpy())
            Wj, bj = O.resample_W(J0 + Jl, h0 + h[j, :D].numpy(), aj, B, z[j].numpy()[O._mask(aj, B)])
            a[j] = torch.from_numpy(aj.astype(np.uint8))
            W[j], bias[j] = torch.from_numpy(Wj), float(bj[0])
        return W, bias, None, None, torch.zeros(n_loc, dtype=torch.int32)
