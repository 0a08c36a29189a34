"""Host-side preparation of the Gaussian / Bernoulli prior terms consumed by the spike-and-slab kernel.

Follows regression.py:138-151 (natural_params) and :210-223 (_prior_sufficient_statistics), batched over
postsynaptic neurons and kept in block form: the reference materialises a dense (D x D) block_diag per neuron
per sweep; the kernel only ever needs the B x B blocks, their closed-form log-normaliser and log-odds of rho.
"""
import numpy as np


def prior_arrays(rho, mu_w, S_w, mu_b, S_b):
    """rho (n,N), mu_w (n,N,B), S_w (n,N,B,B), mu_b (n,), S_b (n,)  ->  dict of float64 arrays
         J0w (n,N,B,B) = S_w^-1               h0w (n,N,B) = J0w mu_w
         J0b (n,) = 1/S_b                     h0b (n,) = J0b mu_b
         cprior (n,N) = 1/2 log|J0w| - 1/2 h0w^T J0w^-1 h0w   (prior part of ml(a_m=1) - ml(a_m=0))
         logit_rho (n,N) = log rho - log(1-rho)
         do_scan (n,) bool: False where rho is deterministic (regression.py:153-155)."""
    rho = np.asarray(rho, dtype=np.float64)
    mu_w = np.asarray(mu_w, dtype=np.float64)
    S_w = np.asarray(S_w, dtype=np.float64)
    mu_b = np.asarray(mu_b, dtype=np.float64).reshape(-1)
    S_b = np.asarray(S_b, dtype=np.float64).reshape(-1)
    B = S_w.shape[-1]
    if B <= 2:
        # closed forms on contiguous component planes: numpy's batched inv / slogdet pay a LAPACK call per B x B block
        # (40000 of them at cfg3), and arithmetic on the strided views S_w[..., i, j] is several times slower than on
        # planes -- this runs on the host between two sweeps, on the critical path of the overlapped sweep
        shp = S_w.shape[:-2]
        if B == 1:
            det = S_w.reshape(shp)
            J0w = 1.0 / S_w
            pos = det > 0
            h0w = J0w.reshape(shp + (1,)) * mu_w
            quad = (mu_w * h0w).reshape(shp)
        else:
            Sp = np.ascontiguousarray(S_w.reshape(-1, 4).T)              # planes s00, s01, s10, s11
            mp = np.ascontiguousarray(mu_w.reshape(-1, 2).T)             # planes mu0, mu1
            det = Sp[0] * Sp[3] - Sp[1] * Sp[2]
            pos = (det > 0) & (Sp[0] > 0)
            inv = 1.0 / det
            Jp = np.empty_like(Sp)
            np.multiply(Sp[3], inv, out=Jp[0])
            np.multiply(Sp[1], -inv, out=Jp[1])
            np.multiply(Sp[2], -inv, out=Jp[2])
            np.multiply(Sp[0], inv, out=Jp[3])
            hp = np.empty_like(mp)
            hp[0] = Jp[0] * mp[0] + Jp[1] * mp[1]
            hp[1] = Jp[2] * mp[0] + Jp[3] * mp[1]
            J0w = np.ascontiguousarray(Jp.T).reshape(S_w.shape)
            h0w = np.ascontiguousarray(hp.T).reshape(mu_w.shape)
            # h0w^T S_w h0w = mu_w^T J0w mu_w = mu_w . h0w
            quad = (mp[0] * hp[0] + mp[1] * hp[1]).reshape(shp)
            det = det.reshape(shp)
        if not np.all(pos):
            raise ValueError("S_w must be positive definite")
        logdet = -np.log(det)
    else:
        J0w = np.linalg.inv(S_w)
        h0w = np.einsum("nmbc,nmc->nmb", J0w, mu_w)
        sign, logdet = np.linalg.slogdet(J0w)
        if np.any(sign <= 0):
            raise ValueError("S_w must be positive definite")
        quad = np.einsum("nmb,nmbc,nmc->nm", h0w, S_w, h0w)
    J0b = 1.0 / S_b
    h0b = J0b * mu_b
    cprior = 0.5 * logdet - 0.5 * quad
    with np.errstate(divide="ignore"):
        logit_rho = np.log(rho) - np.log1p(-rho)
    do_scan = ~np.all((rho < 1e-6) | (rho > 1 - 1e-6), axis=1)
    return dict(J0w=np.ascontiguousarray(J0w), h0w=np.ascontiguousarray(h0w), J0b=J0b, h0b=h0b,
                cprior=np.ascontiguousarray(cprior), logit_rho=np.ascontiguousarray(logit_rho), do_scan=do_scan)
