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
        # closed forms: numpy's batched inv / slogdet pay a LAPACK call per B x B block (40000 of them at cfg3)
        if B == 1:
            det = S_w[..., 0, 0]
            J0w = 1.0 / S_w
        else:
            s00, s01, s10, s11 = S_w[..., 0, 0], S_w[..., 0, 1], S_w[..., 1, 0], S_w[..., 1, 1]
            det = s00 * s11 - s01 * s10
            J0w = np.empty_like(S_w)
            J0w[..., 0, 0], J0w[..., 0, 1], J0w[..., 1, 0], J0w[..., 1, 1] = s11 / det, -s01 / det, -s10 / det, s00 / det
        if np.any(~(det > 0)) or np.any(~(S_w[..., 0, 0] > 0)):
            raise ValueError("S_w must be positive definite")
        logdet = -np.log(det)
        h0w = (J0w * mu_w[..., None, :]).sum(-1)
        # h0w^T S_w h0w = mu_w^T J0w mu_w = mu_w . h0w
        quad = (mu_w * h0w).sum(-1)
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
