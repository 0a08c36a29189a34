"""Numpy prototype of the algebra of csrc/spike_slab_dsm.cu (development aid, CPU only): the collapsed scan with
  * P = (Jp_SS)^-1 kept in SLOT space (a removed block leaves a zeroed tombstone that the next addition reuses),
  * a lookahead table for the next G inactive neurons: c_g, t_g = P c_g, and -- new -- the small matrices
    M = T^T C (all pairs of slots) and r_g = hp_g - c_g^T mu kept current under every flip by O(B^3) updates, so an
    add-evaluation needs no reduction over K at all,
checked step by step against oracle.collapsed_resample_a (log-odds and decisions).

    python profiles/proto_scan_dsm.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import pyglm_oracle as O  # noqa: E402


def scan(Jp, hp, cprior, lrho, a0, perm, us, B, G=4):
    N = len(a0)
    D = N * B + 1
    a = np.array(a0, dtype=bool).copy()
    # slot space: position 0 = bias, then blocks of B
    cidx = [D - 1]
    slot = -np.ones(N, dtype=int)
    for m in range(N):
        if a[m]:
            slot[m] = len(cidx)
            cidx += [m * B + b for b in range(B)]
    Ks = len(cidx)
    cap = D
    P = np.zeros((cap, cap))
    ci = np.array(cidx)
    P[:Ks, :Ks] = np.linalg.inv(Jp[np.ix_(ci, ci)])
    cidx = np.array(cidx + [-1] * (cap - Ks))
    mu = np.zeros(cap)
    mu[:Ks] = P[:Ks, :Ks] @ hp[ci]
    free = []
    cand = [-1] * G
    C = np.zeros((cap, G * B))
    T = np.zeros((cap, G * B))
    M = np.zeros((G * B, G * B))
    rv = np.zeros(G * B)
    logodds = np.zeros(N)
    nrefill = 0

    def blk(g):
        return slice(g * B, (g + 1) * B)

    for step in range(N):
        m = perm[step]
        pos = slot[m]
        if pos >= 0:
            S = P[pos:pos + B, pos:pos + B].copy()
            r = mu[pos:pos + B].copy()
            dpost = 0.5 * np.linalg.slogdet(S)[1] + 0.5 * r @ np.linalg.solve(S, r)
        else:
            if m not in cand:
                nrefill += 1
                g = 0
                cand = [-1] * G
                for i in range(step, N):
                    if slot[perm[i]] < 0 and g < G:
                        cand[g] = perm[i]
                        g += 1
                live = cidx[:Ks] >= 0
                C[:] = 0
                for g in range(G):
                    if cand[g] >= 0:
                        cols = cand[g] * B + np.arange(B)
                        C[:Ks, blk(g)] = np.where(live[:, None], Jp[np.ix_(np.maximum(cidx[:Ks], 0), cols)], 0.0)
                T[:Ks] = P[:Ks, :Ks] @ C[:Ks]
                M = T[:Ks].T @ C[:Ks]
                rv = np.array([hp[cand[g // B] * B + g % B] if cand[g // B] >= 0 else 0.0 for g in range(G * B)]) \
                    - C[:Ks].T @ mu[:Ks]
            g = cand.index(m)
            cols = m * B + np.arange(B)
            S = Jp[np.ix_(cols, cols)] - M[blk(g), blk(g)]
            r = rv[blk(g)].copy()
            dpost = -0.5 * np.linalg.slogdet(S)[1] + 0.5 * r @ np.linalg.solve(S, r)
        lo = dpost + cprior[m] + lrho[m]
        logodds[step] = lo
        v = us[step] > 1.0 / (1.0 + np.exp(lo))
        if pos < 0 and v:                                   # ---- commit add
            Gm = np.linalg.inv(S)
            gr = Gm @ r
            t = T[:Ks, blk(g)].copy()
            E = {}
            Dm = {}
            for g2 in range(G):
                if g2 != g and cand[g2] >= 0:
                    Dm[g2] = Jp[np.ix_(cols, cand[g2] * B + np.arange(B))]
                    E[g2] = M[blk(g), blk(g2)] - Dm[g2]
            if free:
                p = free.pop()
            else:
                p = Ks
                Ks += B
                t = np.vstack([t, np.zeros((B, B))])       # the new rows do not exist yet: zero
            tG = t @ Gm
            P[:Ks, :Ks] += tG @ t.T
            P[:Ks, p:p + B] = -tG
            P[p:p + B, :Ks] = -tG.T
            P[p:p + B, p:p + B] = Gm
            mu[:Ks] -= t @ gr
            mu[p:p + B] = gr
            for g2 in E:
                GE = Gm @ E[g2]
                T[:Ks, blk(g2)] += t @ GE
                T[p:p + B, blk(g2)] = -GE
                C[p:p + B, blk(g2)] = Dm[g2]
                rv[blk(g2)] += E[g2].T @ gr
            for g1 in E:
                for g2 in E:
                    M[blk(g1), blk(g2)] += E[g1].T @ Gm @ E[g2]
            cidx[p:p + B] = cols
            slot[m] = p
            a[m] = True
            cand[g] = -1
            T[:, blk(g)] = 0
            C[:, blk(g)] = 0
        elif pos >= 0 and not v:                            # ---- commit remove
            Gm = np.linalg.inv(S)
            tcol = P[:Ks, pos:pos + B].copy()
            gm = Gm @ r
            vv = {g2: T[pos:pos + B, blk(g2)].copy() for g2 in range(G) if cand[g2] >= 0}
            P[:Ks, :Ks] -= tcol @ Gm @ tcol.T
            P[pos:pos + B, :] = 0
            P[:, pos:pos + B] = 0
            mu[:Ks] -= tcol @ gm
            mu[pos:pos + B] = 0
            for g2 in vv:
                T[:Ks, blk(g2)] -= tcol @ (Gm @ vv[g2])
                T[pos:pos + B, blk(g2)] = 0
                C[pos:pos + B, blk(g2)] = 0
                rv[blk(g2)] += vv[g2].T @ gm
            for g1 in vv:
                for g2 in vv:
                    M[blk(g1), blk(g2)] -= vv[g1].T @ Gm @ vv[g2]
            cidx[pos:pos + B] = -1
            free.append(pos)
            slot[m] = -1
            a[m] = False
        elif pos < 0:
            cand[g] = -1
    return a, logodds, nrefill


def main():
    for (N, B, T, seed) in [(12, 2, 800, 0), (40, 3, 3000, 1), (60, 1, 3000, 2), (90, 2, 6000, 3)]:
        rng = np.random.default_rng(seed)
        Y = (rng.random((T, N)) < 0.1).astype(float)
        X = O.convolve_with_basis(Y, O.cosine_basis(B, 20) / 20).reshape(T, N * B)
        Sw = np.zeros((N, B, B))
        for m in range(N):
            Mx = rng.standard_normal((B, B))
            Sw[m] = Mx @ Mx.T + 0.5 * np.eye(B)
        hy = dict(rho=rng.uniform(0.1, 0.9, N), mu_w=rng.standard_normal((N, B)), S_w=Sw, mu_b=rng.standard_normal(1),
                  S_b=np.array([[rng.uniform(0.5, 2)]]))
        om = rng.random(T) * 0.25
        Jl, hl = O.lkhd_sufficient_statistics(X, om, Y[:, 0] - 0.5)
        J0, h0 = O.prior_sufficient_statistics(hy["mu_w"], hy["S_w"], hy["mu_b"], hy["S_b"])
        a0 = rng.random(N) < 0.5
        perm = rng.permutation(N)
        us = rng.random(N)
        trace = []
        a_ref = O.collapsed_resample_a(J0, h0, J0 + Jl, h0 + hl, a0, hy["rho"], B, perm, us, trace=trace)
        lo_ref = np.array([t[2] - t[1] for t in trace])
        Jw, hw, Jb, hb = O.natural_params(hy["mu_w"], hy["S_w"], hy["mu_b"], hy["S_b"])
        cprior = np.array([0.5 * np.linalg.slogdet(Jw[m])[1] - 0.5 * hw[m] @ np.linalg.solve(Jw[m], hw[m]) for m in range(N)])
        lrho = np.log(hy["rho"]) - np.log(1 - hy["rho"])
        a, lo, nref = scan(J0 + Jl, h0 + hl, cprior, lrho, a0, perm, us, B)
        print(N, B, "decisions equal:", np.array_equal(a, a_ref), " max |logodds diff|: %.2e" % np.max(np.abs(lo - lo_ref)),
              " flips:", int(np.sum(a != a0)), " refills:", nref)
        assert np.array_equal(a, a_ref) and np.max(np.abs(lo - lo_ref)) < 1e-8


if __name__ == "__main__":
    main()
