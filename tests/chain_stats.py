"""Comparison of a candidate Gibbs chain with the REFERENCE's own chains (tests/golden/chains_*.npz, written by
oracle/gen_chain_golden.py from /root/reference's unmodified sampler).  north_star: "posterior mean of A and W plus the
log-likelihood trace against long reference chains on the same synthetic data, via KS and tolerance tests".

No tolerance here is hand-set: every bound is the worst LEAVE-ONE-OUT discrepancy among the reference chains themselves
(chain c against the pool of the other chains) times a fixed slack of 1.5, i.e. the candidate must look like one more
reference chain.  examples/synthetic.py:51-83 is what the summaries mirror (LL trace, P(A), E[W], E[b])."""
import numpy as np
from scipy import stats

SLACK = 1.5
THIN = 5              # post-burn-in LL samples are thinned before the KS test (neighbouring sweeps are correlated)


def load_case(golden, name):
    g = golden(name)
    T, N, B, L = [int(v) for v in g["shape"]]
    Y = np.unpackbits(g["Ybits"])[:T * N].reshape(T, N).astype(np.float64)
    return g, Y, (T, N, B, L)


def summarize(lls, As, Ws, bs, burn):
    """Per-sweep records -> (LL trace, P(A), E[a o W], E[b]) over the post-burn-in sweeps."""
    A = np.array(As[burn:], dtype=np.float64)
    W = np.array(Ws[burn:])
    return (np.array(lls), A.mean(0), (A[..., None] * W).mean(0), np.array(bs[burn:]).mean(0))


def _ks(a, b):
    return stats.ks_2samp(a, b).statistic


def reference_spread(g):
    """Leave-one-out discrepancies of the reference chains: dict of the worst value per statistic."""
    burn = int(g["burn"])
    ll, PA, EW, Eb = g["ll"][:, burn:], g["PA"], g["EW"], g["Eb"]
    R = ll.shape[0]
    out = dict(ks=0.0, llmean=0.0, pa_mad=0.0, pa_max=0.0, ew_mad=0.0, eb_max=0.0, flips=0.0)
    for c in range(R):
        rest = [r for r in range(R) if r != c]
        pool = ll[rest][:, ::THIN].ravel()
        out["ks"] = max(out["ks"], _ks(ll[c][::THIN], pool))
        out["llmean"] = max(out["llmean"], abs(ll[c].mean() - ll[rest].mean()))
        pa_rest, ew_rest, eb_rest = PA[rest].mean(0), EW[rest].mean(0), Eb[rest].mean(0)
        out["pa_mad"] = max(out["pa_mad"], np.abs(PA[c] - pa_rest).mean())
        out["pa_max"] = max(out["pa_max"], np.abs(PA[c] - pa_rest).max())
        out["ew_mad"] = max(out["ew_mad"], np.abs(EW[c] - ew_rest).mean())
        out["eb_max"] = max(out["eb_max"], np.abs(Eb[c] - eb_rest).max())
        sure_on, sure_off = pa_rest > 0.9, pa_rest < 0.1
        out["flips"] = max(out["flips"], float(np.mean(PA[c][sure_on] < 0.5)) if sure_on.any() else 0.0,
                           float(np.mean(PA[c][sure_off] > 0.5)) if sure_off.any() else 0.0)
    return out


def assert_chain_matches_reference(g, cand_ll, cand_PA, cand_EW, cand_Eb):
    """cand_ll: full LL trace of the candidate chain (same number of sweeps as the reference chains)."""
    burn = int(g["burn"])
    ref = reference_spread(g)
    ll_ref = g["ll"][:, burn:]
    cl = np.asarray(cand_ll)[burn:]
    assert len(cl) == ll_ref.shape[1]
    n_par = g["PA"][0].size
    # Monte-Carlo floors: what two perfect samplers would still differ by with this many (correlated) samples
    n_eff = max(len(cl) // THIN, 1)
    ks_floor = 1.36 * np.sqrt(2.0 / n_eff)                 # 5 % critical value of the two-sample KS statistic
    ks = _ks(cl[::THIN], ll_ref[:, ::THIN].ravel())
    assert ks <= max(SLACK * ref["ks"], ks_floor), ("LL distribution (KS)", ks, ref["ks"], ks_floor)
    se = ll_ref.std() / np.sqrt(n_eff)
    assert abs(cl.mean() - ll_ref.mean()) <= SLACK * ref["llmean"] + 2 * se, ("LL plateau", cl.mean(), ll_ref.mean(), ref)
    PA, EW, Eb = g["PA"].mean(0), g["EW"].mean(0), g["Eb"].mean(0)
    floor = 1.0 / np.sqrt(n_eff)
    assert np.abs(cand_PA - PA).mean() <= SLACK * ref["pa_mad"] + floor / np.sqrt(n_par), ("P(A), mean abs dev", ref)
    assert np.abs(cand_PA - PA).max() <= SLACK * ref["pa_max"] + floor, ("P(A), worst entry", np.abs(cand_PA - PA).max(), ref)
    assert np.abs(cand_EW - EW).mean() <= SLACK * ref["ew_mad"] + 0.05 * np.abs(EW).mean(), ("E[a W], mean abs dev", ref)
    assert np.abs(cand_Eb - Eb).max() <= SLACK * ref["eb_max"] + 2 * floor * g["Eb"].std(0).max() + 1e-3, ("E[b]", ref)
    sure_on, sure_off = PA > 0.9, PA < 0.1
    if sure_on.any():
        assert np.mean(cand_PA[sure_on] < 0.5) <= SLACK * ref["flips"] + 1.0 / max(sure_on.sum(), 1), "confident edges lost"
    if sure_off.any():
        assert np.mean(cand_PA[sure_off] > 0.5) <= SLACK * ref["flips"] + 1.0 / max(sure_off.sum(), 1), "spurious edges"
    return dict(ks=ks, ref=ref)
