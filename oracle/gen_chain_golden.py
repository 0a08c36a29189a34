"""ORACLE (test infrastructure): long Gibbs chains of the REFERENCE'S OWN sampler (/root/reference, unmodified files,
run through oracle/ref_shim) on synthetic data, summarised into tests/golden/chains_*.npz.  Run in the build container:

    python oracle/gen_chain_golden.py cfg1      # N=4,  B=1, L=100, T=1e4  (BASELINE configs[0], examples/synthetic.py)
    python oracle/gen_chain_golden.py cfg2s     # N=27, B=3, L=100, T=2e4  (configs[1] with a shorter recording)

north_star: "posterior mean of A and W plus the log-likelihood trace against long reference chains on the same
synthetic data, via KS and tolerance tests".  The data come from the reference's generate() with the true model of
examples/synthetic.py:27-33 (S_w=10, mu_b=-2, self-weights -2); the fitted model is :40-44.  Several chains with
different numpy seeds are run so that the tests can derive their tolerances from the between-chain spread.  The draws
use the shim's stand-ins for pypolyagamma / pybasicbayes (absent third-party code): the chains are the reference's
sweep logic, line for line, on restated stochastic primitives -- PARITY UNPINNED for the primitives themselves.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "ref_shim"))
import sitecustomize_shim  # noqa: E402,F401

import warnings  # noqa: E402
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
from pyglm.models import SparseBernoulliGLM  # noqa: E402
from pyglm.utils.basis import cosine_basis  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")

CASES = {
    "cfg1": dict(N=4, B=1, L=100, T=10000, sweeps=600, burn=100, seeds=(1, 2, 3, 4, 5, 6), data_seed=0),
    "cfg2s": dict(N=27, B=3, L=100, T=20000, sweeps=450, burn=150, seeds=(1, 2, 3, 4, 5, 6, 7, 8), data_seed=0),
}


def run(name):
    c = CASES[name]
    N, B, L, T = c["N"], c["B"], c["L"], c["T"]
    basis = cosine_basis(B=B, L=L) / L
    np.random.seed(c["data_seed"])
    true = SparseBernoulliGLM(N, basis=basis, regression_kwargs=dict(S_w=10.0, mu_b=-2.))
    for n in range(N):
        true.regressions[n].a[n] = True
        true.regressions[n].W[n, :] = -2.0
    _, Y = true.generate(T=T, keep=True)
    ll_true = true.log_likelihood()
    out = dict(Ybits=np.packbits(Y.astype(np.uint8)), shape=np.array([T, N, B, L]), basis=basis,
               true_A=true.adjacency, true_W=true.weights, true_b=true.biases, ll_true=ll_true,
               sweeps=c["sweeps"], burn=c["burn"], seeds=np.array(c["seeds"]))
    lls, PA, EW, Eb, EW2 = [], [], [], [], []
    for seed in c["seeds"]:
        np.random.seed(1000 + seed)
        m = SparseBernoulliGLM(N, basis=basis, regression_kwargs=dict(S_w=10.0, mu_b=-2.))
        m.add_data(Y)
        ll = np.zeros(c["sweeps"])
        sA = np.zeros((N, N)); sW = np.zeros((N, N, B)); sW2 = np.zeros((N, N, B)); sb = np.zeros(N)
        t0 = time.time()
        for it in range(c["sweeps"]):
            m.resample_model()
            ll[it] = m.log_likelihood()
            if it >= c["burn"]:
                A, W = m.adjacency, m.weights
                sA += A; sW += A[:, :, None] * W; sW2 += (A[:, :, None] * W) ** 2; sb += m.biases
            if it % 20 == 0:
                print("%s seed %d sweep %d ll %.1f (true %.1f) %.0fs" % (name, seed, it, ll[it], ll_true, time.time() - t0),
                      flush=True)
        k = c["sweeps"] - c["burn"]
        lls.append(ll); PA.append(sA / k); EW.append(sW / k); EW2.append(sW2 / k); Eb.append(sb / k)
    out.update(ll=np.array(lls), PA=np.array(PA), EW=np.array(EW), EW2=np.array(EW2), Eb=np.array(Eb))
    np.savez_compressed(os.path.join(OUT, "chains_%s.npz" % name), **out)
    print("wrote chains_%s.npz" % name)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES)):
        run(nm)
