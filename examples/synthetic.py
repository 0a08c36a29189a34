"""README configuration of the reference (examples/synthetic.py there): a 4-neuron network with self-inhibition is
simulated, a fresh model is fitted by Gibbs sampling, and the posterior means are compared with the truth.  Same
calls as the reference script, minus the plotting; the sample statistics are collected on the GPU instead of copying
the (T, N) firing rates to the host after every sweep.

    python examples/synthetic.py [--T 10000] [--N 4] [--sweeps 100]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyglm_b200.models import SparseBernoulliGLM  # noqa: E402
from pyglm_b200.utils.basis import cosine_basis  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=10000)
ap.add_argument("--N", type=int, default=4)
ap.add_argument("--B", type=int, default=1)
ap.add_argument("--L", type=int, default=100)
ap.add_argument("--sweeps", type=int, default=100)
args = ap.parse_args()
np.random.seed(0)
T, N, B, L = args.T, args.N, args.B, args.L

basis = cosine_basis(B=B, L=L) / L
true_model = SparseBernoulliGLM(N, basis=basis, regression_kwargs=dict(S_w=10.0, mu_b=-2.))
for n in range(N):
    true_model.regressions[n].a[n] = True
    true_model.regressions[n].W[n, :] = -2.0
_, Y = true_model.generate(T=T, keep=True)
print("simulated %d bins x %d neurons, mean rate %.3f, true log-likelihood %.1f"
      % (T, N, Y.mean(), true_model.log_likelihood()))

test_model = SparseBernoulliGLM(N, basis=basis, regression_kwargs=dict(S_w=10.0, mu_b=-2.))
test_model.add_data(Y)

lps = [test_model.log_likelihood()]
t0 = time.perf_counter()
for itr in range(args.sweeps):
    if itr == args.sweeps // 2:
        test_model.start_collecting(rates=True)        # second half of the chain = posterior samples
    test_model.resample_model()
    lps.append(test_model.log_likelihood())
dt = time.perf_counter() - t0
mom = test_model.posterior_moments()
print("%d sweeps in %.2f s (%.1f sweeps/s incl. a log-likelihood evaluation each)" % (args.sweeps, dt, args.sweeps / dt))
print("log-likelihood: start %.1f -> end %.1f (true model %.1f)" % (lps[0], lps[-1], true_model.log_likelihood()))
np.set_printoptions(precision=2, suppress=True)
print("posterior edge probabilities:\n", mom["A_mean"])
print("posterior mean self-weights:", np.array([mom["W_mean"][n, n].sum() for n in range(N)]),
      "(truth %.1f)" % (-2.0 * B))
print("posterior mean biases:", mom["b_mean"], "(truth", true_model.biases, ")")
err = np.abs(mom["rate_mean"][0] - true_model.means[0]).mean()
print("mean |posterior rate - true rate| = %.4f" % err)
