"""Workload for ncu: BASELINE cfg3 (N=200, B=2, L=100, T=1e5), one warm-up sweep + `--sweeps` sweeps through
the public API.  Not a benchmark: numbers printed under ncu are never bench values."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--sweeps", type=int, default=1)
ap.add_argument("--N", type=int, default=200)
ap.add_argument("--B", type=int, default=2)
ap.add_argument("--T", type=int, default=100000)
ap.add_argument("--dgemm", action="store_true", help="measure cuBLAS DGEMM / FP64 peak and exit")
ap.add_argument("--int8", action="store_true", help="measure cuBLASLt int8 GEMM (torch._int_mm) peak and exit")
args = ap.parse_args()

import torch

if args.dgemm:
    out = {}
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device="cuda")
        b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        for _ in range(2):
            a @ b
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out["dgemm_%d_tflops" % n] = 2 * n ** 3 / (best * 1e-3) / 1e12
    # sustained: back to back for ~3 s
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    out["dgemm_8192_sustained_tflops"] = reps * 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    import json
    print(json.dumps(out))
    sys.exit(0)

if args.int8:
    import json
    out = {}
    for n in (8192, 16384):
        a = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda")
        b = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda").t()   # column-major B (cuBLASLt "TN")
        try:
            for _ in range(3):
                torch._int_mm(a, b)
            best = 1e9
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                torch._int_mm(a, b)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out["int8_gemm_%d_tops" % n] = 2 * n ** 3 / (best * 1e-3) / 1e12
            reps = 60 if n == 8192 else 12
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize()
            out["int8_gemm_%d_sustained_tops" % n] = reps * 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
        except Exception as e:      # noqa: BLE001
            out["int8_gemm_%d_error" % n] = repr(e)[:200]
    print(json.dumps(out))
    sys.exit(0)

from pyglm_b200.models import SparseBernoulliGLM
from pyglm_b200.utils.basis import cosine_basis

np.random.seed(0)
basis = cosine_basis(B=args.B, L=100) / 100
Y = (np.random.default_rng(0).random((args.T, args.N)) < 0.05).astype(np.float64)
m = SparseBernoulliGLM(args.N, basis=basis, regression_kwargs=dict(S_w=10.0, mu_b=-2.), seed=1234)
m.add_data(Y, host_X=False)
for _ in range(1 + args.sweeps):
    m.resample_model()
print("ll", m.log_likelihood())
