"""GPU probe of the tcgen05 Gram (csrc/gram_tc.cu): digit planes and integer sums against a numpy emulation
(bit-exact), J against the FP64 DMMA kernel, and timing at the bench configuration.

    python profiles/probe_gram_tc.py [--big] [--S 4]
"""
import argparse
import math
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pyglm_b200.kernels import CudaKernels, pad_ldn  # noqa: E402
from pyglm_b200.utils.basis import cosine_basis  # noqa: E402
from oracle import pyglm_oracle as ORC  # noqa: E402  (checker only)


def exponent(c):
    return 0 if not c > 0 else math.frexp(c * 1.02)[1]


def digits(scaled, S):
    v = np.rint(scaled).astype(np.int64)
    out = []
    for _ in range(S - 1):
        lo = ((v + 128) & 255) - 128
        out.append(lo)
        v = (v - lo) >> 8
    out.append(v)
    return out[::-1]


def make(K, T, N, B, n, seed=0):
    rng = np.random.default_rng(seed)
    Y = (rng.random((T, N)) < 0.05).astype(np.float64)
    basis = cosine_basis(B=B, L=min(100, max(4, T // 4))) / min(100, max(4, T // 4))
    Xp = K.filter_spikes(K.to_device(Y), K.to_device(basis), True)
    om = K.zeros(T, pad_ldn(n))
    om[:, :n] = torch.from_numpy(0.02 + 0.4 * rng.random((T, n)) ** 3).to(K.device)
    return Xp, om


def check_exact(K, T, N, B, n, S):
    D = N * B + 1
    Xp, om = make(K, T, N, B, n, seed=T + N)
    plan = K.gram_tc_plan(Xp, D, n, S)
    g = plan.geom
    print("case T=%d D=%d n=%d S=%d geom=%s" % (T, D, n, S, g))
    X = Xp.cpu().numpy()[:, :D]
    O = om.cpu().numpy()[:, :n]
    cmax = X.max(0)
    assert np.array_equal(plan.cmax.cpu().numpy(), cmax), "column max mismatch"
    ex = [exponent(c) for c in cmax]
    Zs = plan.Zs.cpu().numpy()
    zd = {}
    bad = 0
    for i in range(D):
        for j in range(i + 1):
            p = i * (i + 1) // 2 + j
            d = ORC.tc_z_digits(X, i, j, ex, S)
            zd[p] = d
            for s in range(S):
                if not np.array_equal(Zs[s, p, :T], (d[s] & 255).astype(np.uint8)):
                    bad += 1
            assert d[0].min() >= 0 and d[0].max() <= 255
    assert Zs[:, :, T:].max(initial=0) == 0 and Zs[:, g["M"]:, :].max(initial=0) == 0
    print("  Z digit planes mismatching rows:", bad)
    plan.slice_omega(om)
    omax = O.max(0)
    assert np.array_equal(plan.omax.cpu().numpy(), omax)
    eo = [exponent(c) for c in omax]
    Os = plan.Os.cpu().numpy()
    od = []
    bad_o = 0
    for c in range(n):
        d = digits(O[:, c] * 2.0 ** (8 * S - eo[c]), S)
        od.append(d)
        for s in range(S):
            if not np.array_equal(Os[s, c, :T], (d[s] & 255).astype(np.uint8)):
                bad_o += 1
    print("  omega digit planes mismatching rows:", bad_o)
    Jint = plan.mma().cpu().numpy()
    torch.cuda.synchronize()
    exp = np.zeros((n, g["M"]), dtype=np.int64)
    for p in range(g["M"]):
        for c in range(n):
            tot = 0
            for a in range(S):
                for b in range(S - a):
                    tot += int(np.dot(zd[p][a], od[c][b])) << (8 * (S - 1 - a - b))
            exp[c, p] = tot
    nbad = int((Jint[:, :g["M"]] != exp).sum())
    print("  Jint mismatches: %d of %d   (max |Jint| = %.3e)" % (nbad, exp.size, np.abs(exp).max()))
    if nbad:
        w = np.argwhere(Jint[:, :g["M"]] != exp)[:10]
        for c, p in w:
            print("    n=%d p=%d got %d want %d" % (c, p, Jint[c, p], exp[c, p]))
    J = plan.gram(om).cpu().numpy()
    Jref = K.weighted_gram(Xp, om, D, n).cpu().numpy()
    il, jl = np.tril_indices(D)
    rel = np.abs(J[:, il, jl] - Jref[:, il, jl]) / np.abs(Jref[:, il, jl]).clip(1e-300)
    print("  J vs FP64 DMMA: max rel %.3e" % rel.max())
    return bad == 0 and bad_o == 0 and nbad == 0


def check_vs_fp64(K, T, N, B, n, S, reps=3):
    D = N * B + 1
    Xp, om = make(K, T, N, B, n, seed=1)
    torch.cuda.synchronize()
    t0 = time.time()
    plan = K.gram_tc_plan(Xp, D, n, S)
    torch.cuda.synchronize()
    print("case T=%d D=%d n=%d S=%d: build_z %.1f ms, geom=%s, Zs %.2f GB" %
          (T, D, n, S, 1e3 * (time.time() - t0), plan.geom, plan.Zs.numel() / 1e9))
    J = K.zeros(n, plan.ldx, plan.ldx)
    plan.gram(om, J)
    Jref = K.weighted_gram(Xp, om, D, n)
    torch.cuda.synchronize()
    il, jl = np.tril_indices(D)
    a = J.cpu().numpy()[:, il, jl]
    b = Jref.cpu().numpy()[:, il, jl]
    rel = np.abs(a - b) / np.abs(b).clip(1e-300)
    print("  J vs FP64 DMMA: max rel %.3e  median %.3e" % (rel.max(), np.median(rel)))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for r in range(reps):
        ev[0].record(); plan.slice_omega(om)
        ev[1].record(); plan.mma()
        ev[2].record(); plan.finalize(J)
        ev[3].record(); torch.cuda.synchronize()
        t = [ev[k].elapsed_time(ev[k + 1]) for k in range(3)]
        flops = n * T * D * (D + 1)
        print("  rep %d: slice %.3f ms  mma %.3f ms  finalize %.3f ms   (%.1f algorithmic TFLOP/s)" %
              (r, t[0], t[1], t[2], flops / (sum(t) * 1e-3) / 1e12))
    for label, arg in (("multicast on ", 0), ("multicast off", -1)):
        ts = []
        for r in range(8):
            ev[0].record(); plan.mma(arg); ev[1].record(); torch.cuda.synchronize()
            ts.append(ev[0].elapsed_time(ev[1]))
        print("  mma %s: min %.3f  median %.3f ms" % (label, min(ts), sorted(ts)[len(ts) // 2]))
    ev[0].record(); plan.mma_probe(); ev[1].record(); torch.cuda.synchronize()
    tp = ev[0].elapsed_time(ev[1])
    S_ = plan.S
    g = plan.geom
    ops = 2.0 * (S_ * (S_ + 1) // 2) * g["Mpad"] * g["Npad"] * g["Tpad"]
    print("  issue-rate probe (no loads/atomics): %.3f ms -> %.0f int8 TOP/s executed (padded shape)" % (tp, ops / tp / 1e9))
    plan.mma()
    ev[0].record(); K.weighted_gram(Xp, om, D, n, J=J); ev[1].record(); torch.cuda.synchronize()
    print("  FP64 DMMA kernel: %.3f ms" % ev[0].elapsed_time(ev[1]))
    return rel.max()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--S", type=int, default=4)
    a = ap.parse_args()
    K = CudaKernels()
    ok = check_exact(K, 1000, 5, 2, 7, a.S)
    ok &= check_exact(K, 333, 3, 1, 3, 3)
    ok &= check_exact(K, 700, 8, 2, 20, 5)
    print("EXACT CHECKS:", "PASS" if ok else "FAIL")
    check_vs_fp64(K, 40000, 16, 2, 40, a.S)
    check_vs_fp64(K, 20000, 10, 2, 200, a.S)
    if a.big:
        check_vs_fp64(K, 100000, 200, 2, 200, a.S)
