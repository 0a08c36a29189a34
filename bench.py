#!/usr/bin/env python
"""Benchmark of the Gibbs hot path: sweeps/sec of SparseBernoulliGLM.resample_model().

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3|cfg2|cfg1] [--impl reference]

A "step" is one Gibbs sweep (psi -> PG -> weighted Gram -> spike-and-slab update of all neurons -> host network
step) over one synthetic recording.  Default workload = BASELINE.json's metric config: N=200, B=2, L=100, T=1e5
(configs[2], "cfg3"), which fits one B200.  With --gpus N (strong scaling: total work fixed) the default is the hybrid
partition: psi / PG / Gram over N time slabs, an exact int64 NCCL reduce-scatter of the integer Gram partials over the
neuron axis, the scan neuron-sharded, one all-gather of the new (a, W, b) rows; --shard neuron is the pure
neuron-sharded layout of BASELINE.json (X and the 32 GB of Z digit planes replicated: HBM-bound on Z).

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream around exactly K sweeps, barrier +
synchronize on both sides, max over ranks.  `value` uses device-resident data and a device-only timed region of
the sweep kernels; `e2e` times the public API call resample_model() with host state in/out every sweep.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "cfg1": dict(N=4, B=1, L=100, T=10000),
    "cfg2": dict(N=27, B=3, L=100, T=100000),
    "cfg3": dict(N=200, B=2, L=100, T=100000),
    # BASELINE.json configs[3] / [4]: 8-GPU workloads (time-sharded long recording; N=1000 neuron-sharded).  "cfg4r" is
    # ONE rank's slab of cfg4 (T/8), runnable on a single GPU; profiles/probe_rank_share.py plays one rank of cfg5.
    "cfg4": dict(N=100, B=3, L=100, T=10000000),
    "cfg4r": dict(N=100, B=3, L=100, T=1250000),
    "cfg5": dict(N=1000, B=1, L=100, T=1000000),
}


def synthetic_spikes(T, N, seed=0):
    """Y = (default_rng(seed).random((T,N)) < 0.05)."""
    return (np.random.default_rng(seed).random((T, N)) < 0.05).astype(np.float64)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), bf16_burst=d.get("bf16_tflops", 1590.0),
                    bf16_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


def int8_peak():
    """cuBLASLt int8 GEMM throughput measured on the pool's B200 (newest profiles/r*_int8_peak.json), or None."""
    import glob
    try:
        p = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_int8_peak.json")))[-1]
        with open(p) as f:
            d = json.load(f)
        return dict(tops=float(d["int8_gemm_8192_tops"]),
                    source="cuBLASLt int8 GEMM 8192^3 (torch._int_mm), best of 10, measured on B200: "
                           "profiles/%s (sustained: %.0f)" % (os.path.basename(p), d.get("int8_gemm_8192_sustained_tops", 0.0)))
    except Exception:
        return None


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed `ncu --set full` capture
    (profiles/<round>_ncu_key_metrics.json, same workload; the newest round present), in bytes; None when absent."""
    import glob
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    try:
        d = None
        for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_key_metrics.json")), reverse=True):
            with open(p) as f:
                allk = json.load(f)
            if kernel in allk:
                d = allk[kernel]
                break
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v, u = d[k].split()
            tot += float(v) * unit[u]
        return tot
    except Exception:
        return None


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index=0):
        self.index, self.samples, self.reasons, self.proc = index, [], set(), None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            try:
                self.samples.append((float(parts[0]), float(parts[1])))
                for nm, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        self.proc.terminate()
        sm = sorted(s[0] for s in self.samples)
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None),
                    sm_max_mhz=(self.samples[0][1] if self.samples else None), reasons=sorted(self.reasons))


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args, cfg, name):
    """The reference's CPU path (numpy/OpenBLAS + OpenMP Devroye PG) as restated by the oracle port, timed on the
    host cores with every thread they offer.  Each step is a BOUNDED SAMPLE of the sweep, sized so that the whole
    --steps K --warmup W run ends within a few minutes (REF_BUDGET_S): `sample_neurons` of the N postsynaptic
    regressions (independent and identically sized: scaled by N / sample_neurons), and -- only when a full-length
    regression does not fit the per-step budget -- the first T_s of the T time bins, with the T-proportional phases
    (psi, PG draws, dgemm Gram; exactly linear in T) scaled by T / T_s and the 2N-Cholesky a-scan + W draw of
    regression.py:282-340 measured in full."""
    from oracle import pyglm_oracle as O
    N, B, L, T = cfg["N"], cfg["B"], cfg["L"], cfg["T"]
    cores = os.cpu_count() or 1
    basis = O.cosine_basis(B, L) / L
    Y = synthetic_spikes(T, N)
    X = O.convolve_with_basis(Y, basis)
    steps = args.steps if args.steps is not None else 2
    warmup = args.warmup if args.warmup is not None else 1
    budget = float(os.environ.get("REF_BUDGET_S", "150")) / max(1, steps + warmup)

    def model_on(Ts):
        m = O.OracleSparseBernoulliGLM(N, basis, S_w=10.0, mu_b=-2.0, seed=0, pg_threads=cores)
        m.add_data(Y[:Ts], X=X[:Ts])
        return m

    # calibration on a short prefix: cost of the T-proportional part per bin, and of the scan
    T_cal = min(T, 10000)
    cal = model_on(T_cal)
    cal.resample_model(neurons=[0])
    aug_per_bin, scan_s = cal.t_aug / T_cal, cal.t_scan
    sample = min(N, args.ref_neurons)
    if sample * (aug_per_bin * T + scan_s) <= budget:
        T_s = T
    else:
        sample = 1
        T_s = int(min(T, max(T_cal, (budget - scan_s) / aug_per_bin)))
    m = model_on(T_s)
    neurons = list(range(sample))
    for _ in range(warmup):
        m.resample_model(neurons=neurons[:1])
    m.t_aug = m.t_scan = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        m.resample_model(neurons=neurons)
    wall = time.perf_counter() - t0
    per_reg = (m.t_aug * (T / float(T_s)) + m.t_scan + max(0.0, wall - m.t_aug - m.t_scan)) / (steps * sample)
    sweep_s = per_reg * N
    val = 1.0 / sweep_s
    if T_s == T:
        desc = "%d of %d regressions per step at full T=%d, scaled by N/%d" % (sample, N, T, sample)
    else:
        desc = ("%d of %d regressions per step on the first %d of T=%d bins: psi / PG / dgemm Gram scaled by T/T_s, "
                "a-scan and W draw measured in full; scaled by N/%d" % (sample, N, T_s, T, sample))
    line = dict(metric="gibbs_sweeps_per_sec", value=val, unit="sweeps/s", n_gpus=0, steps=steps,
                warmup=warmup, ms_per_step=sweep_s * 1e3,
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                impl="reference", config=dict(workload=name, **cfg),
                cpu_baseline=dict(value=val, unit="sweeps/s", cores=cores, kind="port", sample=desc),
                e2e=dict(value=val, unit="sweeps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


# ----------------------------------------------------------------------------------------------- our arm
_JSON_FD = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else any library prints to fd 1 (NCCL's
    version banner, for one) was redirected to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gram", default="auto", choices=["auto", "fp64", "tc"])
    ap.add_argument("--shard", default="auto", choices=["auto", "neuron", "time"],
                    help="auto: one GPU -> neuron; several -> time (psi / PG / Gram over time slabs, exact int64 "
                         "reduce-scatter of the Gram partials, then the neuron-sharded scan and the all-gather of W)")
    ap.add_argument("--ref-neurons", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-neurons", type=int, default=1)
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    name = "%s: SparseBernoulliGLM N=%d B=%d L=%d T=%d" % (args.config, cfg["N"], cfg["B"], cfg["L"], cfg["T"])

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args, cfg, name)
        return

    import torch
    import torch.distributed as dist
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) out of it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from pyglm_b200.models import SparseBernoulliGLM
    from pyglm_b200.utils.basis import cosine_basis

    if args.shard == "auto":
        args.shard = "time" if world > 1 else "neuron"
    steps = args.steps if args.steps is not None else 10
    warmup = max(3, args.warmup if args.warmup is not None else 3)
    N, B, L, T = cfg["N"], cfg["B"], cfg["L"], cfg["T"]
    np.random.seed(0)
    basis = cosine_basis(B=B, L=L) / L
    Y = synthetic_spikes(T, N)
    model = SparseBernoulliGLM(N, basis=basis, regression_kwargs=dict(S_w=10.0, mu_b=-2.), seed=1234,
                               gram=args.gram, shard=args.shard)
    model.add_data(Y, host_X=False)
    eng = model.engine
    K = eng.K

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- e2e: the public API, host state in / out every sweep -------------------------------------------
    for _ in range(warmup):
        model.resample_model()
    barrier()
    h2d0, d2h0, l0 = eng.h2d_bytes, eng.d2h_bytes, K.launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        model.resample_model()
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1) / steps
    h2d = (eng.h2d_bytes - h2d0) // steps
    d2h = (eng.d2h_bytes - d2h0) // steps
    launches = K.launches - l0

    # ---- `value`: whole sweeps with inputs resident in HBM (the engine call), with every kernel phase timed by
    # CUDA events on the launching stream inside the same timed region -------------------------------------
    ds = model._device_datasets()[0]
    A, W, b = model._host_state()
    n_loc = eng.psi_hi - eng.psi_lo
    hyp = model._stacked_hypers()
    barrier()
    eng.profile = {}
    ev0.record()
    for _ in range(steps):
        A, W, b = eng.sweep([ds], A, W, b, hyp)
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1) / steps
    kern_ms = eng.phase_ms()
    eng.profile = None
    clocks = sampler.stop() if rank == 0 else None
    names = sorted(kern_ms)

    t = torch.tensor([e2e_ms, dev_ms] + [kern_ms[nm] for nm in names], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms, dev_ms = float(t[0]), float(t[1])
    kern_ms = {nm: float(t[2 + i]) for i, nm in enumerate(names)}

    if rank == 0:
        pk = peaks()
        Dp = N * B + 1
        T_loc = ds.T
        gram_flop = n_loc * T_loc * Dp * (Dp + 1)             # symmetric minimum, SURVEY 8(d), per rank
        gram_tflops = gram_flop / (kern_ms["weighted_gram"] * 1e-3) / 1e12
        pg_bytes = 16.0 * T_loc * n_loc
        tc = "gram_tc_mma" in kern_ms
        if tc:
            # tcgen05 path: S(S+1)/2 int8 digit products per algorithmic MAC (S = 4 -> 10)
            plan = ds.buffers[("tc_plan", n_loc)]
            S = plan.S
            ops = gram_flop * (S * (S + 1) // 2)
            achieved = ops / (kern_ms["gram_tc_mma"] * 1e-3) / 1e12
            # denominator: cuBLASLt int8 GEMM (torch._int_mm, 8192^3, best of 10) measured on this pool's B200 by
            # profiles/profile_sweep.py --int8 (MEASURED_PEAKS.json has no int8 figure); else 2 x the bf16 burst figure
            i8 = int8_peak()
            peak = i8["tops"] if i8 else 2.0 * pk["bf16_burst"]
            roof = dict(kernel="gram_tc_kernel (tcgen05 kind::i8, %d radix-256 digits, exact int32/int64 sums)" % S,
                        bound="tensor", achieved=achieved, peak=peak, unit="TOP/s", frac=achieved / peak,
                        traffic=ncu_traffic("gram_tc_kernel"),
                        algorithmic="N*T*D*(D+1) FP64 flop x %d int8 digit products" % (S * (S + 1) // 2),
                        max_rel_dev_vs_fp64_kernel=plan.max_rel_dev,
                        peak_source=(i8["source"] if i8 else "2 x bf16_tflops of MEASURED_PEAKS.json (%s): int8 runs on "
                                     "the same tcgen05 pipe at twice the bf16 rate" % pk["source"]))
        else:
            fp64_peak = float(os.environ.get("PYGLM_FP64_PEAK_TFLOPS", "35.5"))
            roof = dict(kernel="gram_kernel (FP64 DMMA)", bound="tensor", achieved=gram_tflops, peak=fp64_peak,
                        unit="TFLOP/s", frac=gram_tflops / fp64_peak, traffic=None,
                        peak_source="FP64 tensor peak: cuBLAS DGEMM measured in profiles/r01_fp64_peak.json "
                                    "(MEASURED_PEAKS.json has no FP64 figure)")
        line = dict(
            metric="gibbs_sweeps_per_sec", value=1e3 / dev_ms, unit="sweeps/s", n_gpus=world, steps=steps,
            warmup=warmup, ms_per_step=dev_ms, higher_is_better=True, scaling="strong", vs_baseline=None,
            dtype="f64 (Gram: int8 digits on tcgen05, exact integer sums, <=1e-9 of FP64)" if tc else "f64",
            data="synthetic",
            config=dict(workload=name, parallelism=("time-sharded psi/PG/Gram + reduce-scatter of the Gram partials (%s), neuron-sharded scan + "
                                     "all-gather of (a, W, b), x%d" % ("exact int64" if tc else "FP64", world)
                                     if eng.shard == "time" and world > 1 else "%s-sharded x%d" % (eng.shard, world)), gram=("tc" if tc else "fp64"),
                        l2="inputs (X %.0f MB, omega %.0f MB%s per rank) exceed the 126 MB L2"
                           % (ds.Xp.numel() * 8 / 1e6, T_loc * eng.K.lib_ldn(n_loc) * 8 / 1e6,
                              ", Z digit planes %.1f GB" % (plan.Zs.numel() / 1e9) if tc else ""), **cfg),
            e2e=dict(value=1e3 / e2e_ms, unit="sweeps/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h)),
            gpu_launches=int(launches),
            roofline=roof,
            kernels_ms=kern_ms,
            dominant_kernel=max((k for k in kern_ms if not k.startswith("gram_")), key=lambda k: kern_ms[k]),
            pg_roofline=dict(bound="hbm", achieved=pg_bytes / (kern_ms["pg_draw"] * 1e-3) / 1e9, peak=pk["hbm_gbs"],
                             unit="GB/s", frac=pg_bytes / (kern_ms["pg_draw"] * 1e-3) / 1e9 / pk["hbm_gbs"],
                             peak_source=pk["source"]),
            weighted_gram_tflops=gram_tflops * world,
            clocks=clocks,
        )
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg, args.cpu_baseline_neurons)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(cfg, sample):
    """The oracle port of the reference sweep timed on the host cores for `sample` regressions (bounded)."""
    from oracle import pyglm_oracle as O
    N, B, L, T = cfg["N"], cfg["B"], cfg["L"], cfg["T"]
    cores = os.cpu_count() or 1
    basis = O.cosine_basis(B, L) / L
    Y = synthetic_spikes(T, N)
    m = O.OracleSparseBernoulliGLM(N, basis, S_w=10.0, mu_b=-2.0, seed=0, pg_threads=cores)
    m.add_data(Y, X=O.convolve_with_basis(Y, basis))
    t0 = time.perf_counter()
    m.resample_model(neurons=list(range(sample)))
    dt = time.perf_counter() - t0
    return dict(value=1.0 / (dt * N / sample), unit="sweeps/s", cores=cores, kind="port",
                sample="%d of %d regressions of one sweep at full T=%d (numpy/OpenBLAS + OpenMP Devroye PG), "
                       "scaled by N/%d" % (sample, N, T, sample))


if __name__ == "__main__":
    main()
