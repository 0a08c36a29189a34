#!/usr/bin/env python
"""Benchmark of the Gibbs hot path: sweeps/sec of SparseBernoulliGLM.resample_model().

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3|cfg2|cfg1|cfg4|cfg5] [--impl reference]

A "step" is one Gibbs sweep (psi -> PG -> weighted Gram -> spike-and-slab update of all neurons -> host network
step) over one synthetic recording.  Default workload = BASELINE.json's metric config: N=200, B=2, L=100, T=1e5
(configs[2], "cfg3"), which fits one B200.  With --gpus N (strong scaling: total work fixed) the default for cfg3 /
cfg4 is the hybrid partition: psi / PG / Gram over N time slabs, an exact int64 reduce-scatter of the integer Gram
partials over the neuron axis, the scan neuron-sharded, one all-gather of the new (a, W, b) rows; --shard neuron is the
pure neuron-sharded layout (X replicated, no exchange but the all-gather), the default for cfg5.
cfg4 = BASELINE configs[3] (N=100, B=3, T=1e7, time-sharded), cfg5 = configs[4] (N=1000, B=1, T=1e6, neuron-sharded,
latent-distance network prior): both meant for --gpus 8.

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream around exactly K sweeps, barrier +
synchronize on both sides, max over ranks.  `value` uses device-resident data and a device-only timed region of
the sweep kernels; `e2e` times the public API call resample_model() with host state in/out every sweep.
"""
import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv:
    # The reference arm is CPU work on rank 0 alone: give numpy / OpenBLAS / OpenMP every host core.  torchrun exports
    # OMP_NUM_THREADS=1 to its workers, which would handicap the baseline; this has to happen before numpy loads.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import argparse  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "cfg1": dict(N=4, B=1, L=100, T=10000),
    "cfg2": dict(N=27, B=3, L=100, T=100000),
    "cfg3": dict(N=200, B=2, L=100, T=100000),
    # BASELINE.json configs[3] / [4]: 8-GPU workloads (time-sharded long recording; N=1000 neuron-sharded with the
    # latent-distance network prior).  "cfg4r" is ONE rank's slab of cfg4 (T/8), runnable on a single GPU.
    "cfg4": dict(N=100, B=3, L=100, T=10000000),
    "cfg4r": dict(N=100, B=3, L=100, T=1250000),
    "cfg5": dict(N=1000, B=1, L=100, T=1000000),
}
SPIKE_CHUNK = 100000


def synthetic_spikes(T, N, seed=0, lo=0, hi=None):
    """Bins [lo, hi) of the synthetic recording.  T <= 1e5 (cfg1-3): SURVEY 8(d)'s recipe, Y = (default_rng(seed)
    .random((T, N)) < 0.05).  Longer recordings: the same recipe per chunk of 1e5 bins, chunk c seeded by (seed, c),
    so that a rank can generate just its slab."""
    hi = T if hi is None else hi
    if T <= SPIKE_CHUNK:
        return (np.random.default_rng(seed).random((T, N)) < 0.05).astype(np.float64)[lo:hi]
    out = np.empty((hi - lo, N), dtype=np.float64)
    for c in range(lo // SPIKE_CHUNK, (hi - 1) // SPIKE_CHUNK + 1):
        c0, c1 = c * SPIKE_CHUNK, min(T, (c + 1) * SPIKE_CHUNK)
        blk = np.random.default_rng([seed, c]).random((c1 - c0, N)) < 0.05
        a, b = max(lo, c0), min(hi, c1)
        out[a - lo:b - lo] = blk[a - c0:b - c0]
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), bf16_burst=d.get("bf16_tflops", 1590.0),
                    bf16_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def int8_peak():
    """cuBLASLt int8 GEMM throughput measured on the pool's B200 (newest profiles/r*_int8_peak.json), or None."""
    import glob
    try:
        p = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_int8_peak.json")))[-1]
        with open(p) as f:
            d = json.load(f)
        return dict(tops=float(d["int8_gemm_8192_tops"]), file="profiles/" + os.path.basename(p))
    except Exception:
        return None


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the newest committed `ncu --set full`
    capture of the SAME workload on one GPU (profiles/r*_ncu_key_metrics.json), in bytes; None when absent."""
    import glob
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    try:
        for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_key_metrics.json")), reverse=True):
            with open(p) as f:
                allk = json.load(f)
            if kernel in allk:
                tot = 0.0
                for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    v, u = allk[kernel][k].split()
                    tot += float(v) * unit[u]
                return tot
    except Exception:
        pass
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
                                          "--format=csv,noheader,nounits", "-lms", "25"],
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
        if not self.samples:
            time.sleep(0.1)                 # a timed region shorter than nvidia-smi's start-up: take what arrives now
        self.proc.terminate()
        sm = sorted(s[0] for s in self.samples)
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None),
                    sm_max_mhz=(self.samples[0][1] if self.samples else None), reasons=sorted(self.reasons))


# ----------------------------------------------------------------------------------------------- reference arm
def _pin_host_threads():
    cores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    return cores


def _spread(xs):
    xs = sorted(xs)
    return dict(min=xs[0], median=xs[len(xs) // 2], max=xs[-1], n=len(xs))


def run_reference(args, cfg, name):
    """The reference's CPU path (numpy/OpenBLAS dgemm Gram + 2N Choleskys per regression + OpenMP Devroye PG) as
    restated by the oracle port, timed on the host cores with every thread they offer (pinned here: torchrun exports
    OMP_NUM_THREADS=1).  The reference is a Python loop over N independent, identically sized regressions
    (models.py:169-171), so each step is a BOUNDED SAMPLE of the sweep: `s` whole regressions at full T, a DIFFERENT
    `s` every step (rotating through the neurons), sized so that the whole --steps K --warmup W run ends within a few
    minutes (REF_BUDGET_S); sweep time = N x the mean time of the regressions timed, whose spread is reported.  Only
    when one full-length regression does not fit the per-step budget (cfg4 / cfg5) are the T-proportional phases
    (psi, PG draws, dgemm Gram: exactly linear in T) timed on a time prefix and scaled."""
    from oracle import pyglm_oracle as O
    N, B, L, T = cfg["N"], cfg["B"], cfg["L"], cfg["T"]
    cores = _pin_host_threads()
    basis = O.cosine_basis(B, L) / L
    steps = args.steps if args.steps is not None else 4
    warmup = args.warmup if args.warmup is not None else 1
    budget = float(os.environ.get("REF_BUDGET_S", "200")) / max(1, steps + warmup)

    # calibration on a short prefix: cost of the T-proportional part per bin, and of the scan
    T_cal = min(T, 10000)
    Yc = synthetic_spikes(T, N, hi=T_cal)
    cal = O.OracleSparseBernoulliGLM(N, basis, S_w=10.0, mu_b=-2.0, seed=0, pg_threads=cores)
    cal.add_data(Yc, X=O.convolve_with_basis(Yc, basis))
    cal.resample_model(neurons=[0])
    aug_per_bin, scan_s = cal.t_aug / T_cal, cal.t_scan
    per_reg_est = aug_per_bin * T + scan_s
    if per_reg_est <= budget:
        T_s, sample = T, int(max(1, min(N, args.ref_neurons, budget // per_reg_est)))
    else:
        T_s, sample = int(min(T, max(T_cal, (budget - scan_s) / aug_per_bin))), 1
    Y = synthetic_spikes(T, N, hi=T_s)
    m = O.OracleSparseBernoulliGLM(N, basis, S_w=10.0, mu_b=-2.0, seed=0, pg_threads=cores)
    m.add_data(Y, X=O.convolve_with_basis(Y, basis))
    for _ in range(warmup):
        m.resample_model(neurons=[N - 1])
    per_reg, nxt = [], 0
    for _ in range(steps):
        for _ in range(sample):
            m.t_aug = m.t_scan = 0.0
            t0 = time.perf_counter()
            m.resample_model(neurons=[nxt % N])
            wall = time.perf_counter() - t0
            per_reg.append(m.t_aug * (T / float(T_s)) + m.t_scan + max(0.0, wall - m.t_aug - m.t_scan))
            nxt += 1
    sweep_s = float(np.mean(per_reg)) * N
    val = 1.0 / sweep_s
    if T_s == T:
        desc = ("%d of %d regressions per step at full T=%d, different ones every step (%d timed in all); "
                "sweep = N x their mean" % (sample, N, T, len(per_reg)))
    else:
        desc = ("1 of %d regressions per step on the first %d of T=%d bins (%d timed in all): psi / PG / dgemm Gram "
                "scaled by T/T_s, a-scan and W draw measured in full; sweep = N x their mean" % (N, T_s, T, len(per_reg)))
    # n_gpus mirrors the --gpus the arm was launched with (the driver pairs the two arms by it); the reference's path
    # runs on the host cores only
    line = dict(metric="gibbs_sweeps_per_sec", value=val, unit="sweeps/s", n_gpus=int(args.gpus), gpus_used=0, steps=steps,
                warmup=warmup, ms_per_step=sweep_s * 1e3,
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                impl="reference", config=dict(workload=name, **cfg),
                cpu_baseline=dict(value=val, unit="sweeps/s", cores=cores, kind="port", sample=desc,
                                  regressions_timed=len(per_reg), per_regression_s=_spread(per_reg),
                                  blas_threads=os.environ.get("OPENBLAS_NUM_THREADS"),
                                  extrapolated=(T_s != T or len(per_reg) < N)),
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
    ap.add_argument("--gram-stream", default="auto", choices=["auto", "on", "off"],
                    help="tensor-core Gram: build the Z digit tiles inside the kernel (on), stream resident digit planes "
                         "(off), or let the engine choose (auto)")
    ap.add_argument("--shard", default="auto", choices=["auto", "neuron", "time"],
                    help="auto: one GPU -> neuron; several -> time (psi / PG / Gram over time slabs, exact int64 "
                         "reduce-scatter of the Gram partials, then the neuron-sharded scan and the all-gather of W); "
                         "cfg5 -> neuron")
    ap.add_argument("--network", default="auto", choices=["auto", "niw", "latent-distance", "block"],
                    help="network prior; auto: the reference's default NIW prior, cfg5: latent distance model")
    ap.add_argument("--ref-neurons", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-neurons", type=int, default=2)
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    name = "%s: SparseBernoulliGLM N=%d B=%d L=%d T=%d" % (args.config, cfg["N"], cfg["B"], cfg["L"], cfg["T"])

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args, cfg, name)
        return

    t_start = time.perf_counter()
    import torch
    import torch.distributed as dist
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) out of it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from pyglm_b200 import networks
    from pyglm_b200.distributed import time_partition
    from pyglm_b200.models import SparseBernoulliGLM
    from pyglm_b200.utils.basis import cosine_basis
    t_init = time.perf_counter()

    if args.shard == "auto":
        args.shard = "neuron" if (world == 1 or args.config == "cfg5") else "time"
    if args.network == "auto":
        args.network = "latent-distance" if args.config == "cfg5" else "niw"
    big = cfg["N"] * cfg["T"] > 5e8
    steps = args.steps if args.steps is not None else (5 if big else 10)
    warmup = max(3, args.warmup if args.warmup is not None else 3)
    N, B, L, T = cfg["N"], cfg["B"], cfg["L"], cfg["T"]
    np.random.seed(0)
    basis = cosine_basis(B=B, L=L) / L
    net = {"niw": None, "latent-distance": networks.NIWLatentDistanceNetwork,
           "block": networks.NIWStochasticBlockNetwork}[args.network]
    model = SparseBernoulliGLM(N, basis=basis, network=None if net is None else net(N, B),
                               regression_kwargs=dict(S_w=10.0, mu_b=-2.), seed=1234, gram=args.gram, shard=args.shard,
                               gram_stream={"auto": "auto", "on": True, "off": False}[args.gram_stream])
    if args.shard == "time" and world > 1:
        lo, hi = time_partition(T, world, rank)
        h0 = max(0, lo - L)
        model.add_data(synthetic_spikes(T, N, lo=h0, hi=hi), host_X=False, slab=(h0, T))
    else:
        model.add_data(synthetic_spikes(T, N), host_X=False)
    eng = model.engine
    K = eng.K
    torch.cuda.synchronize()
    t_data = time.perf_counter()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- e2e: the public API, host state in / out every sweep -------------------------------------------
    model.resample_model()                               # first sweep: operand build + FP64 cross-check of the Gram
    torch.cuda.synchronize()
    t_first = time.perf_counter()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                                  # runs through the warm-up sweeps and both timed regions
    for _ in range(warmup - 1):
        model.resample_model()
    barrier()
    h2d0, d2h0, l0 = eng.h2d_bytes, eng.d2h_bytes, K.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        model.resample_model()
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1) / steps
    h2d = (eng.h2d_bytes - h2d0) // steps
    d2h = (eng.d2h_bytes - d2h0) // steps
    launches = K.launches - l0                           # our kernels launched inside the timed region (e2e loop)

    # ---- `value`: whole sweeps with inputs resident in HBM (the engine call), with every kernel phase timed by
    # CUDA events on the launching stream inside the same timed region -------------------------------------
    ds = model._device_datasets()[0]
    A, W, b = model._host_state()
    n_loc = eng.psi_hi - eng.psi_lo
    hyp = model._stacked_hypers()
    barrier()
    # Large single-GPU models sweep as two neuron groups on two streams (engine._sweep_overlapped): the phases of the
    # groups overlap, so per-phase events inside that region would not add up.  `value` is then timed on the overlapped
    # sweeps as they run in production, and the per-kernel times (and the roofline of the Gram kernel) come from a
    # second pass of the same number of sweeps with the overlap switched off -- same kernels, same chain, one block.
    overlapped = eng._overlap_groups([ds]) is not None
    if not overlapped:
        eng.profile = {}
    coll0 = eng.comm.collective_calls
    ev0.record()
    for _ in range(steps):
        A, W, b = eng.sweep([ds], A, W, b, hyp)
    ev1.record()
    coll = (eng.comm.collective_calls - coll0) / float(steps)
    barrier()
    dev_ms = ev0.elapsed_time(ev1) / steps
    serial_ms = None
    if overlapped:
        eng.overlap = False
        for _ in range(2):                               # builds / verifies the one-block plan, refills the pipeline
            A, W, b = eng.sweep([ds], A, W, b, hyp)
        barrier()
        eng.profile = {}
        ev0.record()
        for _ in range(steps):
            A, W, b = eng.sweep([ds], A, W, b, hyp)
        ev1.record()
        barrier()
        serial_ms = ev0.elapsed_time(ev1) / steps
        eng.overlap = True
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
        plan = ds.buffers.get(("tc_plan", n_loc))
        tc = "gram_tc_mma" in kern_ms and plan is not None
        if tc:
            # tcgen05 path: S(S+1)/2 int8 digit products per algorithmic MAC (S = 4 -> 10)
            S = plan.S
            ops = gram_flop * (S * (S + 1) // 2)
            achieved = ops / (kern_ms["gram_tc_mma"] * 1e-3) / 1e12
            # denominator: 2 x the driver-measured dense bf16 rate (int8 runs on the same tcgen05 pipe at twice the bf16
            # rate; MEASURED_PEAKS.json has no int8 figure).  The repo's own cuBLASLt int8 GEMM measurement is quoted
            # beside it.
            peak = 2.0 * pk["bf16_burst"]
            i8 = int8_peak()
            kname = ("gram_tcm_kernel" if plan.geom["nt"] <= 112 else "gram_tcs_kernel") if plan.stream else "gram_tc_kernel"
            roof = dict(kernel="%s (tcgen05 kind::i8, %d radix-256 digits, exact int32/int64 sums; %s)"
                               % (kname, S, "Z digit tiles built in the kernel" if plan.stream else "resident Z digit planes"),
                        bound="tensor", achieved=achieved, peak=peak, unit="TOP/s", frac=achieved / peak,
                        traffic=(ncu_traffic(kname) if world == 1 and args.config == "cfg3" else None),
                        algorithmic="N*T*D*(D+1) FP64 flop x %d int8 digit products" % (S * (S + 1) // 2),
                        max_rel_dev_vs_fp64_kernel=plan.max_rel_dev, max_rel_dev_spot_checks=plan.max_rel_dev_spot,
                        peak_source="2 x bf16_tflops (burst) of %s" % pk["source"],
                        frac_vs_2x_bf16_sustained=achieved / (2.0 * pk["bf16_sustained"]),
                        frac_vs_cublaslt_int8=(achieved / i8["tops"] if i8 else None),
                        cublaslt_int8_tops=(i8["tops"] if i8 else None),
                        cublaslt_int8_source=(i8["file"] if i8 else None))
        else:
            fp64_peak = float(os.environ.get("PYGLM_FP64_PEAK_TFLOPS", "35.5"))
            roof = dict(kernel="gram_kernel (FP64 DMMA)", bound="tensor", achieved=gram_tflops, peak=fp64_peak,
                        unit="TFLOP/s", frac=gram_tflops / fp64_peak, traffic=None,
                        peak_source="FP64 tensor peak: cuBLAS DGEMM measured in profiles/r01_fp64_peak.json "
                                    "(MEASURED_PEAKS.json has no FP64 figure)")
        if eng.shard == "time" and world > 1:
            par = ("time-sharded psi/PG/Gram + reduce-scatter of the Gram partials (%s), neuron-sharded scan + "
                   "all-gather of (a, W, b), x%d" % ("exact int64" if tc else "FP64", world))
            if eng.peer is not None:
                par += ("; exchanges by our own kernels over peer-mapped memory (reduce-scatter fused into the Gram "
                        "finalize pass, state rows pushed over NVLink), %.2f torch.distributed collectives per sweep" % coll)
            else:
                par += "; NCCL collectives (%.2f per sweep)" % coll
        else:
            par = "%s-sharded x%d" % (eng.shard, world)
            if overlapped:
                par = ("one GPU, two neuron groups on two streams: the scan of one group overlaps the psi / PG / Gram of "
                       "the other (dynamic item queue in the Gram kernel)")
            if world > 1:
                par += ("; state rows pushed over NVLink by our own kernel, %.2f torch.distributed collectives per sweep"
                        % coll) if eng.peer is not None else "; NCCL all-gather (%.2f collectives per sweep)" % coll
        resident = (", Z digit planes %.1f GB" % (plan.Zs.numel() / 1e9)) if (tc and not plan.stream) else ""
        line = dict(
            metric="gibbs_sweeps_per_sec", value=1e3 / dev_ms, unit="sweeps/s", n_gpus=world, steps=steps,
            warmup=warmup, ms_per_step=dev_ms, higher_is_better=True, scaling="strong", vs_baseline=None,
            dtype="f64 (Gram: int8 digits on tcgen05, exact integer sums, <=1e-9 of FP64)" if tc else "f64",
            data="synthetic",
            config=dict(workload=name, parallelism=par, gram=("tc" if tc else "fp64"), network=args.network,
                        l2="inputs (X %.0f MB, omega %.0f MB%s per rank) exceed the 126 MB L2"
                           % (ds.Xp.numel() * 8 / 1e6, T_loc * eng.K.lib_ldn(n_loc) * 8 / 1e6, resident), **cfg),
            e2e=dict(value=1e3 / e2e_ms, unit="sweeps/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h)),
            gpu_launches=int(launches), gpu_launches_per_step=int(launches // steps),
            roofline=roof,
            kernels_ms=kern_ms,
            dominant_kernel=max((k for k in kern_ms if not k.startswith("gram_")), key=lambda k: kern_ms[k]),
            pg_roofline=dict(bound="hbm", achieved=pg_bytes / (kern_ms["pg_draw"] * 1e-3) / 1e9, peak=pk["hbm_gbs"],
                             unit="GB/s", frac=pg_bytes / (kern_ms["pg_draw"] * 1e-3) / 1e9 / pk["hbm_gbs"],
                             peak_source=pk["source"]),
            weighted_gram_tflops=gram_tflops * world,
            setup_s=dict(import_and_nccl_init=t_init - t_start, spikes_and_add_data=t_data - t_init,
                         first_sweep_with_operand_build_and_fp64_check=t_first - t_data,
                         note="outside every timed region"),
            clocks=clocks,
        )
        if overlapped:
            line["overlap"] = dict(groups=2, one_block_ms_per_step=serial_ms,
                                   note="ms_per_step / value / e2e: overlapped sweeps; kernels_ms, roofline, pg_roofline: "
                                        "a second pass of the same sweeps with the overlap off (one block, one stream), "
                                        "where each kernel has the GPU to itself")
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg, args.cpu_baseline_neurons)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(cfg, sample):
    """The oracle port of the reference sweep timed on the host cores: whole regressions at full T when one fits ~15 s
    (cfg1-3), else one regression on a time prefix with the T-proportional phases scaled (bounded: ~10-30 s)."""
    from oracle import pyglm_oracle as O
    N, B, L, T = cfg["N"], cfg["B"], cfg["L"], cfg["T"]
    cores = _pin_host_threads()
    basis = O.cosine_basis(B, L) / L
    T_s = int(min(T, max(10000, 4e7 // (N * B))))        # cfg3: full T; cfg4 / cfg5: a prefix
    Y = synthetic_spikes(T, N, hi=T_s)
    m = O.OracleSparseBernoulliGLM(N, basis, S_w=10.0, mu_b=-2.0, seed=0, pg_threads=cores)
    m.add_data(Y, X=O.convolve_with_basis(Y, basis))
    per_reg = []
    for n in range(min(sample, N)):
        m.t_aug = m.t_scan = 0.0
        t0 = time.perf_counter()
        m.resample_model(neurons=[n])
        wall = time.perf_counter() - t0
        per_reg.append(m.t_aug * (T / float(T_s)) + m.t_scan + max(0.0, wall - m.t_aug - m.t_scan))
    val = 1.0 / (float(np.mean(per_reg)) * N)
    return dict(value=val, unit="sweeps/s", cores=cores, kind="port",
                sample="%d of %d regressions of one sweep on %s (numpy/OpenBLAS + OpenMP Devroye PG); sweep = N x their "
                       "mean" % (len(per_reg), N, "the full T=%d" % T if T_s == T else
                                 "the first %d of T=%d bins, psi / PG / Gram scaled by T/T_s" % (T_s, T)),
                per_regression_s=_spread(per_reg))


if __name__ == "__main__":
    main()
