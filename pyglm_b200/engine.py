"""Device-resident state and the Gibbs sweep driver.

One GibbsEngine per process / GPU.  It owns the padded design matrices, the spike matrix, the per-sweep
work buffers (psi, omega, J, the P workspace) and launches the five hot-path kernels in the order of the
reference's sweep (models.py:166-171 -> regression.py:265-280):

    Wt <- (a o W, b)            host build + one H2D copy of N*D doubles
    psi = Xp Wt                 activation.cu
    omega ~ PG(1, psi)          polyagamma.cu
    J_n = Xp^T diag(omega_n) Xp gram.cu        (h_n = Xp^T (y_n - 1/2) is sweep-invariant: cached per dataset)
    [time-sharded: reduce-scatter J over the neuron axis]
    (a, W, b) update            spike_slab.cu
    [multi-GPU: all-gather (a, W, b)]

X, Y, psi, omega, J never leave the device; only O(N^2 B) state crosses PCIe per sweep.
"""
import os

import numpy as np
import torch

from .distributed import Comm, PeerExchange, StateExchange, block_partition
from .kernels import CudaKernels, gram_tc_bytes, pad_ldn, pad_ldx
from .priors import prior_arrays

_KERNELS = {}


def default_kernels(device=None):
    """Process-wide CudaKernels for a device (raises without a B200 / the built library: no fallback)."""
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("pyglm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device not in _KERNELS:
        _KERNELS[device] = CudaKernels(device)
    return _KERNELS[device]


class DeviceDataset(object):
    """One (X, Y) pair resident in HBM.  Xp: (T, ldx) padded design; Y: (T, N) float64; t_off: global index of
    the first local time bin (time-sharded runs); h_cache: sweep-invariant X~^T (Y - 1/2) per neuron range."""

    def __init__(self, Xp, Y, t_off=0, T_global=None):
        self.Xp, self.Y = Xp, Y
        self.T = Xp.shape[0]
        self.t_off = t_off
        self.T_global = self.T if T_global is None else T_global
        self.h_cache = {}
        self.buffers = {}


class DeviceMoments(object):
    """Running first and second moments of the Gibbs samples, kept in HBM (the reference's example
    loops pull the (T, N) rates and the full state to the host every sweep, examples/synthetic.py:51-59).  After each
    sweep the new state rows [a | W | b] -- already on the device for the all-gather -- and, optionally, the firing
    rates logistic(psi) of this rank's (time, neuron) block are added to sum / sum-of-squares buffers; nothing
    crosses PCIe until moments() is called."""

    def __init__(self, rates=False):
        self.rates = bool(rates)
        self.n = 0
        self.s1 = self.s2 = None            # (N, N + N*B + 1) sums of the state rows
        self.r1, self.r2 = {}, {}           # dataset index -> (T_loc, n_psi) sums of the rates

    def add_state(self, state, width):
        x = state[:, :width]
        if self.s1 is None:
            self.s1, self.s2 = torch.zeros_like(x), torch.zeros_like(x)
        self.s1 += x
        self.s2 += x * x
        self.n += 1

    def add_rates(self, di, psi, n):
        mu = torch.sigmoid(psi[:, :n])
        if di not in self.r1:
            self.r1[di], self.r2[di] = torch.zeros_like(mu), torch.zeros_like(mu)
        self.r1[di] += mu
        self.r2[di] += mu * mu


class GibbsEngine(object):
    def __init__(self, N, B, kernels=None, seed=0, comm=None, shard="neuron", gram="auto", gram_digits=4,
                 gram_stream="auto"):
        """gram: "fp64" = FP64 DMMA kernel (gram.cu); "tc" = tcgen05 integer-digit kernel (gram_tc.cu), checked
        against the FP64 kernel on the first sweep (<= 5e-10 relative, else one more digit, else FP64);
        "auto" = "tc" when the design is non-negative, the contraction is large enough to matter and the digit
        planes of Z fit in HBM, else "fp64".  gram_digits: radix-256 digits the tc path starts with (4 or 5)."""
        assert shard in ("neuron", "time")
        assert gram in ("auto", "fp64", "tc") and gram_digits in (4, 5) and gram_stream in ("auto", True, False)
        self.gram_mode, self.gram_digits = gram, gram_digits
        env = os.environ.get("PYGLM_TC_STREAM")
        self.gram_stream = gram_stream if env is None else (env != "0")
        self.N, self.B = N, B
        self.D = N * B + 1
        self.ldx = pad_ldx(self.D)
        self.K = default_kernels() if kernels is None else kernels
        self.comm = Comm() if comm is None else comm
        self.shard = shard
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.calls = 0                       # Philox call_id counter: identical on every rank
        lo, hi, n_max = block_partition(N, self.comm.world, self.comm.rank)
        self.scan_lo, self.scan_hi, self.n_max = lo, hi, n_max
        if shard == "neuron" or self.comm.world == 1:
            self.psi_lo, self.psi_hi = lo, hi
        else:
            self.psi_lo, self.psi_hi = 0, N
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._ws = {}
        # Test hook: dict(omega=[(T,N) per dataset], perm=(N,N) int, us=(N,N), z=(N,D)) replaces the device
        # draws of the next sweeps with the given host arrays (row n = postsynaptic neuron n), so that a whole
        # sweep can be compared with the reference on identical randomness.  None in production.
        self.inject = None
        # Optional per-phase device timing: set to {} and every sweep appends (start, end) CUDA-event pairs per
        # phase name; phase_ms() averages them.  Events are recorded on the launching stream.
        self.profile = None
        # Software pipelining of consecutive sweeps (see sweep()); the pre-launched Gram of the next sweep.
        self.pipeline = True
        self._pending = None
        # Optional on-device sample statistics (DeviceMoments); None = off.
        self.moments = None
        # Opt-in (PYGLM_OVERLAP=1 or .overlap = True): single-GPU sweeps of >= OVERLAP_MIN_N neurons run as two neuron
        # groups on two streams (see _sweep_overlapped), the scan of one group sharing the GPU with the psi / PG / Gram of
        # the other.  Off by default: measured at cfg3 it only pays while the chain is sparse (16.0 against 18.5 ms per
        # sweep over the first sweeps); at the equilibrium density the scan is L2-bandwidth bound, a group's single wave
        # (3.8 ms, stretched by the co-running psi / PG streams) costs as much as its share of the one-block scan
        # (6.5 ms for 1.35 waves, whose tail wave runs with less contention), and both modes take 19.3-19.8 ms
        # (profiles/r02ac_sweep_times.log).
        self.overlap = os.environ.get("PYGLM_OVERLAP", "0") == "1"
        self.overlap_min_n = self.OVERLAP_MIN_N
        self._ovl = None
        # Exchange steps over peer-mapped memory with our own kernels (distributed.PeerExchange) when the ranks are GPUs
        # of one NVLink domain; otherwise (gloo CPU tests, PYGLM_PEER_EXCHANGE=0, allocation failure) the NCCL / gloo
        # collectives of Comm.  All ranks decide alike (the constructor is collective).
        self.peer = None
        self._state_xchg = None
        if PeerExchange.usable(self.comm, self.K.device):
            ok = 1
            try:
                self.peer = PeerExchange(self.comm, self.K)
                self._state_xchg = StateExchange(self.peer, self.n_max, self._state_width())
            except Exception as e:                       # pragma: no cover - depends on the box
                import warnings
                warnings.warn("peer-memory exchange unavailable (%s): using NCCL collectives" % (e,))
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=self.K.device)
            if not bool(self.comm.all_reduce_min(flag).item()):
                self.peer = self._state_xchg = None
            else:
                self.comm.peer = self.peer

    def _state_width(self):
        """Doubles per state row [a (N) | W (N*B) | b | status], padded to an even count (16-byte rows for the push)."""
        w = self.N + self.N * self.B + 2
        return w + (w & 1)

    def _gather_state(self, state):
        """(n_max, width) rows of this rank's scan block -> (world * n_max, width) rows of all ranks."""
        if self._state_xchg is not None:
            return self._state_xchg.all_gather_rows(state)
        return self.comm.all_gather_rows(state)

    def _mark(self, name, start=None):
        if self.profile is None:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        if start is not None:
            self.profile.setdefault(name, []).append((start, ev))
        return ev

    def phase_ms(self):
        torch.cuda.synchronize()
        return {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in (self.profile or {}).items()}

    # ------------------------------------------------------------------ data
    def make_dataset(self, Y, basis=None, X=None, t_off=0, T_global=None):
        """Y (T,N) host float64; either basis (L,B) -> filter on device, or X (T,N,B)/(T,NB) host -> pack."""
        K = self.K
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        Yd = K.to_device(Y)
        self.h2d_bytes += Y.nbytes
        if X is None:
            basis = np.ascontiguousarray(basis, dtype=np.float64)
            clip = bool(np.amin(basis) >= 0 and np.amin(Y) >= 0)
            Xp = K.filter_spikes(Yd, K.to_device(basis), clip)
        else:
            Xh = np.ascontiguousarray(np.reshape(X, (Y.shape[0], self.N * self.B)), dtype=np.float64)
            self.h2d_bytes += Xh.nbytes
            Xp = K.pack_design(K.to_device(Xh))
        return DeviceDataset(Xp, Yd, t_off, T_global)

    def _buf(self, ds, name, shape, dtype=torch.float64, zero=False):
        key = (name, tuple(shape), dtype)
        if key not in ds.buffers:
            ds.buffers[key] = (self.K.zeros if zero else self.K.empty)(*shape, dtype=dtype)
        return ds.buffers[key]

    def _wsbuf(self, name, shape, dtype=torch.float64, zero=False):
        key = (name, tuple(shape), dtype)
        if key not in self._ws:
            self._ws[key] = (self.K.zeros if zero else self.K.empty)(*shape, dtype=dtype)
        return self._ws[key]

    # ------------------------------------------------------------------ Gram dispatch
    TC_STREAM_DEFAULT = True    # "auto": build Z tiles in the kernel even when the resident planes would fit (measured:
                                # cfg3 10.7 ms against 11.4 ms, and no 32 GB of planes / 37 ms build per data set)
    TC_MIN_WORK = 5e9           # pairs * T * neurons below which the FP64 kernel is used (cfg2 = 9e9: tc 0.49 ms vs 1.28 ms)
    TC_ACCEPT = 5e-10           # accepted max relative deviation from the FP64 kernel (stated tolerance 1e-9, 2x margin)

    def _time_sharded(self):
        return self.shard == "time" and self.comm.world > 1

    def _agree(self, ok):
        """Time-sharded ranks take the Gram path decisions together (the reduce-scatter that follows must see the
        same dtype and shape on every rank): a choice holds only if it holds everywhere."""
        if not self._time_sharded():
            return bool(ok)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.K.device)
        return bool(self.comm.all_reduce_min(flag).item())

    def _tc_stream(self, ds, n, digits):
        """True: the kernel builds the Z digit tiles in shared memory (4 digits only); False: resident digit planes;
        None: neither is possible here.  gram_stream True / False force one; "auto" streams when the planes would not
        fit in HBM, and otherwise follows TC_STREAM_DEFAULT (the measured choice).  All ranks decide alike."""
        free = torch.cuda.mem_get_info(self.K.device)[0]
        fits = self._agree(gram_tc_bytes(self.D, n, ds.T, digits) < 0.9 * free)
        if self.gram_stream is True:
            return True if digits == 4 else None
        if self.gram_stream is False:
            return False if fits else None
        if digits == 4 and (self.TC_STREAM_DEFAULT or not fits):
            return True
        return False if fits else None

    @staticmethod
    def _tc_key(n, tag=None):
        return ("tc_plan", n) if tag is None else ("tc_plan", n, tag)

    def _tc_build(self, ds, n, digits, share=None):
        stream = self._tc_stream(ds, n, digits)
        if stream is None:
            return None
        need = gram_tc_bytes(self.D, n, ds.T, digits, stream=stream)
        free = torch.cuda.mem_get_info(self.K.device)[0]
        if not self._agree(need < 0.9 * free):
            return None
        try:
            if self._time_sharded():
                plan = self.K.gram_tc_plan(ds.Xp, self.D, n, digits, comm=self.comm, t_off=ds.t_off, stream=stream)
                # Jint with the row padding the reduce-scatter over the neuron axis needs (extra rows stay zero)
                rows = self.comm.world * self.n_max
                if self.peer is not None:
                    # peer-mapped: the owner of a neuron block reads these partial sums from every rank's HBM
                    plan.Jint, plan.peer_hdl = self.peer.alloc((rows, plan.geom["Mpad"]), torch.int64, zero=True)
                elif rows != n:
                    plan.Jint = torch.zeros(rows, plan.geom["Mpad"], dtype=torch.int64, device=self.K.device)
                return plan
            if share is not None and (share.S, share.stream) != (digits, bool(stream)):
                share = None
            return self.K.gram_tc_plan(ds.Xp, self.D, n, digits, stream=stream, share=share)
        except ValueError:
            return None                  # signed design: the digits of Z assume x >= 0

    def _tc_plan(self, ds, n, tag=None):
        """The dataset's tensor-core Gram plan for n local neurons, or None when the FP64 kernel should run.
        tag: neuron group of the overlapped single-GPU sweep (its own per-sweep buffers, the operand shared)."""
        if self.gram_mode == "fp64" or not hasattr(self.K, "gram_tc_plan"):
            return None
        key = self._tc_key(n, tag)
        if key not in ds.buffers:
            M = self.D * (self.D + 1) // 2
            T_eff = ds.T_global / self.comm.world if self._time_sharded() else ds.T
            want = self.gram_mode == "tc" or float(M) * T_eff * n >= self.TC_MIN_WORK
            share = None
            if tag is not None:
                share = next((v for k, v in ds.buffers.items() if k[0] == "tc_plan" and v is not None), None)
            plan = self._tc_build(ds, n, self.gram_digits, share=share) if want else None
            if plan is not None:
                plan.key = key
            if plan is None and self.gram_mode == "tc":
                raise RuntimeError("gram='tc' needs a non-negative design and %.1f GB of free HBM"
                                   % (gram_tc_bytes(self.D, n, ds.T, self.gram_digits,
                                                    stream=self.gram_stream is True) / 1e9))
            ds.buffers[key] = plan
        return ds.buffers[key]

    def _tc_verified(self, ds, omega, n, plan, tag=None):
        """First use of a plan: run it beside the FP64 kernel on the sweep's own omega and keep it only if the
        two agree to TC_ACCEPT on every lower-triangle entry; otherwise add a digit, then give up (FP64).  The
        deviation of the integer-digit product depends on the data (sparser trains -> smaller entries relative to
        the fixed-point scale), so it is measured, not assumed.  Time-sharded: what is compared are the COMPLETE
        Grams of each rank's neuron block after the reduce-scatter (the integer totals are the single-GPU ones, so
        the decision is too); the worst rank decides for all."""
        key = self._tc_key(n, tag)
        rows = max(self.scan_hi - self.scan_lo, 1) if self._time_sharded() else n
        J_ref = self._gram_fp64(ds, omega, n, self.K.zeros(rows, self.ldx, self.ldx))
        tril = torch.tril(torch.ones(self.D, self.D, dtype=torch.bool, device=J_ref.device))
        while plan is not None:
            J_tc = self._gram_tc(plan, omega, self.K.zeros(rows, self.ldx, self.ldx))
            a, b = J_tc[:, :self.D, :self.D], J_ref[:, :self.D, :self.D]
            dev = torch.where(tril & (b != 0), (a - b).abs() / b.abs(), torch.zeros_like(b))
            worst = dev.max().reshape(1)
            if self._time_sharded():
                self.comm.all_reduce_max(worst)
            plan.max_rel_dev = float(worst)
            del J_tc, dev
            if plan.max_rel_dev <= self.TC_ACCEPT:
                plan.verified = True
                break
            import warnings
            warnings.warn("tensor-core Gram with %d digits deviates from the FP64 kernel by %.2e (> %.1e) on the first "
                          "sweep: trying %s" % (plan.S, plan.max_rel_dev, self.TC_ACCEPT,
                                                "5 digits" if plan.S < 5 else "the FP64 kernel"))
            digits = plan.S + 1
            ds.buffers[key] = plan = None
            torch.cuda.empty_cache()
            if digits <= 5:
                plan = self._tc_build(ds, n, digits)
                if plan is not None:
                    plan.key = key
        if plan is None and self.gram_mode == "tc":
            raise RuntimeError("gram='tc': the integer-digit Gram deviates from FP64 by more than %.1e" % self.TC_ACCEPT)
        ds.buffers[key] = plan
        return plan

    def _gram_tc(self, plan, omega, J):
        e0 = self._mark("gram_tc_slice")
        plan.slice_omega(omega)
        e1 = self._mark("gram_tc_slice", e0)
        plan.mma()
        e2 = self._mark("gram_tc_mma", e1)
        if self._time_sharded() and getattr(plan, "peer_hdl", None) is not None:
            # exact int64 reduce-scatter fused into the finalize pass (csrc/gram_tc.cu gram_tc_finalize_peers_kernel):
            # barrier (every rank's partial sums are complete) -> read, add, scale -> barrier (the buffers may be reused)
            nS = self.scan_hi - self.scan_lo
            self.peer.barrier(plan.peer_hdl, channel=1)
            if nS > 0:
                plan.finalize_peers(J, plan.peer_hdl, self.comm.world, self.scan_lo, nS,
                                    plan.omax[self.scan_lo:self.scan_hi])
            self.peer.barrier(plan.peer_hdl, channel=1)
        elif self._time_sharded():
            nS = self.scan_hi - self.scan_lo
            Jloc = self.comm.reduce_scatter_rows(plan.Jint)
            e2 = self._mark("gram_reduce_scatter", e2)
            if nS > 0:
                plan.finalize(J, Jint=Jloc[:nS], omax=plan.omax[self.scan_lo:self.scan_hi])
        else:
            plan.finalize(J)
        self._mark("gram_tc_finalize", e2)
        return J

    def _gram_fp64(self, ds, omega, n, J):
        if self._time_sharded():
            Jp = self._wsbuf("J_partial", (self.comm.world * self.n_max, self.ldx, self.ldx), zero=True)
            self.K.weighted_gram(ds.Xp, omega, self.D, n, J=Jp)
            e0 = self._mark("gram_reduce_scatter")
            Jloc = self.comm.reduce_scatter_rows(Jp)
            self._mark("gram_reduce_scatter", e0)
            J.copy_(Jloc[:J.shape[0]])
        else:
            self.K.weighted_gram(ds.Xp, omega, self.D, n, J=J)
        return J

    TC_RECHECK_EVERY = 16       # sweeps between spot checks of the tensor-core Gram against the FP64 kernel
    TC_RECHECK_NEURONS = 2      # neurons compared per spot check (rotating through the local block)
    TC_RECHECK_TILE_STRIDE = 8  # every 8th (i, j) tile of the FP64 kernel's tile list per check (rotating offset): a check
                                # costs 1 ms instead of 8 ms at cfg3 and after 8 checks every entry has been compared

    def _tc_spot_check(self, ds, omega, n, plan):
        """The first-sweep check (_tc_verified) sees one omega; the deviation of the fixed-point Gram depends on the
        data, and omega changes every sweep.  So every TC_RECHECK_EVERY-th use of a plan, TC_RECHECK_NEURONS neurons
        (rotating) are recomputed by the FP64 kernel from the sweep's own omega and compared entry by entry with what
        the tensor-core kernel has just produced (time-sharded: the slab's partial sums on both sides).  Nothing
        waits: the worst relative deviation goes to pinned memory behind an event and is read at the start of a later
        sweep (_tc_poll), where a deviation above TC_ACCEPT retires the plan (FP64 kernel from then on) with a warning.
        Time-sharded ranks all-reduce (max) the figure so that they retire the plan at the same sweep."""
        k = min(self.TC_RECHECK_NEURONS, n)
        lo = (plan.checks * k) % max(n - k + 1, 1)
        plan.checks += 1
        om = self._wsbuf("tc_check_omega", (ds.T, pad_ldn(k)), zero=True)
        om[:, :k] = omega[:, lo:lo + k]
        tiles = self.K.gram_tiles(self.D, 0)
        stride = max(1, min(self.TC_RECHECK_TILE_STRIDE, tiles.shape[0]))
        tiles = tiles[(plan.checks - 1) % stride::stride].contiguous()
        # entries outside the chosen tiles stay zero in J_ref and are skipped by the comparison below (b != 0)
        J_ref = self.K.weighted_gram(ds.Xp, om, self.D, k, J=self._wsbuf("tc_check_Jref", (k, self.ldx, self.ldx)).zero_(),
                                     tiles=tiles)
        J_tc = plan.finalize(self._wsbuf("tc_check_J", (k, self.ldx, self.ldx), zero=True),
                             Jint=plan.Jint[lo:lo + k], omax=plan.omax[lo:lo + k])
        if "tril" not in self._ws:
            self._ws["tril"] = torch.tril(torch.ones(self.D, self.D, dtype=torch.bool, device=J_ref.device))
        a, b = J_tc[:, :self.D, :self.D], J_ref[:, :self.D, :self.D]
        dev = torch.where(self._ws["tril"] & (b != 0), (a - b).abs() / b.abs(), torch.zeros_like(b))
        worst = dev.max().reshape(1)
        if self._time_sharded():
            self.comm.all_reduce_max(worst)
        stage = self._pinned("tc_check", (1,), torch.float64)
        stage.copy_(worst, non_blocking=True)
        ev = torch.cuda.Event() if self.K.device.type == "cuda" else None
        if ev is not None:
            ev.record()
        self._tc_pending_check = (ev, stage, ds, n, plan, self.calls)

    def _tc_poll(self, force=False):
        """Read a finished spot check (see _tc_spot_check).  Multi-rank runs decide at a fixed sweep distance from the
        check (waiting for the event if need be) so that every rank switches kernels at the same sweep."""
        pend = getattr(self, "_tc_pending_check", None)
        if pend is None:
            return
        ev, stage, ds, n, plan, at_call = pend
        if self.comm.world > 1 or force:
            if not force and self.calls < at_call + 2:
                return
            if ev is not None:
                ev.synchronize()
        elif ev is not None and not ev.query():
            return
        self._tc_pending_check = None
        worst = float(stage[0])
        plan.max_rel_dev_spot = max(worst, plan.max_rel_dev_spot or 0.0)
        # time-sharded: what was compared are the slabs' PARTIAL sums, whose relative deviation is sqrt(world) times
        # that of the totals the scan consumes (independent rounding errors add in quadrature, the sums add linearly)
        limit = self.TC_ACCEPT * (float(self.comm.world) ** 0.5 if self._time_sharded() else 1.0)
        if not worst <= limit:
            import warnings
            warnings.warn("tensor-core Gram deviates from the FP64 kernel by %.2e (> %.1e) on a spot check: "
                          "falling back to the FP64 kernel for this data set" % (worst, limit))
            if self.gram_mode == "tc":
                raise RuntimeError("gram='tc': spot check of the integer-digit Gram failed (%.2e)" % worst)
            ds.buffers[getattr(plan, "key", ("tc_plan", n))] = None
            self._pending = None

    def weighted_gram(self, ds, omega, n, J, tag=None):
        """J (rows of the scan block) = this dataset's weighted Gram.  Neuron-sharded / single GPU: n = local
        neurons, J has n rows.  Time-sharded: n = N, the slab's partial sums of ALL neurons are reduce-scattered
        over the neuron axis (exact int64 sums on the tensor-core path, FP64 otherwise) and J receives the complete
        Grams of this rank's neuron block."""
        plan = self._tc_plan(ds, n, tag)
        if plan is not None and not plan.verified:
            plan = self._tc_verified(ds, omega, n, plan, tag)
        if plan is not None:
            self._gram_tc(plan, omega, J)
            plan.uses += 1
            if self.TC_RECHECK_EVERY and plan.uses % self.TC_RECHECK_EVERY == 0 \
                    and getattr(self, "_tc_pending_check", None) is None:
                self._tc_spot_check(ds, omega, n, plan)
            return J
        return self._gram_fp64(ds, omega, n, J)

    # ------------------------------------------------------------------ coefficients
    def build_Wt(self, A, W, b, lo, hi):
        """Host (N,N) bool, (N,N,B), (N,) -> device Wt (ldx, ldn) for neurons [lo, hi)
        (regression.py:199: coefficients are a o W, then the bias)."""
        n = hi - lo
        NB = self.N * self.B
        Wt = np.zeros((self.ldx, pad_ldn(n)))
        Wt[:NB, :n] = (A[lo:hi, :, None] * W[lo:hi]).reshape(n, NB).T
        Wt[NB, :n] = b[lo:hi]
        self.h2d_bytes += Wt.nbytes
        return self.K.to_device(Wt)

    def build_Wt_device(self, state, lo, hi, tag=None):
        """The same from the device copy of the new state (rows [a (N) | W (N*B) | b | status] per neuron), so the
        next sweep's psi / PG / Gram can be enqueued without a host round trip."""
        n = hi - lo
        N, NB = self.N, self.N * self.B
        Wt = self._wsbuf("Wt" if tag is None else "Wt@%s" % (tag,), (self.ldx, pad_ldn(n)), zero=True)
        rows = state[lo:hi]
        Wt[:NB, :n] = (rows[:, :N].unsqueeze(-1) * rows[:, N:N + NB].reshape(n, N, self.B)).reshape(n, NB).t()
        Wt[NB, :n] = rows[:, N + NB]
        return Wt

    # ------------------------------------------------------------------ sweep-invariant h
    def _h_lkhd(self, ds, lo, hi):
        """h[j, :] = Xp^T (Y[:, lo+j] - 1/2) for j < hi-lo (regression.py:259-260 with :510-511)."""
        key = (lo, hi)
        if key not in ds.h_cache:
            n = hi - lo
            kap = self.K.zeros(ds.T, pad_ldn(n))
            kap[:, :n] = ds.Y[:, lo:hi] - 0.5
            ds.h_cache[key] = self.K.xt_kappa(ds.Xp, kap, self.D, n)
            del kap
        return ds.h_cache[key]

    # ------------------------------------------------------------------ the sweep
    # Philox call ids of sweep k: k * CALL_STRIDE + (index of the data set) for the PG draws, k * CALL_STRIDE +
    # CALL_STRIDE - 1 for the scan's permutation / uniforms / normals.
    CALL_STRIDE = 64

    def _check_call_ids(self, datasets):
        if len(datasets) >= self.CALL_STRIDE - 1:
            raise ValueError("at most %d data sets per model: the Philox call ids of a sweep are laid out in strides "
                             "of %d (PG draws of data set i, then the scan)" % (self.CALL_STRIDE - 2, self.CALL_STRIDE))

    def _augment(self, datasets, Wt, call_base, blk=None):
        """psi -> omega ~ PG(1, psi) -> J for the scan block, for every dataset (regression.py:496-508, :225-262).
        Everything is enqueued on the current stream; nothing here waits for the device.
        blk = (lo, hi, tag): the same for one neuron group of the overlapped single-GPU sweep, with the group's own
        buffers (the PG draws are keyed by the global (bin, neuron) index, so grouping does not change them)."""
        K, N, D, ldx = self.K, self.N, self.D, self.ldx
        p_lo, p_hi = (self.psi_lo, self.psi_hi) if blk is None else blk[:2]
        tag = None if blk is None else blk[2]
        sfx = "" if tag is None else "@%s" % (tag,)
        nP = p_hi - p_lo
        nS = self.scan_hi - self.scan_lo if blk is None else nP
        ldn = Wt.shape[1]
        J = self._wsbuf("J" + sfx, (max(nS, 1) if self._time_sharded() else nP, ldx, ldx), zero=True)
        for di, ds in enumerate(datasets):
            psi = self._buf(ds, "psi" + sfx, (ds.T, ldn))
            omega = self._buf(ds, "omega" + sfx, (ds.T, ldn), zero=True)
            e0 = self._mark("activation")
            K.activation(ds.Xp, Wt, D, nP, out=psi)
            e1 = self._mark("activation", e0)
            if self.inject is None:
                K.pg_draw(psi, nP, omega, self.seed, call_base + di, ds.t_off, p_lo, N)
            else:
                om = np.asarray(self.inject["omega"][di])[ds.t_off:ds.t_off + ds.T, p_lo:p_hi]
                omega[:, :nP] = K.to_device(om)
            e2 = self._mark("pg_draw", e1)
            if di == 0:
                self.weighted_gram(ds, omega, nP, J, tag)
            else:
                Jd = self._wsbuf("J_extra" + sfx, tuple(J.shape), zero=True)
                self.weighted_gram(ds, omega, nP, Jd, tag)
                J += Jd
            self._mark("weighted_gram", e2)
        self._aug_end = self._mark("idle_before_scan")        # start of the gap until the scan kernel is enqueued
        return J[:nS] if self._time_sharded() else J

    def _prelaunched(self, datasets, A, W, b):
        """The Gram enqueued at the end of the previous sweep, if it was computed from exactly this state and these
        datasets (users may edit regressions[n].a / .W / .b or swap data between sweeps: then it is discarded)."""
        pend, self._pending = self._pending, None
        if pend is None or self.inject is not None:
            return None
        if len(pend) != 5:
            return None                                      # left by the overlapped sweep (per-group Grams)
        ds_ids, A0, W0, b0, J = pend
        if ds_ids != [id(ds) for ds in datasets]:
            return None
        if not (np.array_equal(A0, A) and np.array_equal(W0, W) and np.array_equal(b0, b)):
            return None
        return J

    def _pinned(self, name, shape, dtype):
        key = ("pinned", name, tuple(shape), dtype)
        if key not in self._ws:
            self._ws[key] = torch.empty(*shape, dtype=dtype, pin_memory=(self.K.device.type == "cuda"))
        return self._ws[key]

    def _upload_priors(self, pr, a_host, do_scan):
        """One pinned staging buffer and one asynchronous copy for all prior terms of the scan block (and one for the
        two byte arrays) instead of eight pageable, stream-synchronising copies."""
        K = self.K
        names = ("J0w", "h0w", "J0b", "h0b", "cprior", "logit_rho")
        sizes = [pr[k].size for k in names]
        stage = self._pinned("prior", (sum(sizes),), torch.float64)
        host = stage.numpy()
        off = 0
        for k, sz in zip(names, sizes):
            host[off:off + sz] = pr[k].ravel()
            off += sz
        dev = self._wsbuf("prior_dev", (sum(sizes),))
        dev.copy_(stage, non_blocking=True)
        out, off = {}, 0
        for k, sz in zip(names, sizes):
            out[k] = dev[off:off + sz].view(pr[k].shape)
            off += sz
        nb = a_host.size + do_scan.size
        bstage = self._pinned("prior_bytes", (nb,), torch.uint8)
        bh = bstage.numpy()
        bh[:a_host.size] = a_host.ravel()
        bh[a_host.size:] = do_scan.astype(np.uint8)
        bdev = self._wsbuf("prior_bytes_dev", (nb,), dtype=torch.uint8)
        bdev.copy_(bstage, non_blocking=True)
        self.h2d_bytes += stage.numel() * 8 + nb
        return out, bdev[:a_host.size].view(a_host.shape), bdev[a_host.size:]

    # ------------------------------------------------------------------ overlapped single-GPU sweep
    OVERLAP_MIN_N = 128

    def _overlap_groups(self, datasets):
        """Neuron groups of the overlapped sweep, or None when the sweep runs as one block (multi-GPU runs, small
        models, injected randomness, no data, on-device rate moments -- they read the psi buffer of the whole block)."""
        if not (self.overlap and self.pipeline and self.inject is None and datasets and self.comm.world == 1
                and self.K.device.type == "cuda" and self.N >= max(2, self.overlap_min_n)):
            return None
        if self.moments is not None and self.moments.rates:
            return None
        h = (self.N + 1) // 2
        return [(0, h, 0), (h, self.N, 1)]

    def _overlap_streams(self):
        if self._ovl is None:
            dev = self.K.device
            # the scan stream has the higher priority: when the Gram of a group retires, the scan CTAs of that group
            # take their SMs before the next group's psi / PG / Gram kernels fill the machine
            self._ovl = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev, priority=-1))
        return self._ovl

    def _overlap_drain(self):
        """Order the current stream behind whatever the two side streams still hold (mode switches, scoring calls)."""
        if self._ovl is not None:
            cur = torch.cuda.current_stream(self.K.device)
            for st in self._ovl:
                cur.wait_stream(st)

    def _sweep_overlapped(self, groups, datasets, A, W, b, hypers, call_base):
        """The single-GPU sweep as TWO neuron groups on two streams.  Regressions are independent given the
        hyper-parameters (models.py:169-171 is a loop over them), so the groups only meet in the host network step:

            stream `aug` :  psi/PG/Gram_k(g1)   | psi/PG/Gram_k+1(g0) | psi/PG/Gram_k+1(g1) | psi/PG/Gram_k+2(g0) | ...
            stream `scan`:  scan_k(g0)          | scan_k(g1)          | scan_k+1(g0)        | scan_k+1(g1)        | ...
                            `------------- sweep k ------------'      `------------ sweep k+1 ------------'

        One group's scan is a single wave of <= 148 one-CTA neurons, latency bound (~2.5 ms at cfg3) and leaves the
        other SMs idle; here those SMs -- and all of them once the scan CTAs retire -- work on the other group's
        augmentation.  Measured at cfg3 (profiles/r02z_timeline_aug_first.log): psi + PG of the other group finish
        inside the scan's 2.5 ms and the Gram follows at full width -- 2 x 8.0 ms per sweep while the chain is sparse,
        but no gain at its equilibrium density, where the scan is L2-bound (see the note at `self.overlap`).  Should a
        Gram start while scan CTAs still hold SMs, it draws its items from a device counter (gram_tc.cu,
        tc_work_ticket), so CTAs that start late just take fewer items.
        Call k submits aug_k(g1) first -- it needs nothing from the host, and keeps the device busy while the host
        prepares the prior terms (2 ms of numpy at cfg3) -- then scan_k(g0), scan_k(g1) and the NEXT sweep's aug(g0),
        which is on the device while the caller runs its network step.  A host that is late with scan_k(g0) loses the
        overlap (the persistent Gram CTAs of g1 are not preempted), nothing else.  The chain is the one the one-block
        sweep produces: PG draws and scan randomness are keyed by global (bin, neuron) / neuron indices, the Gram sums
        are exact integers, and every other kernel works neuron by neuron."""
        K, N, B, D, ldx = self.K, self.N, self.B, self.D, self.ldx
        NB = N * B
        (lo0, hi0, tag0), (lo1, hi1, tag1) = groups
        s_aug, s_scan = self._overlap_streams()
        main = torch.cuda.current_stream(K.device)
        pend, self._pending = self._pending, None
        aug0 = prev_state = None
        if pend is not None and len(pend) == 6 and pend[5] == "groups":
            ds_ids, A0, W0, b0, (aug0, prev_state), _ = pend
            if ds_ids != [id(ds) for ds in datasets] or not (np.array_equal(A0, A) and np.array_equal(W0, W)
                                                              and np.array_equal(b0, b)):
                aug0 = prev_state = None                     # the user edited the state or swapped data: start over
        s_aug.wait_stream(main)
        s_scan.wait_stream(main)
        for ds in datasets:
            if "side_streams" not in ds.buffers:
                # allocated on the caller's stream, used on the side streams: a later free must wait for them
                for t in (ds.Xp, ds.Y):
                    t.record_stream(s_aug)
                    t.record_stream(s_scan)
                ds.buffers["side_streams"] = True

        def augment(lo, hi, tag, src_state, base):
            """psi / PG / Gram of one group on the aug stream, from device state rows (or the host state)."""
            with torch.cuda.stream(s_aug):
                Wt = self.build_Wt(A, W, b, lo, hi) if src_state is None else \
                    self.build_Wt_device(src_state, lo, hi, tag=tag)
                J = self._augment(datasets, Wt, base, blk=(lo, hi, tag))
                ev = torch.cuda.Event()
                ev.record()
            return J, ev

        if aug0 is None:
            aug0 = augment(lo0, hi0, tag0, None, call_base)
        # group 1's augmentation of THIS sweep: from the rows the previous sweep left on the device
        aug1 = augment(lo1, hi1, tag1, prev_state, call_base)
        width = self._state_width()
        with torch.cuda.stream(s_scan):
            state = self._wsbuf("state@%d" % (self.calls & 1), (N, width), zero=True)
            h_S = self._h_for_scan(datasets)
            pr = prior_arrays(hypers["rho"], hypers["mu_w"], hypers["S_w"], hypers["mu_b"], hypers["S_b"])
            do_scan = pr.pop("do_scan")
            a_host = np.array(A, dtype=np.uint8)
            det = ~do_scan
            a_host[det] = np.round(hypers["rho"][det]).astype(np.uint8)      # regression.py:274-275
            prior, a_dev, do_scan_dev = self._upload_priors(pr, a_host, do_scan)
            perm, us, z = K.scan_randomness(N, B, N, 0, self.seed, call_base + self.CALL_STRIDE - 1)

        def scan(lo, hi, tag, J_g, ev_aug):
            n = hi - lo
            with torch.cuda.stream(s_scan):
                s_scan.wait_event(ev_aug)
                P_ws = self._wsbuf("P@%s" % (tag,), (n * D * D,))
                e3 = self._mark("spike_slab")
                W_new, b_new, _, _, status = K.spike_slab_update(
                    N, B, J_g, h_S[lo:hi], {k: v[lo:hi] for k, v in prior.items()}, perm[lo:hi], us[lo:hi], z[lo:hi],
                    do_scan_dev[lo:hi], a_dev[lo:hi], P_ws=P_ws)
                self._mark("spike_slab", e3)
                state[lo:hi, :N] = a_dev[lo:hi]
                state[lo:hi, N:N + NB] = W_new.reshape(n, NB)
                state[lo:hi, N + NB] = b_new
                state[lo:hi, N + NB + 1] = status
                ev = torch.cuda.Event()
                ev.record()
            return ev

        ev_scan0 = scan(lo0, hi0, tag0, *aug0)
        scan(lo1, hi1, tag1, *aug1)
        with torch.cuda.stream(s_scan):
            stage = self._pinned("state", (N, width), torch.float64)
            stage.copy_(state, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
            if self.moments is not None:
                self.moments.add_state(state, N + NB + 1)
        # group 0's augmentation of the NEXT sweep, behind its scan: on the device while the host does its part
        s_aug.wait_event(ev_scan0)
        nxt0 = augment(lo0, hi0, tag0, state, call_base + self.CALL_STRIDE)
        main.wait_event(done)            # the caller's stream (and its timing events) sees the sweep's result
        done.synchronize()
        return self._finish_sweep(stage, datasets, (nxt0, state), groups=True)

    def _finish_sweep(self, stage, datasets, pend_J, groups=False):
        """Host side of the end of a sweep: unpack the pinned state rows, surface failed neurons, remember what the
        pre-launched augmentation was computed from."""
        N, NB, B = self.N, self.N * self.B, self.B
        host = stage.numpy()
        self.d2h_bytes += host.nbytes
        A_out = host[:, :N] != 0
        W_out = host[:, N:N + NB].reshape(-1, N, B).copy()
        b_out = host[:, N + NB].copy()
        st = host[:, N + NB + 1]
        if st.any():
            self._pending = None
            bad = np.nonzero(st)[0]
            raise FloatingPointError("spike-and-slab update lost positive definiteness for neuron(s) %s "
                                     "(ill-conditioned posterior precision)" % bad[:8].tolist())
        if pend_J is not None:
            # private copies: the caller's regressions alias the returned arrays and may edit them in place
            self._pending = ([id(ds) for ds in datasets], A_out.copy(), W_out.copy(), b_out.copy(), pend_J) \
                + (("groups",) if groups else ())
        return A_out, W_out, b_out

    def sweep(self, datasets, A, W, b, hypers):
        """One resample_regressions() (models.py:169-171) for all neurons.
        A (N,N) bool, W (N,N,B), b (N,) host state;  hypers: dict rho (N,N), mu_w (N,N,B), S_w (N,N,B,B),
        mu_b (N,), S_b (N,) host arrays, row n = regression n.  Returns new host (A, W, b).

        Sweeps are software-pipelined: psi / PG / Gram depend only on (a, W, b), not on the hyper-parameters, so as
        soon as the scan of sweep k has produced the new state on the device the augmentation of sweep k+1 is
        enqueued behind it, and the host part (D2H of the state, network step, prior terms) overlaps with it."""
        K, N, B, D, ldx = self.K, self.N, self.B, self.D, self.ldx
        comm = self.comm
        p_lo, p_hi = self.psi_lo, self.psi_hi
        s_lo, s_hi = self.scan_lo, self.scan_hi
        nP, nS = p_hi - p_lo, s_hi - s_lo
        NB = N * B
        self.calls += 1
        call_base = self.calls * self.CALL_STRIDE
        self._check_call_ids(datasets)
        self._tc_poll()
        groups = self._overlap_groups(datasets)
        if groups is not None:
            return self._sweep_overlapped(groups, datasets, A, W, b, hypers, call_base)
        self._overlap_drain()

        J_S = h_S = None
        if datasets:
            J_S = self._prelaunched(datasets, A, W, b)
            if J_S is None and nP > 0:
                J_S = self._augment(datasets, self.build_Wt(A, W, b, p_lo, p_hi), call_base)
        # new state rows [a (N) | W (N*B) | b | status], one row per neuron of the scan block (padded to n_max)
        width = self._state_width()
        state = K.zeros(self.n_max if comm.world > 1 else nS, width)
        if nS > 0:
            if datasets:
                h_S = self._h_for_scan(datasets)
            else:
                # no data: the posterior is the prior (regression.py:237, datas = [] leaves J_lkhd = h_lkhd = 0)
                J_S = self._wsbuf("J_nodata", (nS, ldx, ldx), zero=True)
                h_S = self._wsbuf("h_nodata", (nS, ldx), zero=True)
            pr = prior_arrays(hypers["rho"][s_lo:s_hi], hypers["mu_w"][s_lo:s_hi], hypers["S_w"][s_lo:s_hi],
                              hypers["mu_b"][s_lo:s_hi], hypers["S_b"][s_lo:s_hi])
            do_scan = pr.pop("do_scan")
            a_host = np.array(A[s_lo:s_hi], dtype=np.uint8)
            # deterministic sparsity: a = round(rho) (regression.py:274-275)
            det = ~do_scan
            a_host[det] = np.round(hypers["rho"][s_lo:s_hi][det]).astype(np.uint8)
            prior, a_dev, do_scan_dev = self._upload_priors(pr, a_host, do_scan)
            if self.inject is None:
                perm, us, z = K.scan_randomness(N, B, nS, s_lo, self.seed, call_base + self.CALL_STRIDE - 1)
            else:
                perm = K.to_device(np.asarray(self.inject["perm"][s_lo:s_hi], dtype=np.int32))
                us = K.to_device(np.asarray(self.inject["us"][s_lo:s_hi], dtype=np.float64))
                z = K.to_device(np.asarray(self.inject["z"][s_lo:s_hi], dtype=np.float64))
            P_ws = self._wsbuf("P", (nS * D * D,))
            e3 = self._mark("spike_slab")
            if e3 is not None and getattr(self, "_aug_end", None) is not None:
                self.profile.setdefault("idle_before_scan", []).append((self._aug_end, e3))
                self._aug_end = None
            W_new, b_new, _, _, status = K.spike_slab_update(N, B, J_S, h_S, prior, perm, us, z, do_scan_dev, a_dev,
                                                             P_ws=P_ws)
            e4 = self._mark("spike_slab", e3)
            state[:nS, :N] = a_dev
            state[:nS, N:N + NB] = W_new.reshape(nS, NB)
            state[:nS, N + NB] = b_new
            state[:nS, N + NB + 1] = status
        # exchange: ONE all-gather of the new rows (the only collective of the neuron-sharded sweep)
        if comm.world > 1:
            state = self._gather_state(state)[:N]
        # state -> host through pinned memory, asynchronously ...
        stage = self._pinned("state", tuple(state.shape), torch.float64)
        stage.copy_(state, non_blocking=True)
        done = None
        if K.device.type == "cuda":
            done = torch.cuda.Event()
            done.record()
        # ... while the device already starts on the next sweep's psi / PG / Gram
        pend_J = None
        mom = self.moments
        if mom is not None and state.shape[0] == N:
            mom.add_state(state, N + NB + 1)
        if self.pipeline and self.inject is None and datasets and nP > 0 and state.shape[0] == N:
            Wt_next = self.build_Wt_device(state, p_lo, p_hi)
            if nS > 0 and datasets:
                self._mark("exchange", e4)
            pend_J = self._augment(datasets, Wt_next, call_base + self.CALL_STRIDE)
            if mom is not None and mom.rates:              # psi of the NEW state is already in the buffers
                for di, ds in enumerate(datasets):
                    mom.add_rates(di, self._buf(ds, "psi", (ds.T, Wt_next.shape[1])), nP)
        elif mom is not None and mom.rates and datasets and nP > 0 and state.shape[0] == N:
            Wt_next = self.build_Wt_device(state, p_lo, p_hi)
            for di, ds in enumerate(datasets):
                psi = self._buf(ds, "psi", (ds.T, Wt_next.shape[1]))
                K.activation(ds.Xp, Wt_next, D, nP, out=psi)
                mom.add_rates(di, psi, nP)
        if done is not None:
            done.synchronize()
        return self._finish_sweep(stage, datasets, pend_J)

    # ------------------------------------------------------------------ Gaussian observations
    def _gaussian_stats(self, datasets):
        """Sweep-invariant X~^T X~ (ldx, ldx), X~^T Y (N, ldx) and the number of bins, summed over the data sets
        (regression.py:225-262 with omega = 1/eta, kappa = y/eta factored out: both are sums over time only)."""
        K, N, D = self.K, self.N, self.D
        G = H = None
        T = 0
        for ds in datasets:
            # cached ON the data set (an id()-keyed engine cache could be hit by a new object reusing the address)
            if "gauss" not in ds.buffers:
                ones = K.zeros(ds.T, pad_ldn(1))
                ones[:, 0] = 1.0
                Gd = K.weighted_gram(ds.Xp, ones, D, 1)[0]
                Yp = K.zeros(ds.T, pad_ldn(N))
                Yp[:, :N] = ds.Y
                ds.buffers["gauss"] = (Gd, K.xt_kappa(ds.Xp, Yp, D, N))
            Gd, Hd = ds.buffers["gauss"]
            G = Gd if G is None else G + Gd
            H = Hd if H is None else H + Hd
            T += ds.T
        return G, H, T

    def residual_ss(self, datasets, A, W, b):
        """sum_t (y_{t,n} - psi_{t,n})^2 for every neuron, over all data sets: (N,) host array."""
        K, N, D = self.K, self.N, self.D
        Wt = self.build_Wt(A, W, b, 0, N)
        rss = K.zeros(N)
        for ds in datasets:
            psi = self._buf(ds, "psi", (ds.T, Wt.shape[1]))
            K.activation(ds.Xp, Wt, D, N, out=psi)
            rss += ((ds.Y - psi[:, :N]) ** 2).sum(0)
        self.d2h_bytes += 8 * N
        return rss.cpu().numpy()

    def sweep_gaussian(self, datasets, A, W, b, hypers, eta):
        """resample_regressions() for Gaussian observations (regression.py:426-430 without the eta draw, which is
        the caller's host step): J_n = X~^T X~ / eta_n, h_n = X~^T y_n / eta_n from the cached sums, then the same
        spike-and-slab kernel.  Single process or neuron-sharded.  Returns (A, W, b, rss) with rss_n the residual
        sum of squares under the NEW coefficients (what _resample_eta needs, regression.py:432-445)."""
        K, N, B, D, ldx = self.K, self.N, self.B, self.D, self.ldx
        comm = self.comm
        assert not self._time_sharded(), "Gaussian observations: use shard='neuron'"
        s_lo, s_hi = self.scan_lo, self.scan_hi
        nS = s_hi - s_lo
        NB = N * B
        self.calls += 1
        call_base = self.calls * self.CALL_STRIDE
        self._check_call_ids(datasets)
        self._pending = None
        width = self._state_width()
        state = K.zeros(self.n_max if comm.world > 1 else nS, width)
        if nS > 0:
            if datasets:
                G, H, _ = self._gaussian_stats(datasets)
                inv_eta = K.to_device(1.0 / np.asarray(eta, dtype=np.float64)[s_lo:s_hi])
                J_S = G.unsqueeze(0) * inv_eta[:, None, None]
                h_S = H[s_lo:s_hi] * inv_eta[:, None]
            else:
                # no data: the posterior is the prior, as in sweep() (regression.py:237 with datas = [])
                J_S = self._wsbuf("J_nodata", (nS, ldx, ldx), zero=True)
                h_S = self._wsbuf("h_nodata", (nS, ldx), zero=True)
            pr = prior_arrays(hypers["rho"][s_lo:s_hi], hypers["mu_w"][s_lo:s_hi], hypers["S_w"][s_lo:s_hi],
                              hypers["mu_b"][s_lo:s_hi], hypers["S_b"][s_lo:s_hi])
            do_scan = pr.pop("do_scan")
            a_host = np.array(A[s_lo:s_hi], dtype=np.uint8)
            det = ~do_scan
            a_host[det] = np.round(hypers["rho"][s_lo:s_hi][det]).astype(np.uint8)
            prior, a_dev, do_scan_dev = self._upload_priors(pr, a_host, do_scan)
            if self.inject is None:
                perm, us, z = K.scan_randomness(N, B, nS, s_lo, self.seed, call_base + self.CALL_STRIDE - 1)
            else:
                perm = K.to_device(np.asarray(self.inject["perm"][s_lo:s_hi], dtype=np.int32))
                us = K.to_device(np.asarray(self.inject["us"][s_lo:s_hi], dtype=np.float64))
                z = K.to_device(np.asarray(self.inject["z"][s_lo:s_hi], dtype=np.float64))
            P_ws = self._wsbuf("P", (nS * D * D,))
            W_new, b_new, _, _, status = K.spike_slab_update(N, B, J_S.contiguous(), h_S.contiguous(), prior, perm,
                                                             us, z, do_scan_dev, a_dev, P_ws=P_ws)
            state[:nS, :N] = a_dev
            state[:nS, N:N + NB] = W_new.reshape(nS, NB)
            state[:nS, N + NB] = b_new
            state[:nS, N + NB + 1] = status
        if comm.world > 1:
            state = self._gather_state(state)[:N]
        host = state.cpu().numpy()
        self.d2h_bytes += host.nbytes
        if host[:, N + NB + 1].any():
            bad = np.nonzero(host[:, N + NB + 1])[0]
            raise FloatingPointError("spike-and-slab update lost positive definiteness for neuron(s) %s" % bad[:8].tolist())
        A_out = host[:, :N] != 0
        W_out = host[:, N:N + NB].reshape(-1, N, B).copy()
        b_out = host[:, N + NB].copy()
        if self.moments is not None:
            self.moments.add_state(state, N + NB + 1)
        rss = self.residual_ss(datasets, A_out, W_out, b_out) if datasets else np.zeros(N)
        return A_out, W_out, b_out, rss

    def activations(self, ds, A, W, b):
        """(T, N) psi for one data set: the mean of a Gaussian regression (regression.py:429-430)."""
        N = self.N
        Wt = self.build_Wt(A, W, b, 0, N)
        psi = self.K.activation(ds.Xp, Wt, self.D, N)
        out = psi[:, :N].cpu().numpy()
        self.d2h_bytes += out.nbytes
        return out

    def _h_for_scan(self, datasets):
        """Likelihood h for the scan block, summed over datasets (and over time slabs when time-sharded)."""
        s_lo, s_hi = self.scan_lo, self.scan_hi
        if self.shard == "time" and self.comm.world > 1:
            # the slab sums, all-reduced once per data set and kept ON the data set (h_cache): replacing a data_list
            # entry can then never leave a stale vector behind
            h = None
            for ds in datasets:
                key = ("time_total", s_lo, s_hi)
                if key not in ds.h_cache:
                    h_all = self._h_lkhd(ds, 0, self.N).clone()
                    self.comm.all_reduce_sum(h_all)
                    ds.h_cache[key] = h_all[s_lo:s_hi].contiguous()
                    del ds.h_cache[(0, self.N)]
                h = ds.h_cache[key] if h is None else h + ds.h_cache[key]
            return h
        h = None
        for ds in datasets:
            hd = self._h_lkhd(ds, s_lo, s_hi)
            h = hd if h is None else h + hd
        return h

    # ------------------------------------------------------------------ scoring
    def log_likelihood(self, datasets, A, W, b):
        """models.py:82-96 over the given datasets; partial sums are all-reduced across ranks."""
        lo, hi = self.psi_lo, self.psi_hi
        tot = self.K.zeros(1)
        if hi > lo and datasets:
            Wt = self.build_Wt(A, W, b, lo, hi)
            for ds in datasets:
                tot += self.K.loglik(ds.Xp, Wt, self.D, hi - lo, ds.Y, lo)
        self.comm.all_reduce_sum(tot)
        self.d2h_bytes += 8
        return float(tot.cpu()[0])

    def means(self, ds, A, W, b):
        """(T_local, N) firing probabilities for one dataset (models.py:153-163)."""
        lo, hi = self.psi_lo, self.psi_hi
        n = hi - lo
        mu = self.K.means(ds.Xp, self.build_Wt(A, W, b, lo, hi), self.D, n) if n > 0 else self.K.zeros(ds.T, 0)
        if self.comm.world > 1 and self.shard == "neuron":
            pad = self.n_max - n
            mt = mu.t().contiguous()
            if pad:
                mt = torch.cat([mt, self.K.zeros(pad, ds.T)])
            mu = self.comm.all_gather_rows(mt)[:self.N].t().contiguous()
        out = mu.cpu().numpy()
        self.d2h_bytes += out.nbytes
        return out
