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
import numpy as np
import torch

from .distributed import Comm, block_partition
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


class GibbsEngine(object):
    def __init__(self, N, B, kernels=None, seed=0, comm=None, shard="neuron", gram="auto", gram_digits=4):
        """gram: "fp64" = FP64 DMMA kernel (gram.cu); "tc" = tcgen05 integer-digit kernel (gram_tc.cu), checked
        against the FP64 kernel on the first sweep (<= 5e-10 relative, else one more digit, else FP64);
        "auto" = "tc" when the design is non-negative, the contraction is large enough to matter and the digit
        planes of Z fit in HBM, else "fp64".  gram_digits: radix-256 digits the tc path starts with (4 or 5)."""
        assert shard in ("neuron", "time")
        assert gram in ("auto", "fp64", "tc") and gram_digits in (3, 4, 5)
        self.gram_mode, self.gram_digits = gram, gram_digits
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
    TC_MIN_WORK = 2e11          # pairs * T * neurons below which the FP64 kernel is already sub-millisecond
    TC_ACCEPT = 5e-10           # accepted max relative deviation from the FP64 kernel (stated tolerance 1e-9, 2x margin)

    def _tc_build(self, ds, n, digits):
        need = gram_tc_bytes(self.D, n, ds.T, digits)
        free = torch.cuda.mem_get_info(self.K.device)[0]
        if need >= 0.9 * free:
            return None
        try:
            return self.K.gram_tc_plan(ds.Xp, self.D, n, digits)
        except ValueError:
            return None                  # signed design: the digits of Z assume x >= 0

    def _tc_plan(self, ds, n):
        """The dataset's tensor-core Gram plan for n local neurons, or None when the FP64 kernel should run."""
        if self.gram_mode == "fp64" or not hasattr(self.K, "gram_tc_plan"):
            return None
        key = ("tc_plan", n)
        if key not in ds.buffers:
            M = self.D * (self.D + 1) // 2
            want = self.gram_mode == "tc" or float(M) * ds.T * n >= self.TC_MIN_WORK
            plan = self._tc_build(ds, n, self.gram_digits) if want else None
            if plan is None and self.gram_mode == "tc":
                raise RuntimeError("gram='tc' needs a non-negative design and %.1f GB of free HBM"
                                   % (gram_tc_bytes(self.D, n, ds.T, self.gram_digits) / 1e9))
            ds.buffers[key] = plan
        return ds.buffers[key]

    def _tc_verified(self, ds, omega, n, plan):
        """First use of a plan: run it beside the FP64 kernel on the sweep's own omega and keep it only if the
        two agree to TC_ACCEPT on every lower-triangle entry; otherwise add a digit, then give up (FP64).  The
        deviation of the integer-digit product depends on the data (sparser trains -> smaller entries relative to
        the fixed-point scale), so it is measured, not assumed."""
        key = ("tc_plan", n)
        J_ref = self.K.weighted_gram(ds.Xp, omega, self.D, n)
        tril = torch.tril(torch.ones(self.D, self.D, dtype=torch.bool, device=J_ref.device))
        while plan is not None:
            J_tc = plan.gram(omega)
            a, b = J_tc[:, :self.D, :self.D], J_ref[:, :self.D, :self.D]
            dev = torch.where(tril & (b != 0), (a - b).abs() / b.abs(), torch.zeros_like(b))
            plan.max_rel_dev = float(dev.max())
            del J_tc, dev
            if plan.max_rel_dev <= self.TC_ACCEPT:
                plan.verified = True
                break
            digits = plan.S + 1
            ds.buffers[key] = plan = None
            torch.cuda.empty_cache()
            if digits <= 5:
                plan = self._tc_build(ds, n, digits)
        if plan is None and self.gram_mode == "tc":
            raise RuntimeError("gram='tc': the integer-digit Gram deviates from FP64 by more than %.1e" % self.TC_ACCEPT)
        ds.buffers[key] = plan
        return plan

    def weighted_gram(self, ds, omega, n, J):
        plan = self._tc_plan(ds, n)
        if plan is not None and not plan.verified:
            plan = self._tc_verified(ds, omega, n, plan)
        if plan is not None:
            e0 = self._mark("gram_tc_slice")
            plan.slice_omega(omega)
            e1 = self._mark("gram_tc_slice", e0)
            plan.mma()
            e2 = self._mark("gram_tc_mma", e1)
            plan.finalize(J)
            self._mark("gram_tc_finalize", e2)
        else:
            self.K.weighted_gram(ds.Xp, omega, self.D, n, J=J)
        return J

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
    def sweep(self, datasets, A, W, b, hypers):
        """One resample_regressions() (models.py:169-171) for all neurons.
        A (N,N) bool, W (N,N,B), b (N,) host state;  hypers: dict rho (N,N), mu_w (N,N,B), S_w (N,N,B,B),
        mu_b (N,), S_b (N,) host arrays, row n = regression n.  Returns new host (A, W, b)."""
        K, N, B, D, ldx = self.K, self.N, self.B, self.D, self.ldx
        comm = self.comm
        p_lo, p_hi = self.psi_lo, self.psi_hi
        s_lo, s_hi = self.scan_lo, self.scan_hi
        nP, nS = p_hi - p_lo, s_hi - s_lo
        self.calls += 1
        call_base = self.calls * 64

        J_S = h_S = None
        if nP > 0 and datasets:
            Wt = self.build_Wt(A, W, b, p_lo, p_hi)
            ldn = Wt.shape[1]
            J = self._wsbuf("J", (nP, ldx, ldx), zero=True)
            for di, ds in enumerate(datasets):
                psi = self._buf(ds, "psi", (ds.T, ldn))
                omega = self._buf(ds, "omega", (ds.T, ldn), zero=True)
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
                    self.weighted_gram(ds, omega, nP, J)
                else:
                    Jd = self._wsbuf("J_extra", (nP, ldx, ldx), zero=True)
                    self.weighted_gram(ds, omega, nP, Jd)
                    J += Jd
                self._mark("weighted_gram", e2)
            if self.shard == "time" and comm.world > 1:
                # partial Grams of ALL neurons over the local time slab -> complete Grams of the local block
                Jpad = J
                if N != comm.world * self.n_max:
                    Jpad = self._wsbuf("J_pad", (comm.world * self.n_max, ldx, ldx), zero=True)
                    Jpad[:N] = J
                J_S = comm.reduce_scatter_rows(Jpad)[:nS]
            else:
                J_S = J
        if nS > 0 and datasets:
            h_S = self._h_for_scan(datasets)
            pr = prior_arrays(hypers["rho"][s_lo:s_hi], hypers["mu_w"][s_lo:s_hi], hypers["S_w"][s_lo:s_hi],
                              hypers["mu_b"][s_lo:s_hi], hypers["S_b"][s_lo:s_hi])
            do_scan = pr.pop("do_scan")
            a_host = np.array(A[s_lo:s_hi], dtype=np.uint8)
            # deterministic sparsity: a = round(rho) (regression.py:274-275)
            det = ~do_scan
            a_host[det] = np.round(hypers["rho"][s_lo:s_hi][det]).astype(np.uint8)
            prior = {k: K.to_device(v) for k, v in pr.items()}
            self.h2d_bytes += sum(v.nbytes for v in pr.values()) + a_host.nbytes + do_scan.size
            a_dev = K.to_device(a_host)
            if self.inject is None:
                perm, us, z = K.scan_randomness(N, B, nS, s_lo, self.seed, call_base + 63)
            else:
                perm = K.to_device(np.asarray(self.inject["perm"][s_lo:s_hi], dtype=np.int32))
                us = K.to_device(np.asarray(self.inject["us"][s_lo:s_hi], dtype=np.float64))
                z = K.to_device(np.asarray(self.inject["z"][s_lo:s_hi], dtype=np.float64))
            P_ws = self._wsbuf("P", (nS * D * D,))
            e3 = self._mark("spike_slab")
            W_new, b_new, _, _, status = K.spike_slab_update(N, B, J_S, h_S, prior, perm, us, z,
                                                             K.to_device(do_scan.astype(np.uint8)), a_dev, P_ws=P_ws)
            self._mark("spike_slab", e3)
        else:
            a_dev = K.zeros(0, N, dtype=torch.uint8)
            W_new, b_new = K.zeros(0, N, B), K.zeros(0)
            status = K.zeros(0, dtype=torch.int32)
        # exchange: all-gather the new rows (the only collective of the neuron-sharded sweep)
        if comm.world > 1:
            pad = self.n_max - nS
            if pad:
                a_dev = torch.cat([a_dev, K.zeros(pad, N, dtype=torch.uint8)])
                W_new = torch.cat([W_new, K.zeros(pad, N, B)])
                b_new = torch.cat([b_new, K.zeros(pad)])
                status = torch.cat([status, K.zeros(pad, dtype=torch.int32)])
            a_dev = comm.all_gather_rows(a_dev)[:N]
            W_new = comm.all_gather_rows(W_new)[:N]
            b_new = comm.all_gather_rows(b_new)[:N]
            status = comm.all_gather_rows(status)[:N]
        A_out = a_dev.cpu().numpy().astype(bool)
        W_out = W_new.cpu().numpy()
        b_out = b_new.cpu().numpy()
        st = status.cpu().numpy()
        self.d2h_bytes += A_out.size + W_out.nbytes + b_out.nbytes + st.nbytes
        if st.any():
            bad = np.nonzero(st)[0]
            raise FloatingPointError("spike-and-slab update lost positive definiteness for neuron(s) %s "
                                     "(ill-conditioned posterior precision)" % bad[:8].tolist())
        return A_out, W_out, b_out

    def _h_for_scan(self, datasets):
        """Likelihood h for the scan block, summed over datasets (and over time slabs when time-sharded)."""
        s_lo, s_hi = self.scan_lo, self.scan_hi
        if self.shard == "time" and self.comm.world > 1:
            key = "h_time"
            if key not in self._ws:
                h_all = None
                for ds in datasets:
                    hd = self._h_lkhd(ds, 0, self.N)
                    h_all = hd.clone() if h_all is None else h_all + hd
                self.comm.all_reduce_sum(h_all)
                self._ws[key] = (len(datasets), h_all[s_lo:s_hi].contiguous())
            n_ds, h = self._ws[key]
            if n_ds != len(datasets):
                del self._ws[key]
                return self._h_for_scan(datasets)
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
