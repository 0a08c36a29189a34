"""Typed wrappers around the C ABI (include/pyglm_b200.h) operating on torch CUDA tensors.

torch is used only as the owner of device memory and streams; every computation below is one of the
hand-written sm_100a kernels in pyglm_b200/csrc.  There is no fallback: constructing CudaKernels
without a B200-class device, or without the built library, raises.
"""
import ctypes
import os

import numpy as np
import torch

from . import cabi


def round_up(x, m):
    return (x + m - 1) // m * m


def pad_ldx(D):
    """Row pitch of the padded design matrix / J: D = N*B+1 rounded up to a multiple of 32."""
    return round_up(D, 32)


def pad_ldn(n):
    """Row pitch of psi / omega / Wt: local neuron count rounded up to a multiple of 64."""
    return round_up(n, 64)


class CudaKernels(object):
    """The compute backend.  One instance per device."""

    name = "cuda-sm100a"

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise cabi.PyglmCudaError("pyglm_b200 needs a CUDA device (B200, sm_100a); none is visible and "
                                      "there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.lib = cabi.load()
        with torch.cuda.device(self.device):
            cabi.check(self.lib.pyglm_device_check(), "pyglm_device_check")
        self.launches = 0          # number of kernel-launching ABI calls made (bench.py reports it)
        self._tiles = {}

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)

    @staticmethod
    def lib_ldn(n):
        return pad_ldn(n)

    def empty(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def zeros(self, *shape, dtype=torch.float64):
        return torch.zeros(*shape, dtype=dtype, device=self.device)

    def to_device(self, arr, dtype=None):
        t = torch.from_numpy(np.ascontiguousarray(arr))
        if dtype is not None:
            t = t.to(dtype)
        return t.to(self.device, non_blocking=False)

    def _call(self, name, *args, launches=1):
        with torch.cuda.device(self.device):
            cabi.call(name, *args)
        self.launches += launches

    # ------------------------------------------------------------------ (1) filter
    def filter_spikes(self, S, basis, clip):
        """S (T,N), basis (L,B) device f64 -> padded design Xp (T, ldx)."""
        T, N = S.shape
        L, B = basis.shape
        ldx = pad_ldx(N * B + 1)
        Xp = self.empty(T, ldx)
        self._call("pyglm_filter_spikes", self._p(S), self._p(basis), T, N, L, B, int(bool(clip)),
                   self._p(Xp), ldx, self._stream(), launches=2)
        return Xp

    def pack_design(self, X):
        """dense (T, NB) device f64 -> padded design."""
        T, NB = X.shape
        ldx = pad_ldx(NB + 1)
        Xp = self.empty(T, ldx)
        self._call("pyglm_pack_design", self._p(X), T, NB, self._p(Xp), ldx, self._stream())
        return Xp

    def unpack_design(self, Xp, NB):
        T, ldx = Xp.shape
        X = self.empty(T, NB)
        self._call("pyglm_unpack_design", self._p(Xp), T, NB, ldx, self._p(X), self._stream())
        return X

    # ------------------------------------------------------------------ (5) activation / LL / means
    def activation(self, Xp, Wt, D, n, out=None):
        T, ldx = Xp.shape
        ldn = Wt.shape[1]
        psi = self.empty(T, ldn) if out is None else out
        self._call("pyglm_activation", self._p(Xp), ldx, self._p(Wt), ldn, T, D, n, self._p(psi), psi.shape[1],
                   self._stream())
        return psi

    def loglik(self, Xp, Wt, D, n, Y, y_col0):
        T, ldx = Xp.shape
        ws = self.empty(((T + 127) // 128) * ((n + 7) // 8) + 1)
        ll = self.empty(1)
        self._call("pyglm_loglik", self._p(Xp), ldx, self._p(Wt), Wt.shape[1], T, D, n, self._p(Y), Y.shape[1],
                   y_col0, self._p(ll), self._p(ws), self._stream(), launches=2)
        return ll

    def means(self, Xp, Wt, D, n):
        T, ldx = Xp.shape
        mu = self.empty(T, n)
        self._call("pyglm_means", self._p(Xp), ldx, self._p(Wt), Wt.shape[1], T, D, n, self._p(mu), n, self._stream())
        return mu

    # ------------------------------------------------------------------ (2) Polya-gamma
    def pg_draw(self, psi, n_valid, omega, seed, call_id, t_off, n_off, n_total):
        """omega ~ PG(1, psi).  Default: the branch-compacted two-pass kernels (same draws, element by element, as
        the one-pass kernel, which PYGLM_PG_VARIANT=1 selects for comparison)."""
        T = psi.shape[0]
        if os.environ.get("PYGLM_PG_VARIANT", "2") == "1":
            self._call("pyglm_pg_draw", self._p(psi), psi.shape[1], T, n_valid, self._p(omega), omega.shape[1],
                       seed, call_id, t_off, n_off, n_total, self._stream())
            return omega
        need = int(self.lib.pyglm_pg_draw_ws_bytes(T, n_valid))
        # one workspace per stream: two draws enqueued on different streams must not share their index lists
        cache = self.__dict__.setdefault("_pg_ws", {})
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = cache.get(key)
        if ws is None or ws.numel() < need:
            ws = cache[key] = torch.empty(need, dtype=torch.uint8, device=self.device)
        self._call("pyglm_pg_draw_ws", self._p(psi), psi.shape[1], T, n_valid, self._p(omega), omega.shape[1],
                   seed, call_id, t_off, n_off, n_total, self._p(ws), ws.numel(), self._stream(), launches=3)
        return omega

    def philox_uniforms(self, seed, call_id, elem0, n_elem, count):
        out = self.empty(n_elem, count)
        self._call("pyglm_philox_uniforms", seed, call_id, elem0, n_elem, count, self._p(out), self._stream())
        return out

    # ------------------------------------------------------------------ (3) weighted Gram, h
    def gram_tiles(self, D, mode):
        key = (D, mode)
        if key not in self._tiles:
            n = self.lib.pyglm_gram_tiles(D, mode, None, 0)
            buf = np.zeros((n, 2), dtype=np.int32)
            self.lib.pyglm_gram_tiles(D, mode, buf.ctypes.data_as(ctypes.c_void_p), n)
            self._tiles[key] = self.to_device(buf)
        return self._tiles[key]

    def weighted_gram(self, Xp, Om, D, n_valid, J=None, nslabs=None, tiles=None):
        """J[n, i, j] = sum_t Xp[t,i] Xp[t,j] Om[t,n] on the lower triangle (tile granularity).
        tiles: a subset of the rows of gram_tiles(D, 0) (device int32 (k, 2), contiguous) -- only those (i, j) tiles
        are computed, the rest of J is left as it was."""
        T, ldx = Xp.shape
        if tiles is None:
            tiles = self.gram_tiles(D, 0)
        if J is None:
            J = self.zeros(n_valid, ldx, ldx)
        if nslabs is None:
            nslabs = self.lib.pyglm_gram_slabs(tiles.shape[0], n_valid, T)
        ws = self.empty(nslabs * n_valid * ldx * ldx) if nslabs > 1 else None
        self._call("pyglm_weighted_gram", self._p(Xp), ldx, T, self._p(Om), Om.shape[1], n_valid, self._p(tiles),
                   tiles.shape[0], nslabs, self._p(J), ldx * ldx, ldx, 0, self._p(ws), self._stream(),
                   launches=2 if nslabs > 1 else 1)
        return J

    def xt_kappa(self, Xp, kappa, D, n_valid):
        """h[n, d] = sum_t Xp[t,d] kappa[t,n]: the bias row of the weighted Gram with weights kappa."""
        T, ldx = Xp.shape
        tiles = self.gram_tiles(D, 1)
        i_base = ((D - 1) // 8) * 8
        rows = self.zeros(n_valid, 8, ldx)
        nslabs = self.lib.pyglm_gram_slabs(tiles.shape[0], n_valid, T)
        ws = self.empty(nslabs * n_valid * 8 * ldx) if nslabs > 1 else None
        self._call("pyglm_weighted_gram", self._p(Xp), ldx, T, self._p(kappa), kappa.shape[1], n_valid,
                   self._p(tiles), tiles.shape[0], nslabs, self._p(rows), 8 * ldx, ldx, i_base, self._p(ws),
                   self._stream(), launches=2 if nslabs > 1 else 1)
        return rows[:, D - 1 - i_base, :].contiguous()

    # ------------------------------------------------------------------ (3') weighted Gram on tcgen05
    def gram_tc_geometry(self, D, n_valid, T, S):
        g = (ctypes.c_longlong * 8)()
        cabi.call("pyglm_gram_tc_geometry", D, n_valid, T, S, ctypes.cast(g, ctypes.c_void_p))
        keys = ("M", "Mpad", "Tpad", "Npad", "nt", "n_ntiles", "n_chunks", "blocks_per_chunk")
        return dict(zip(keys, [int(v) for v in g]))

    def column_max(self, A, ncols):
        """(cmax (ncols,) f64, has_negative bool) of the first ncols columns of a (T, ld) device matrix."""
        T, ld = A.shape
        cmax = self.empty(ncols)
        neg = self.empty(1, dtype=torch.int32)
        self._call("pyglm_column_max", self._p(A), ld, T, ncols, self._p(cmax), self._p(neg), self._stream())
        return cmax, bool(neg.item())

    def gram_tc_plan(self, Xp, D, n_valid, S=4, comm=None, t_off=0, stream=False, share=None):
        """Sweep-invariant operand of the tensor-core Gram (once per dataset) and its per-sweep buffers.
        stream=False: the digit planes of Z = X~_i X~_j stay resident in HBM (S * pairs * T bytes); stream=True: only
        the fixed-point design (4 * D * T bytes) is kept and the kernel builds the Z tiles in shared memory (S = 4).
        Both give the same integer sums.  Raises ValueError when the design has negative entries.  comm / t_off:
        time-sharded runs (Xp is the slab starting at global bin t_off; scales are all-reduced over `comm`).
        share: another plan of the SAME design (same Xp, S, mode) whose sweep-invariant operand is reused -- plans of
        different neuron groups of one data set then differ only in their per-sweep buffers (Os, omax, Jint)."""
        return TcGramPlan(self, Xp, D, n_valid, S, comm=comm, t_off=t_off, stream=stream, share=share)

    def gram_tc_stream_tiles(self, D):
        key = ("tc_stream_tiles", D)
        if key not in self._tiles:
            n = self.lib.pyglm_gram_tc_stream_tiles(D, None, 0)
            buf = np.zeros((n, 2), dtype=np.int32)
            self.lib.pyglm_gram_tc_stream_tiles(D, buf.ctypes.data_as(ctypes.c_void_p), n)
            self._tiles[key] = self.to_device(buf)
        return self._tiles[key]

    # ------------------------------------------------------------------ (6) forward simulation
    def generate(self, Wm, bias, basis, T, seed, call_id, want_uniforms=False, gauss_sd=-1.0):
        """Wm (N, N*B), bias (N), basis (L, B) device f64 -> (Xp (T, ldx) padded design, Y (T, N), U or None).
        gauss_sd >= 0: Gaussian observations with that standard deviation instead of Bernoulli spikes."""
        N, NB = Wm.shape
        L, B = basis.shape
        assert NB == N * B and bias.shape[0] == N and T > 0
        ldx = pad_ldx(NB + 1)
        Xp = self.zeros(T, ldx)
        Xp[:, NB] = 1.0
        Y = self.empty(T, N)
        U = self.empty(T, N) if want_uniforms else None
        self._call("pyglm_generate", self._p(Wm.contiguous()), self._p(bias.contiguous()), self._p(basis.contiguous()),
                   N, B, L, T, seed, call_id, ctypes.c_double(gauss_sd), self._p(Xp), ldx, self._p(Y), self._p(U),
                   self._stream())
        return Xp, Y, U

    # ------------------------------------------------------------------ (4) spike and slab
    def scan_randomness(self, N, B, n_loc, n_off, seed, call_id):
        D = N * B + 1
        perm = self.empty(n_loc, N, dtype=torch.int32)
        us = self.empty(n_loc, N)
        z = self.empty(n_loc, D)
        self._call("pyglm_scan_randomness", N, B, n_loc, n_off, seed, call_id, self._p(perm), self._p(us),
                   self._p(z), D, self._stream())
        return perm, us, z

    def spike_slab_update(self, N, B, J, h, prior, perm, us, z, do_scan, a, P_ws=None, want_logodds=False,
                          want_ml=False):
        """J (n_loc, ldj, ldj) likelihood Gram (lower triangle), h (n_loc, ldh).  prior: dict of device
        tensors J0w (n_loc,N,B,B), h0w (n_loc,N,B), J0b, h0b (n_loc,), cprior, logit_rho (n_loc,N).
        a (n_loc, N) uint8 is updated in place.  Returns (W, bias, logodds|None, ml|None)."""
        n_loc = a.shape[0]
        D = N * B + 1
        if P_ws is None:
            P_ws = self.empty(n_loc * D * D)
        W = self.empty(n_loc, N, B)
        bias = self.empty(n_loc)
        logodds = self.zeros(n_loc, N) if want_logodds else None
        ml = self.empty(n_loc) if want_ml else None
        status = self.zeros(n_loc, dtype=torch.int32)
        self._call("pyglm_spike_slab_update", N, B, n_loc, self._p(J), J.shape[1] * J.shape[2], J.shape[2],
                   self._p(h), h.shape[1], self._p(prior["J0w"]), self._p(prior["h0w"]), self._p(prior["J0b"]),
                   self._p(prior["h0b"]), self._p(prior["cprior"]), self._p(prior["logit_rho"]), self._p(perm),
                   self._p(us), self._p(z), z.shape[1], self._p(do_scan), self._p(a), self._p(W), self._p(bias),
                   self._p(P_ws), self._p(logodds), self._p(ml), self._p(status), self._stream())
        return W, bias, logodds, ml, status


def gram_tc_bytes(D, n_valid, T, S=4, stream=False):
    """HBM bytes the tensor-core Gram keeps resident for one dataset: digit planes of Z (resident mode) or the
    fixed-point design and the rounding addends (streaming mode), digit planes of omega, Jint."""
    M = D * (D + 1) // 2
    Mpad, Tpad = round_up(M, 128), round_up(T, 64)
    operand = (4 * round_up(D, 16) + 8) * Tpad if stream else S * Mpad * Tpad
    return operand + S * round_up(n_valid, 16) * Tpad * 2 + n_valid * Mpad * 8


class TcGramPlan(object):
    """Resident state of the tcgen05 Gram for one dataset: Zs (S, Mpad, Tpad) uint8 digit planes of the
    Khatri-Rao operand (built once) -- or, in streaming mode, xq (D, Tpad) uint32 fixed-point design + rw (Tpad)
    rounding addends, from which the kernel builds the same digit tiles in shared memory -- Os (S, Npad, Tpad) digit
    planes of omega and Jint (n, Mpad) int64 (per sweep).

    Time-sharded runs pass `comm` and the slab's global offset `t_off`: the fixed-point scales (column maxima of X
    and of omega) are all-reduced (max) over the ranks and the rounding dither is keyed by the global time bin, so
    the integer partial sums of the slabs add up -- exactly, in int64 -- to the Jint a single GPU would compute."""

    def __init__(self, K, Xp, D, n_valid, S=4, comm=None, t_off=0, stream=False, share=None):
        self.K, self.D, self.n, self.S = K, D, n_valid, S
        self.stream = bool(stream)
        if self.stream and S != 4:
            raise ValueError("the streaming tensor-core Gram builds 4-digit tiles (S=4), not S=%d" % S)
        self.comm = comm if (comm is not None and comm.world > 1) else None
        self.verified, self.max_rel_dev = False, None      # set by GibbsEngine._tc_verified
        self.uses = self.checks = 0                        # sweeps served / spot checks run (GibbsEngine._tc_spot_check)
        self.max_rel_dev_spot = None
        self.T, self.ldx = Xp.shape
        g = K.gram_tc_geometry(D, n_valid, self.T, S)
        self.geom = g
        self.neg = K.empty(1, dtype=torch.int32)
        if share is not None:
            assert (share.D, share.S, share.T, share.ldx, share.stream) == (D, S, self.T, self.ldx, self.stream)
            self.cmax, self.Zs, self.tiles = share.cmax, share.Zs, getattr(share, "tiles", None)
            self.xq, self.rw = getattr(share, "xq", None), getattr(share, "rw", None)
            self.neg.zero_()
            self.Os = torch.zeros(S, g["Npad"], g["Tpad"], dtype=torch.uint8, device=K.device)
            self.omax = K.empty(n_valid)
            self.Jint = K.empty(n_valid, g["Mpad"], dtype=torch.int64)
            return
        self.cmax = K.empty(D)
        K._call("pyglm_column_max", K._p(Xp), self.ldx, self.T, D, K._p(self.cmax), K._p(self.neg), K._stream())
        if self.comm is not None:
            self.comm.all_reduce_max(self.cmax)
            self.comm.all_reduce_max(self.neg)
        if bool(self.neg.item()):
            raise ValueError("tensor-core Gram needs a non-negative design matrix")
        if self.stream:
            self.Zs = None
            # tiled: xq[t // 32][column][t % 32], Dp = D rounded up to 16 rows per block        (uint32 bit patterns)
            self.xq = torch.zeros(g["Tpad"] // 32, round_up(D, 16), 32, dtype=torch.int32, device=K.device)
            self.rw = torch.empty(g["Tpad"], dtype=torch.int64, device=K.device)          # uint64 bit patterns
            self.tiles = K.gram_tc_stream_tiles(D)
            K._call("pyglm_gram_tc_quantize", K._p(Xp), self.ldx, self.T, int(t_off), D, K._p(self.cmax),
                    K._p(self.xq), K._p(self.rw), g["Tpad"], K._stream())
        else:
            self.Zs = torch.zeros(S, g["Mpad"], g["Tpad"], dtype=torch.uint8, device=K.device)
            K._call("pyglm_gram_tc_build_z_slab", K._p(Xp), self.ldx, self.T, int(t_off), D, K._p(self.cmax), S,
                    K._p(self.Zs), g["Mpad"], g["Tpad"], K._stream())
        self.Os = torch.zeros(S, g["Npad"], g["Tpad"], dtype=torch.uint8, device=K.device)
        self.omax = K.empty(n_valid)
        self.Jint = K.empty(n_valid, g["Mpad"], dtype=torch.int64)

    def slice_omega(self, Om):
        K, g = self.K, self.geom
        if self.comm is None:
            K._call("pyglm_gram_tc_slice_omega", K._p(Om), Om.shape[1], self.T, self.n, self.S, K._p(self.omax),
                    K._p(self.neg), K._p(self.Os), g["Npad"], g["Tpad"], int(self.stream), K._stream(), launches=2)
            return
        K._call("pyglm_column_max", K._p(Om), Om.shape[1], self.T, self.n, K._p(self.omax), K._p(self.neg),
                K._stream())
        self.comm.all_reduce_max(self.omax)
        K._call("pyglm_gram_tc_slice_digits", K._p(Om), Om.shape[1], self.T, self.n, self.S, K._p(self.omax),
                K._p(self.Os), g["Npad"], g["Tpad"], int(self.stream), K._stream())

    def mma(self, max_ctas=0):
        K, g = self.K, self.geom
        if self.stream:
            K._call("pyglm_gram_tc_mma_stream", K._p(self.xq), K._p(self.rw), K._p(self.Os), self.D, self.n, self.T,
                    self.S, K._p(self.tiles), self.tiles.shape[0], K._p(self.Jint), g["Mpad"], max_ctas, K._stream())
        else:
            K._call("pyglm_gram_tc_mma", K._p(self.Zs), K._p(self.Os), self.D, self.n, self.T, self.S,
                    K._p(self.Jint), g["Mpad"], max_ctas, K._stream())
        return self.Jint

    def mma_probe(self):
        """Issue-rate probe of the MMA schedule (no loads, no atomics): measures the int8 tensor-pipe peak."""
        K, g = self.K, self.geom
        assert not self.stream, "the issue-rate probe belongs to the resident kernel"
        K._call("pyglm_gram_tc_mma_probe", K._p(self.Zs), K._p(self.Os), self.D, self.n, self.T, self.S,
                K._p(self.Jint), g["Mpad"], K._stream())

    def finalize(self, J, Jint=None, omax=None):
        """J[n] (FP64, lower triangle) from the integer sums; Jint / omax default to the plan's own (all n neurons),
        or are the rows a reduce-scatter handed this rank together with the matching slice of omax."""
        K, g = self.K, self.geom
        Jint = self.Jint if Jint is None else Jint
        omax = self.omax if omax is None else omax
        n = Jint.shape[0]
        assert Jint.shape[1] == g["Mpad"] and Jint.is_contiguous() and omax.shape[0] >= n and J.shape[0] >= n
        K._call("pyglm_gram_tc_finalize", K._p(Jint), g["Mpad"], K._p(self.cmax), K._p(omax), self.D,
                n, self.S, K._p(J), J.shape[1] * J.shape[2], J.shape[2], K._stream())
        return J

    def finalize_peers(self, J, hdl, world, row_off, n, omax):
        """J[0:n] = scaled sum over the ranks' peer-mapped Jint rows [row_off, row_off + n): the time-sharded
        reduce-scatter and the finalize pass as one kernel (hdl: symmetric-memory handle of self.Jint)."""
        K, g = self.K, self.geom
        assert J.shape[0] >= n and omax.shape[0] >= n and omax.is_contiguous()
        K._call("pyglm_gram_tc_finalize_peers", hdl.buffer_ptrs_dev, int(world), int(row_off), g["Mpad"],
                K._p(self.cmax), K._p(omax), self.D, int(n), self.S, K._p(J), J.shape[1] * J.shape[2], J.shape[2],
                K._stream())
        return J

    def gram(self, Om, J=None):
        """J[n, i, j] (i >= j) = sum_t Xp[t,i] Xp[t,j] Om[t,n] for Om > 0."""
        if J is None:
            J = self.K.zeros(self.n, self.ldx, self.ldx)
        self.slice_omega(Om)
        self.mma()
        return self.finalize(J)
