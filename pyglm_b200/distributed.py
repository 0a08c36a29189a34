"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in the
CPU tests) for the only two exchange steps the Gibbs sweep has:

  neuron-sharded : postsynaptic neurons are split into contiguous blocks, X is replicated; the regressions are
                   conditionally independent given the data (models.py:169-171), so the ONLY collective is one
                   all-gather of the new (a, W, b) rows per sweep.
  time-sharded   : the time axis is split; every rank forms partial Gram matrices for all neurons over its slab;
                   ONE reduce-scatter over the neuron axis hands rank r the complete J for its neuron block; then
                   the same all-gather of (a, W, b).

The partition and the collectives are written against torch tensors on any device so the host-side logic is
covered by world_size-2 gloo tests on CPU.
"""
import torch
import torch.distributed as dist


def block_partition(N, world, rank):
    """Contiguous neuron block of `rank`: [lo, hi) with block size ceil(N / world)."""
    n_max = (N + world - 1) // world
    lo = min(N, rank * n_max)
    hi = min(N, lo + n_max)
    return lo, hi, n_max


def time_partition(T, world, rank):
    """Contiguous time slab of `rank`: [lo, hi)."""
    per = (T + world - 1) // world
    lo = min(T, rank * per)
    return lo, min(T, lo + per)


_HOST_GROUPS = {}


class Comm(object):
    """Thin wrapper over a torch.distributed process group (None -> single process).

    Host objects (the network state rank 0 draws each sweep) travel over a gloo side group: an NCCL object
    broadcast is ordered behind whatever the compute stream holds -- with pipelined sweeps that is the whole next
    psi / PG / Gram -- and would serialise the host step with it.  Constructing a Comm is collective."""

    def __init__(self, group=None):
        self.collective_calls = 0        # torch.distributed collectives issued through this object (bench.py reports
                                         # how many a steady-state sweep makes: 0 with the peer-memory exchange)
        self.enabled = dist.is_available() and dist.is_initialized()
        self.group = group
        self.world = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0
        self.host_group = group
        if self.enabled and self.world > 1 and dist.get_backend(group) != "gloo":
            ranks = tuple(dist.get_process_group_ranks(group if group is not None else dist.group.WORLD))
            if ranks not in _HOST_GROUPS:
                _HOST_GROUPS[ranks] = dist.new_group(ranks=list(ranks), backend="gloo")
            self.host_group = _HOST_GROUPS[ranks]

    def all_gather_rows(self, local):
        """local (n_max, ...) on every rank -> (world * n_max, ...), rank-major."""
        if self.world == 1:
            return local
        local = local.contiguous()
        out = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        self.collective_calls += 1
        dist.all_gather_into_tensor(out, local, group=self.group)
        return out

    def reduce_scatter_rows(self, full):
        """full (world * n_max, ...) partial sums on every rank -> this rank's (n_max, ...) block of the total."""
        if self.world == 1:
            return full
        full = full.contiguous()
        n_max = full.shape[0] // self.world
        out = torch.empty((n_max,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
        if full.device.type == "cpu":
            # gloo has no reduce_scatter: all-reduce and slice (CPU tests only)
            tmp = full.clone()
            self.collective_calls += 1
            dist.all_reduce(tmp, group=self.group)
            out.copy_(tmp[self.rank * n_max:(self.rank + 1) * n_max])
        else:
            self.collective_calls += 1
            dist.reduce_scatter_tensor(out, full, group=self.group)
        return out

    def all_reduce_sum(self, t):
        if self.world > 1:
            self.collective_calls += 1
            dist.all_reduce(t, group=self.group)
        return t

    def all_reduce_max(self, t):
        if self.world > 1:
            peer = getattr(self, "peer", None)
            if peer is not None and peer.handles_small(t):
                return peer.all_reduce_max_small(t)
            self.collective_calls += 1
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def all_reduce_min(self, t):
        if self.world > 1:
            self.collective_calls += 1
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return t

    def broadcast_object(self, obj, src=0):
        if self.world == 1:
            return obj
        box = [obj if self.rank == src else None]
        dist.broadcast_object_list(box, src=src, group=self.host_group)
        return box[0]

    def barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)


class PeerExchange(object):
    """The exchange steps of the sweep over PEER-MAPPED memory, driven by this package's own kernels instead of NCCL
    calls.  torch's symmetric-memory allocator supplies what a kernel cannot do for
    itself: buffers of identical layout on every rank, the peers' base pointers (a device array, `buffer_ptrs_dev`) and a
    device-side barrier on the current stream (signal pads; no host synchronisation).  On top of it:

      state rows   csrc/peer.cu pyglm_peer_push: every rank stores its block of the new [a | W | b | status] rows into
                   ALL ranks' buffers over NVLink (the all-gather of models.py:169-171's results), then one barrier;
      Gram         csrc/gram_tc.cu pyglm_gram_tc_finalize_peers: the owner of a neuron block READS the other ranks'
                   int64 partial sums of its rows from their HBM, adds them exactly and scales to FP64 in one pass
                   (reduce-scatter fused into the finalize kernel), between two barriers;
      omega maxima the per-sweep all-reduce(max) of N doubles: push into slot `rank` of every peer, barrier, local max.

    Buffers written by peers are double-buffered by call parity, so one barrier per exchange suffices.  Constructing
    buffers is collective (rendezvous): every rank must allocate in the same order."""

    TIMEOUT_MS = 60000
    SMALL_CAP = 4096            # doubles per rank in the small all-reduce scratch

    def __init__(self, comm, kernels):
        from torch.distributed import _symmetric_memory as symm
        self.symm, self.comm, self.K = symm, comm, kernels
        self.group = comm.group if comm.group is not None else dist.group.WORLD
        self._small = None
        self._small_par = 0

    @staticmethod
    def usable(comm, device):
        import os
        if not comm.enabled:
            return False
        if comm.world <= 1 or torch.device(device).type != "cuda" or os.environ.get("PYGLM_PEER_EXCHANGE", "1") == "0":
            return False
        try:
            from torch.distributed import _symmetric_memory  # noqa: F401
        except Exception:
            return False
        return dist.get_backend(comm.group) == "nccl"

    def alloc(self, shape, dtype, zero=False):
        """(tensor, handle): a symmetric buffer of this shape on every rank.  Collective."""
        t = self.symm.empty(*shape, dtype=dtype, device=self.K.device)
        if zero:
            t.zero_()
        hdl = self.symm.rendezvous(t, self.group)
        return t, hdl

    def barrier(self, hdl, channel=0):
        hdl.barrier(channel=channel, timeout_ms=self.TIMEOUT_MS)

    def push(self, src, hdl, dst_off_bytes):
        """src (contiguous, a multiple of 16 bytes) -> byte offset dst_off_bytes of every rank's buffer."""
        self.K._call("pyglm_peer_push", self.K._p(src), src.numel() * src.element_size(), hdl.buffer_ptrs_dev,
                     self.comm.world, int(dst_off_bytes), self.K._stream())

    def handles_small(self, t):
        return t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() <= self.SMALL_CAP

    def all_reduce_max_small(self, t):
        W, cap = self.comm.world, self.SMALL_CAP
        if self._small is None:
            self._small = [self.alloc((W, cap), torch.float64, zero=True) for _ in range(2)]
        buf, hdl = self._small[self._small_par]
        self._small_par ^= 1
        n = t.numel()
        npad = (n + 1) & ~1
        src = t.reshape(-1)
        if npad != n:
            src = torch.cat([src, src[-1:]])
        elif src.data_ptr() % 16:
            src = src.clone()
        self.push(src.contiguous(), hdl, self.comm.rank * cap * 8)
        self.barrier(hdl)
        t.copy_(buf[:, :n].max(0).values.reshape(t.shape))
        return t


class StateExchange(object):
    """All-gather of the state rows through PeerExchange: two symmetric (world * n_max, width) buffers used alternately."""

    def __init__(self, peer, n_max, width):
        self.peer, self.n_max, self.width = peer, n_max, width
        W = peer.comm.world
        self.bufs = [peer.alloc((W * n_max, width), torch.float64, zero=True) for _ in range(2)]
        self.par = 0

    def all_gather_rows(self, local):
        buf, hdl = self.bufs[self.par]
        self.par ^= 1
        assert local.shape == (self.n_max, self.width) and local.is_contiguous()
        self.peer.push(local, hdl, self.peer.comm.rank * self.n_max * self.width * 8)
        self.peer.barrier(hdl)
        return buf
