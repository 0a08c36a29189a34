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
            dist.all_reduce(tmp, group=self.group)
            out.copy_(tmp[self.rank * n_max:(self.rank + 1) * n_max])
        else:
            dist.reduce_scatter_tensor(out, full, group=self.group)
        return out

    def all_reduce_sum(self, t):
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return t

    def all_reduce_max(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def all_reduce_min(self, t):
        if self.world > 1:
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
