"""Node sharding of the collocation path over several GPUs (one process per
GPU, ``torch.distributed`` for the plumbing).

Constraint node ``i`` only reads trajectory columns ``i`` and ``i + 1``
(opty/direct_collocation.py:2145, 2153-2155), so the ``N - 1`` nodes split
into contiguous ranges with a one-column halo and no data-path collective
(SURVEY.md §8e).  A rank's Jacobian values are one contiguous slice of the
node-major array (opty/direct_collocation.py:2681-2684); its residuals are
``M`` strided segments of the eom-major array (opty/direct_collocation.py:
2446), which is why gathered residual blocks have to be re-tiled.

The gather is only needed when a consumer wants the full vectors in one
place: with IPOPT on the host of rank 0 use :func:`gather_to_host`; for
device-resident consumers :meth:`ShardedCollocator.allgather_device` runs an
NCCL ``all_gather_into_tensor`` on the shards' device buffers.
"""

import numpy as np


def node_shard(num_collocation_nodes, rank, world_size):
    """Contiguous, balanced range ``(lo, hi)`` of the ``N - 1`` constraint
    nodes owned by ``rank``; the first ``(N-1) % world_size`` ranks get one
    extra node."""
    nn = num_collocation_nodes - 1
    if not 0 <= rank < world_size:
        raise ValueError('rank must be in [0, world_size)')
    if world_size > nn:
        raise ValueError('more ranks than constraint nodes')
    base, extra = divmod(nn, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def bind_to_gpu_numa_node(device):
    """Pins the calling process to the CPU cores that are local to GPU
    ``device`` (NVML's CPU affinity of the device) before the pinned host
    buffers are allocated, so that the Jacobian lands in host memory on the
    GPU's own socket.  With one process per GPU the device->host copies of the
    ranks then do not all cross the inter-socket link.  Returns the CPU set or
    ``None`` when NVML is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words)
                for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:       # no NVML / not permitted: keep the default
        return None
    return None


def all_shards(num_collocation_nodes, world_size):
    return [node_shard(num_collocation_nodes, r, world_size)
            for r in range(world_size)]


def assemble_constraints(blocks, shards, num_eom):
    """Re-tiles per-shard eom-major residual blocks ``(M * nn_g,)`` into the
    eom-major vector of the whole problem ``(M * (N-1),)``."""
    total = shards[-1][1]
    out = np.empty((num_eom, total))
    for blk, (lo, hi) in zip(blocks, shards):
        out[:, lo:hi] = np.asarray(blk).reshape(num_eom, hi - lo)
    return out.ravel()


def assemble_jacobian(blocks):
    """Node-major Jacobian blocks are contiguous slices: concatenate."""
    return np.concatenate([np.asarray(b) for b in blocks])


def _all_gather_ragged(local, lengths, dist, group=None):
    """all_gather of 1-D tensors of different lengths (padded to the longest,
    ``all_gather_into_tensor`` needs equal sizes); returns a list of
    tensors trimmed to ``lengths``."""
    import torch
    longest = max(lengths)
    padded = local
    if local.numel() < longest:
        padded = torch.zeros(longest, dtype=local.dtype, device=local.device)
        padded[:local.numel()] = local
    out = torch.empty(len(lengths) * longest, dtype=local.dtype,
                      device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    return [out[r * longest:r * longest + n] for r, n in enumerate(lengths)]


def gather_vectors(local_con, local_jac, num_collocation_nodes, num_eom,
                   dist, group=None):
    """Collective: every rank contributes its shard's residuals and Jacobian
    values (1-D torch tensors on the backend's device: CPU for gloo, CUDA for
    nccl) and receives the full eom-major residual vector and the full
    node-major Jacobian value vector (EOM parts only).

    With equal shards the node-major Jacobian blocks land directly in the
    final vector (``all_gather_into_tensor`` into the result, no staging
    copy) and the eom-major residuals take one gather plus ONE permute
    ``(G, M, nn) -> (M, G*nn)``; ragged shards (``N - 1`` not divisible by
    the number of ranks) go through a padded gather and one concatenation
    (NCCL's own gather of unequal blocks into views of the final vector --
    one broadcast per rank -- measured slower: 17.2 against 13.5 ms for the
    8.4 GB of the 50-link chain on two GPUs)."""
    import torch
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    shards = all_shards(num_collocation_nodes, world)
    sizes = [hi - lo for lo, hi in shards]
    nn = sizes[rank]
    K = local_jac.numel() // nn
    total = shards[-1][1]
    if len(set(sizes)) == 1:
        jac = torch.empty(total * K, dtype=local_jac.dtype,
                          device=local_jac.device)
        dist.all_gather_into_tensor(jac, local_jac.contiguous(), group=group)
        staged = torch.empty(world * num_eom * nn, dtype=local_con.dtype,
                             device=local_con.device)
        dist.all_gather_into_tensor(staged, local_con.contiguous(),
                                    group=group)
        con = staged.view(world, num_eom, nn).permute(1, 0, 2).reshape(-1)
        return con, jac
    con_blocks = _all_gather_ragged(
        local_con, [num_eom * n for n in sizes], dist, group)
    jac_blocks = _all_gather_ragged(
        local_jac, [K * n for n in sizes], dist, group)
    con = torch.empty((num_eom, total), dtype=local_con.dtype,
                      device=local_con.device)
    for blk, (lo, hi) in zip(con_blocks, shards):
        con[:, lo:hi] = blk.view(num_eom, hi - lo)
    return con.reshape(-1), torch.cat(jac_blocks)


class _CudaArray(object):
    """Exposes a raw device pointer through ``__cuda_array_interface__`` so
    that ``torch.as_tensor`` can wrap it without copying."""

    def __init__(self, ptr, count, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {
            'shape': (count,), 'typestr': '<f8', 'data': (ptr, False),
            'version': 2, 'strides': None}


class ShardedCollocator(object):
    """One rank's part of a node-sharded collocation problem.

    Builds an ``opty_b200.ConstraintCollocator`` restricted to this rank's
    node range on this rank's GPU and offers the gathers.  Instance
    constraints are evaluated by every rank from ``free`` on the host (they
    are ``o <~ 100`` scalar expressions) and appended after the gathered EOM
    parts, as in opty/direct_collocation.py:2985-2991.
    """

    def __init__(self, *args, rank=None, world_size=None, device=None,
                 **kwargs):
        import torch.distributed as dist
        from .direct_collocation import ConstraintCollocator
        self.dist = dist
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        if world_size is None:
            world_size = dist.get_world_size() if dist.is_initialized() else 1
        self.rank, self.world_size = rank, world_size
        N = args[2] if len(args) > 2 else kwargs['num_collocation_nodes']
        self.shards = all_shards(N, world_size)
        if world_size > 1 and device is not None:
            bind_to_gpu_numa_node(device)
        self.collocator = ConstraintCollocator(
            *args, node_range=self.shards[rank], device=device, **kwargs)
        self._con = self.collocator.generate_constraint_function()
        self._jac = self.collocator.generate_jacobian_function()

    @property
    def node_range(self):
        return self.shards[self.rank]

    def constraints_local(self, free):
        return self._con(free)

    def jacobian_local(self, free):
        return self._jac(free)

    def jacobian_indices_local(self):
        return self.collocator.jacobian_indices()

    def allgather_device(self, free=None):
        """Evaluates this rank's shard on its GPU (uploading ``free`` first if
        given) and all-gathers residual and Jacobian blocks over NCCL straight
        from the shard's device buffers.  Returns CUDA tensors
        ``(con (M*(N-1),), jac ((N-1)*M*P,))`` holding the full vectors."""
        import torch
        col = self.collocator
        handle = col._evaluator.handle
        if free is not None:
            handle.upload_free(np.ascontiguousarray(free, dtype=np.float64))
        handle.eval_device(sync=True)
        bufs = handle.device_buffers()
        dev = torch.device('cuda', col.device)
        con = torch.as_tensor(_CudaArray(bufs['con'], handle.con_len, handle),
                              device=dev)
        jac = torch.as_tensor(_CudaArray(bufs['jac'], handle.jac_len, handle),
                              device=dev)
        return gather_vectors(con, jac, col.num_collocation_nodes,
                              col.num_eom, self.dist)

    def gather_to_host(self, free):
        """Collective through the process group's default backend with host
        staging: returns NumPy ``(con, jac)`` of the whole problem incl. the
        instance-constraint parts."""
        import torch
        col = self.collocator
        lo, hi = self.node_range
        M = col.num_eom
        K = M * col._evaluator.program.P
        con = torch.from_numpy(np.array(self._con(free))[:M * (hi - lo)])
        jac = torch.from_numpy(np.array(self._jac(free))[:K * (hi - lo)])
        device = None
        if self.dist.get_backend() == 'nccl':
            device = torch.device('cuda', col.device)
            con, jac = con.to(device), jac.to(device)
        full_con, full_jac = gather_vectors(
            con, jac, col.num_collocation_nodes, M, self.dist)
        full_con, full_jac = full_con.cpu().numpy(), full_jac.cpu().numpy()
        if col.instance_constraints is not None:
            full_con = np.hstack((full_con,
                                  col.eval_instance_constraints(free)))
            full_jac = np.hstack((
                full_jac,
                col.eval_instance_constraints_jacobian_values(free)))
        return full_con, full_jac

    def close(self):
        self.collocator.close()
