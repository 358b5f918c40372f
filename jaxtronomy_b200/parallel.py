"""Multi-GPU: the particle batch is the only axis that shards.

One process per GPU (torch.distributed, backend "nccl" on GPUs, "gloo" in the CPU tests).  Parameter sets are
independent, so each rank evaluates a contiguous slice of the [B, P] parameter matrix with a full replica of the
plan and the only data-path exchange is ONE all-gather of the per-particle log-likelihoods (8 bytes per particle;
64 KiB per rank for 65,536 particles on 8 GPUs -- latency-bound, over NVLink/NVSwitch).  There is no reduction and
no halo: the sampler picks the best particle on the host after the gather (Sampling/Samplers/pso.py:51-66).
"""

import numpy as np


def shard_bounds(n, world_size, rank):
    """Contiguous, balanced slice [lo, hi) of n items for `rank` (the first n % world ranks get one extra)."""
    base, extra = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n, world_size):
    return [shard_bounds(n, world_size, r)[1] - shard_bounds(n, world_size, r)[0] for r in range(world_size)]


def all_gather_logl(full, world_size, rank, n):
    """In-place all-gather: on entry `full[lo:hi]` holds this rank's values, on exit `full` holds all of them.
    Works for CUDA tensors (NCCL) and CPU tensors (gloo)."""
    if world_size == 1:
        return full
    import torch.distributed as dist

    sizes = shard_sizes(n, world_size)
    if len(set(sizes)) == 1:
        lo, hi = shard_bounds(n, world_size, rank)
        dist.all_gather_into_tensor(full, full[lo:hi].clone())
    else:
        # uneven shards: pad every contribution to the largest shard, gather, then unpack
        import torch

        m = max(sizes)
        lo, hi = shard_bounds(n, world_size, rank)
        mine = torch.zeros(m, dtype=full.dtype, device=full.device)
        mine[: hi - lo] = full[lo:hi]
        gathered = torch.empty(world_size * m, dtype=full.dtype, device=full.device)
        dist.all_gather_into_tensor(gathered, mine)
        for r in range(world_size):
            a, b = shard_bounds(n, world_size, r)
            full[a:b] = gathered[r * m: r * m + (b - a)]
    return full


def sharded_logl(logl_fn, params, world_size, rank, out=None):
    """Evaluate ``logl_fn(params_slice) -> [n_local]`` on this rank's slice of ``params`` ([B, P] tensor present on
    every rank) and all-gather the result -> [B] on every rank."""
    import torch

    B = params.shape[0]
    lo, hi = shard_bounds(B, world_size, rank)
    if out is None:
        out = torch.empty(B, dtype=torch.float64, device=params.device)
    local = logl_fn(params[lo:hi])
    out[lo:hi] = local
    return all_gather_logl(out, world_size, rank, B)


def sharded_logl_grad(logl_grad_fn, params, world_size, rank):
    """Gradient counterpart of sharded_logl: ``logl_grad_fn(params_slice) -> ([n_local], [n_local, P])`` on this rank's
    slice, then ONE all-gather of the packed ``[logL | grad]`` rows ((1 + P) x 8 bytes per set) -> ([B], [B, P]) on
    every rank.  The chains of a gradient-descent run are independent, exactly like the particles of a swarm."""
    import torch

    B, P = params.shape
    lo, hi = shard_bounds(B, world_size, rank)
    logl, grad = logl_grad_fn(params[lo:hi])
    if world_size == 1:
        return logl, grad
    packed = torch.empty((B, P + 1), dtype=torch.float64, device=params.device)
    packed[lo:hi, 0] = logl
    packed[lo:hi, 1:] = grad
    import torch.distributed as dist

    sizes = shard_sizes(B, world_size)
    m = max(sizes)
    mine = torch.zeros((m, P + 1), dtype=torch.float64, device=params.device)
    mine[: hi - lo] = packed[lo:hi]
    gathered = torch.empty((world_size * m, P + 1), dtype=torch.float64, device=params.device)
    dist.all_gather_into_tensor(gathered, mine)
    for r in range(world_size):
        a, b = shard_bounds(B, world_size, r)
        packed[a:b] = gathered[r * m: r * m + (b - a)]
    return packed[:, 0].contiguous(), packed[:, 1:].contiguous()
