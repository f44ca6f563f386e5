"""Multi-GPU: shard terminal samples by index, one process per GPU, no collective during the
solve, ONE all-gather of the endpoint records afterwards.

The reference has no distributed code (single-process jax.vmap); this is the north-star extension:
"Trajectory batches shard across the 8 GPUs by terminal-sample index, with only an NCCL all-gather
of endpoint (x, lambda, v) datasets for the value-function fit."  torch.distributed is the plumbing
(backend nccl on GPUs, gloo in the CPU tests of this logic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_size(n: int, world: int) -> int:
    """Every rank owns the same number of slots (the last ones may be padding)."""
    return (n + world - 1) // world


def shard_indices(n: int, rank: int, world: int, interleaved: bool = False) -> torch.Tensor:
    """Global indices of rank's shard, padded with -1 up to shard_size(n, world).
    contiguous: rank r owns [r*m, (r+1)*m).  interleaved: rank r owns r, r+W, r+2W, ... -- use it
    when the step count correlates with the sample index (e.g. the orbits theta sweep)."""
    m = shard_size(n, world)
    j = torch.arange(m, dtype=torch.int64)
    idx = j * world + rank if interleaved else rank * m + j
    return torch.where(idx < n, idx, torch.full_like(idx, -1))


def take_shard(y: torch.Tensor, n: int, rank: int, world: int, interleaved: bool = False) -> torch.Tensor:
    """y [rows, n] (struct-of-arrays) -> this rank's [rows, m] shard; padding columns are NaN, which
    the integrator retires immediately (result 4)."""
    idx = shard_indices(n, rank, world, interleaved).to(y.device)
    out = y[:, idx.clamp(min=0)].clone()
    if (idx < 0).any():
        if out.is_floating_point():
            out[:, idx < 0] = float("nan")
        else:
            out[:, idx < 0] = -1
    return out.contiguous()


def allgather_endpoints(local: torch.Tensor, n: int, interleaved: bool = False, group=None) -> torch.Tensor:
    """All-gather of per-rank records local [rows, m] into the global [rows, n] array, in the
    original sample order, on every rank.  One collective (all_gather_into_tensor)."""
    world = dist.get_world_size(group)
    rows, m = local.shape
    assert m == shard_size(n, world), "local shard has the wrong size"
    flat = torch.empty((world * rows, m), dtype=local.dtype, device=local.device)  # rank-major concat
    dist.all_gather_into_tensor(flat, local.contiguous(), group=group)
    gathered = flat.view(world, rows, m)
    if interleaved:
        full = gathered.permute(1, 2, 0).reshape(rows, m * world)  # column j*W + r
    else:
        full = gathered.permute(1, 0, 2).reshape(rows, world * m)  # column r*m + j
    return full[:, :n].contiguous()
