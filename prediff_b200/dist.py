"""Ensemble / batch sharding across the GPUs of one box.

Every ensemble member is an independent chain (no cross-sample op in the UNet, VAE or sampler), so the path shards
with NO data-path collective: the global z_T is drawn from one seed and sliced contiguously by rank, weights are
replicated, and exactly one all-gather of the decoded frames happens at the end. One process per GPU
(torch.distributed, NCCL on GPUs; the same code runs under gloo on CPU for the host-logic tests)."""
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split rows [lo, hi) of rank `rank`; the first (global_rows % world) ranks get one extra row."""
    base, rem = divmod(global_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def sample_ensemble(run_local: Callable[..., torch.Tensor], z_T: torch.Tensor, cond: torch.Tensor,
                    gathered: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Runs `run_local(z_T[lo:hi], cond[lo:hi]) -> frames` on this rank's rows of the GLOBAL inputs and all-gathers
    the decoded frames so every rank returns the full (G, ...) tensor. Results are invariant to the world size
    because rows never interact. Ragged splits (G % world != 0) are padded to the largest shard for the gather.

    In-place variant: with a persistent `gathered` buffer of shape (G, ...) and G % world == 0,
    `run_local(z, c, out=gathered[lo:hi])` is asked to write its frames straight into this rank's slice (the decoder's
    last kernel does, `LatentDiffusion.sample(out=...)`), and the all-gather runs in place on that buffer - no
    allocation and no staging copy per call."""
    rank, world = world_info()
    G = z_T.shape[0]
    lo, hi = shard_bounds(G, rank, world)
    if gathered is not None and G % world == 0 and hi > lo:
        assert gathered.shape[0] == G and gathered.is_contiguous()
        mine = gathered[lo:hi]
        res = run_local(z_T[lo:hi], cond[lo:hi], out=mine)
        if res.data_ptr() != mine.data_ptr():   # run_local ignored `out`
            mine.copy_(res)
        if world > 1:
            dist.all_gather_into_tensor(gathered, mine)
        return gathered
    local = run_local(z_T[lo:hi], cond[lo:hi]) if hi > lo else None
    if world == 1:
        return local
    max_rows = shard_bounds(G, 0, world)[1]
    if local is None:
        raise RuntimeError("sample_ensemble: fewer ensemble members than ranks")
    tail = local.shape[1:]
    send = local
    if local.shape[0] < max_rows:
        send = torch.cat([local, local.new_zeros((max_rows - local.shape[0],) + tuple(tail))])
    gathered = local.new_empty((world * max_rows,) + tuple(tail))
    dist.all_gather_into_tensor(gathered, send.contiguous())
    if G % world == 0:
        return gathered
    parts = []
    for r in range(world):
        l, h = shard_bounds(G, r, world)
        parts.append(gathered[r * max_rows:r * max_rows + (h - l)])
    return torch.cat(parts)
