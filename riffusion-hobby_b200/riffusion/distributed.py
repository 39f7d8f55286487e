"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL on GPUs, gloo in the CPU tests).

Requests are independent (the reference serves them one at a time), so the data path has no collective: each
rank owns a contiguous block of the request list and a full replica of the frozen weights, which rank 0
broadcasts once at start-up.
"""
from __future__ import annotations

import typing as T

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous block partition of `n_items` requests: the first n % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def broadcast_state_dict(spec: T.Sequence[T.Tuple[str, T.Tuple[int, ...]]],
                         state_dict: T.Optional[T.Mapping[str, torch.Tensor]], src: int = 0,
                         device: T.Union[str, torch.device] = "cpu", dtype: torch.dtype = torch.float16,
                         group=None) -> T.Dict[str, torch.Tensor]:
    """Broadcast the weights named by `spec` from rank `src` as ONE flat buffer (a single large collective instead
    of ~700 small ones) and return views into it.  Ranks other than `src` pass state_dict=None."""
    total = sum(int(torch.Size(shape).numel()) for _, shape in spec)
    flat = torch.empty(total, dtype=dtype, device=device)
    if dist.get_rank(group) == src:
        assert state_dict is not None
        off = 0
        for name, shape in spec:
            n = int(torch.Size(shape).numel())
            flat[off: off + n].copy_(state_dict[name].reshape(-1))
            off += n
    if dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
    out, off = {}, 0
    for name, shape in spec:
        n = int(torch.Size(shape).numel())
        out[name] = flat[off: off + n].view(shape)
        off += n
    return out


def gather_results(local: torch.Tensor, n_items: int, group=None) -> T.Optional[torch.Tensor]:
    """Collect per-rank result blocks (same trailing shape, block sizes from `shard_range`) on rank 0."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [len(shard_range(n_items, r, world)) for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    gathered = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([g[:s] for g, s in zip(gathered, sizes)])
