"""Multi-GPU plumbing for the sampling path: edit requests are independent, so request *i* goes to
rank ``i mod world`` and the only exchange is an optional final gather of the finished latents / images."""
from __future__ import annotations

from typing import List, Tuple

import torch


def shard_requests(n_requests: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_requests, world))


def gather_latents(my_ids: List[int], my_latents: torch.Tensor, n_requests: int) -> Tuple[List[int], torch.Tensor]:
    """all_gather (NCCL on GPUs, gloo on CPU) of per-rank results back into request order."""
    import torch.distributed as dist
    world = dist.get_world_size()
    per = (n_requests + world - 1) // world
    pad = torch.zeros((per,) + tuple(my_latents.shape[1:]), dtype=my_latents.dtype, device=my_latents.device)
    pad[: len(my_ids)] = my_latents
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    out = torch.zeros((n_requests,) + tuple(my_latents.shape[1:]), dtype=my_latents.dtype, device=my_latents.device)
    ids = []
    for r in range(world):
        rid = shard_requests(n_requests, r, world)
        out[rid] = bufs[r][: len(rid)]
        ids += rid
    return sorted(ids), out
