"""Shard-by-key search across the GPUs of one box: the one exchange step of the path.

The reference shards an index the same way on the CPU (`Indexes`, /root/reference/python/lib.cpp:74-107):
every query is searched in every shard and the per-shard results are merged by distance
(`search_typed(dense_indexes_py_t&)`, python/lib.cpp:321-402 → `merge_into`, index.hpp:2650-2670). There the
order in which shards reach the per-query lock decides ties; here the merge is deterministic:
ascending distance, then shard rank, then position within the shard.

One process per GPU (torch.distributed, NCCL on GPUs, gloo in the CPU tests). Each rank holds one complete
sub-index; the query batch is replicated; `merge_topk` is the only collective: one all-gather of the
`[nq, k]` keys and distances plus `[nq]` counts, then a stable sort of `world * k` candidates per query.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

SNAN_BITS = 0x7FA00000  # padding distance written by search_result_t::dump_to (index.hpp:2715-2720)


def shard_of(key: int, world: int) -> int:
    return int(key) % world


def merge_topk(keys: torch.Tensor, distances: torch.Tensor, counts: torch.Tensor, k: int,
               group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """All-gather per-shard results and merge them to the global top-k on every rank.

    keys: int64 [nq, k] (uint64 keys viewed as int64), distances: float32 [nq, k], counts: int32/int64 [nq].
    Rows are valid up to `counts`; the rest is padding (key 0, NaN). Returns the same triple, merged.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    nq = keys.shape[0]
    counts = counts.to(torch.int64)
    if world == 1:
        return keys, distances, counts
    gk = [torch.empty_like(keys) for _ in range(world)]
    gd = [torch.empty_like(distances) for _ in range(world)]
    gc = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(gk, keys.contiguous(), group=group)
    dist.all_gather(gd, distances.contiguous(), group=group)
    dist.all_gather(gc, counts.contiguous(), group=group)
    return merge_gathered(gk, gd, gc, k)


def merge_gathered(gk, gd, gc, k: int):
    """The local half of `merge_topk`: shard-major concatenation + stable sort = (distance, shard, rank)."""
    nq = gk[0].shape[0]
    col = torch.arange(gk[0].shape[1], device=gk[0].device)[None, :]
    masked = []
    for d, c in zip(gd, gc):
        valid = col < c[:, None]
        masked.append(torch.where(valid, d, torch.full_like(d, float("inf"))))
    cat_d = torch.cat(masked, dim=1)
    cat_k = torch.cat(gk, dim=1)
    order = torch.sort(cat_d, dim=1, stable=True).indices[:, :k]
    out_d = torch.gather(cat_d, 1, order)
    out_k = torch.gather(cat_k, 1, order)
    total = torch.stack(gc, 0).sum(0).clamp(max=k)
    pad = torch.arange(k, device=out_d.device)[None, :] >= total[:, None]
    nan = torch.tensor(SNAN_BITS, dtype=torch.int32, device=out_d.device).view(torch.float32)
    out_d = torch.where(pad, nan.expand_as(out_d), out_d)
    out_k = torch.where(pad, torch.zeros_like(out_k), out_k)
    return out_k, out_d, total
