"""Shard-by-key search across the GPUs of one box: host-side glue for the sharded search of the C ABI.

The reference shards an index the same way on the CPU (`Indexes`, /root/reference/python/lib.cpp:74-107):
every query is searched in every shard and the per-shard results are merged by distance
(`search_typed(dense_indexes_py_t&)`, python/lib.cpp:321-402 -> `merge_into`, index.hpp:2650-2670). There the
order in which shards reach the per-query lock decides ties; here the merge is deterministic:
ascending distance, then shard rank, then position within the shard.

The data path lives in the library (usearch_b200/csrc/shards.cu): this shard's search writes one packed payload, ONE
`ncclAllGather` moves it, a merge kernel produces the global top-k on every rank. This module only
  * moves the 128-byte group id from rank 0 to the others over torch.distributed (`join`) - the control plane;
  * restates the merge in torch (`merge_gathered`) as the specification the merge kernel is tested against, usable on CPU
    tensors with the gloo backend (tests/test_sharded.py).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

SNAN_BITS = 0x7FA00000  # padding distance written by search_result_t::dump_to (index.hpp:2715-2720)


def shard_of(key: int, world: int) -> int:
    return int(key) % world


def join(index, group: Optional[dist.ProcessGroup] = None) -> None:
    """Make `index` (one per process) a shard of the group: rank 0 creates the id, a broadcast hands it out."""
    from .index import shards_unique_id
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [shards_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    index.join_shards(rank, world, box[0])


def merge_topk(keys: torch.Tensor, distances: torch.Tensor, counts: torch.Tensor, k: int,
               group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Specification of the exchange step in torch collectives: all-gather per-shard results, merge on every rank.

    keys: int64 [nq, k] (uint64 keys viewed as int64), distances: float32 [nq, k], counts: int32/int64 [nq].
    Rows are valid up to `counts`; the rest is padding (key 0, NaN). Returns the same triple, merged.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
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
    """Shard-major concatenation ordered by (padding last, NaN after numbers, distance, shard, position)."""
    cols = gk[0].shape[1]
    col = torch.arange(cols, device=gk[0].device)[None, :]
    cat_d = torch.cat(gd, dim=1)
    cat_k = torch.cat(gk, dim=1)
    valid = torch.cat([col < c.to(torch.int64)[:, None] for c in gc], dim=1)
    # rank of every candidate under the composite order, built from stable sorts (last key first)
    order = torch.sort(torch.nan_to_num(cat_d, nan=0.0, posinf=float("inf"), neginf=-float("inf")), dim=1, stable=True).indices
    is_nan = torch.gather(torch.isnan(cat_d), 1, order)
    order = torch.gather(order, 1, torch.sort(is_nan.to(torch.int8), dim=1, stable=True).indices)
    is_pad = ~torch.gather(valid, 1, order)
    order = torch.gather(order, 1, torch.sort(is_pad.to(torch.int8), dim=1, stable=True).indices)
    if order.shape[1] < k:  # fewer candidates than requested: the tail is padding by construction
        fill = torch.zeros((order.shape[0], k - order.shape[1]), dtype=order.dtype, device=order.device)
        order = torch.cat([order, fill], dim=1)
    order = order[:, :k]
    out_d = torch.gather(cat_d, 1, order)
    out_k = torch.gather(cat_k, 1, order)
    total = torch.stack([c.to(torch.int64) for c in gc], 0).sum(0).clamp(max=k)
    pad = torch.arange(k, device=out_d.device)[None, :] >= total[:, None]
    nan = torch.tensor(SNAN_BITS, dtype=torch.int32, device=out_d.device).view(torch.float32)
    out_d = torch.where(pad, nan.expand_as(out_d), out_d)
    out_k = torch.where(pad, torch.zeros_like(out_k), out_k)
    return out_k, out_d, total
