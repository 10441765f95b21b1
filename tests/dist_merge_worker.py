"""Worker for tests/test_sharded.py: world_size-2 gloo run of the sharded merge on CPU tensors."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from usearch_b200.sharded import SNAN_BITS, merge_topk  # noqa: E402


def fake_shard(rank: int, nq: int, k: int):
    """Per-shard results with heavy ties (integer distances, like Hamming) and short rows."""
    rng = np.random.default_rng(100 + rank)
    d = np.sort(rng.integers(0, 6, size=(nq, k)).astype(np.float32), axis=1)
    keys = (rng.permutation(nq * k).reshape(nq, k) * 2 + rank).astype(np.int64)  # shard = key mod 2
    counts = rng.integers(0, k + 1, size=nq).astype(np.int64)
    counts[0], counts[1] = 0, k
    pad = np.arange(k)[None, :] >= counts[:, None]
    d[pad] = np.array(SNAN_BITS, dtype=np.uint32).view(np.float32)
    keys[pad] = 0
    return keys, d, counts


def reference_merge(shards, k):
    nq = shards[0][0].shape[0]
    out_k = np.zeros((nq, k), np.int64)
    out_d = np.full((nq, k), np.array(SNAN_BITS, dtype=np.uint32).view(np.float32), np.float32)
    out_c = np.zeros(nq, np.int64)
    for q in range(nq):
        items = []
        for r, (keys, d, c) in enumerate(shards):
            items += [(float(d[q, i]), r, i, int(keys[q, i])) for i in range(int(c[q]))]
        items.sort(key=lambda t: (t[0], t[1], t[2]))
        items = items[:k]
        out_c[q] = len(items)
        for i, (dd, _, _, kk) in enumerate(items):
            out_k[q, i], out_d[q, i] = kk, dd
    return out_k, out_d, out_c


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    nq, k = 64, 10
    keys, d, counts = fake_shard(rank, nq, k)
    mk, md, mc = merge_topk(torch.from_numpy(keys), torch.from_numpy(d), torch.from_numpy(counts), k)
    want_k, want_d, want_c = reference_merge([fake_shard(r, nq, k) for r in range(world)], k)
    assert np.array_equal(mc.numpy(), want_c), "counts"
    assert np.array_equal(mk.numpy(), want_k), "keys"
    assert np.array_equal(md.numpy().view(np.uint32), want_d.view(np.uint32)), "distances"
    dist.barrier()
    if rank == 0:
        print("MERGE_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
