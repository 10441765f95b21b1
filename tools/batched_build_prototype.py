"""Research prototype for DESIGN.md §9 (GPU-assisted add): does BATCHED insertion — candidates for a whole batch taken
from the graph as it stood before the batch, plus brute force inside the batch — give the reference's search quality?

Everything here runs on the CPU with the oracle as the search engine (it stands in for the GPU search kernel), so this
is a development tool, not product code. It builds the same collection twice:
  * with the unmodified reference, sequentially (one thread);
  * with the batched scheme below (layer-0 candidates from PortIndex.search on the pre-batch graph + exact distances
    to the earlier members of the batch; link selection = the `refine_` heuristic; reverse links with re-pruning),
writes both as v2 files and measures recall@10 and computed_distances of the REFERENCE search on each.

usage: python tools/batched_build_prototype.py [n] [d] [M] [batch_fraction] [intra_batch: 1|0] [latent_rank]
"""
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bindings  # noqa: E402
from usearch_b200 import datagen, v2format  # noqa: E402


def l2sq(a, b):
    diff = a - b
    return np.einsum("...i,...i->...", diff, diff)


def refine(vectors, new_vec, cand, cand_d, limit):
    """index.hpp:4276-4318: walk candidates by distance, keep one unless an already kept neighbour is closer to it."""
    order = np.argsort(cand_d, kind="stable")
    kept = []
    for i in order:
        c = cand[i]
        if len(kept) >= limit:
            break
        if kept:
            d_to_kept = l2sq(vectors[kept], vectors[c])
            if (d_to_kept < cand_d[i]).any():
                continue
        kept.append(int(c))
    return kept


def build_batched(base, m, ef_add, batch_fraction, seed=1, intra_batch=True):
    n, d = base.shape
    m0 = 2 * m
    rng = np.random.default_rng(seed)
    levels = np.minimum((-np.log(rng.random(n)) * (1.0 / math.log(m))).astype(np.int64), 12)
    nbrs = [[[] for _ in range(int(levels[i]) + 1)] for i in range(n)]
    entry, max_level = 0, int(levels[0])
    done = 1
    searches = 0
    while done < n:
        batch = list(range(done, min(n, done + max(1, int(done * batch_fraction)))))
        # a member above the current top level is inserted alone (it becomes the entry point)
        for j, i in enumerate(batch):
            if levels[i] > max_level and j > 0:
                batch = batch[:j]
                break
        # the graph as it stands BEFORE the batch, searched by the oracle (stand-in for the GPU kernel)
        graph = v2format.Graph("l2sq", "f32", d, m, m0, base[:done].view(np.uint8).reshape(done, -1),
                               np.arange(done, dtype=np.uint64), levels[:done].astype(np.int16),
                               [nbrs[i] for i in range(done)], max_level, entry)
        port = bindings.PortIndex(v2format.dumps(graph), ef_add)
        k = min(ef_add, done)
        found, found_d, counts = port.search(base[batch], k, threads=8)[:3]
        searches += len(batch)
        for j, i in enumerate(batch):
            vec = base[i]
            for level in range(min(int(levels[i]), max_level), -1, -1):
                if level == 0:
                    cand = found[j, :int(counts[j])].astype(np.int64)
                    cand_d = found_d[j, :int(counts[j])].astype(np.float64)
                else:  # upper levels hold 1/M of the members: exact candidates among those present on the level
                    pool = np.array([s for s in range(done) if levels[s] >= level], dtype=np.int64)
                    cand_d = l2sq(base[pool], vec).astype(np.float64)
                    top = np.argsort(cand_d, kind="stable")[:ef_add]
                    cand, cand_d = pool[top], cand_d[top]
                earlier = np.array([s for s in batch[:j] if levels[s] >= level and intra_batch], dtype=np.int64)
                if earlier.size:  # what the frozen graph cannot see: the batch itself, by brute force
                    cand = np.concatenate([cand, earlier])
                    cand_d = np.concatenate([cand_d, l2sq(base[earlier], vec).astype(np.float64)])
                    top = np.argsort(cand_d, kind="stable")[:ef_add]
                    cand, cand_d = cand[top], cand_d[top]
                limit = m0 if level == 0 else m
                chosen = refine(base, vec, cand, cand_d, m)        # form_links_to_closest_: connectivity, not base
                nbrs[i][level] = chosen
                for c in chosen:                                    # form_reverse_links_ (index.hpp:3848-3900)
                    lst = nbrs[c][level]
                    if i in lst:
                        continue
                    if len(lst) < limit:
                        lst.append(i)
                    else:
                        pool = np.array(lst + [i], dtype=np.int64)
                        nbrs[c][level] = refine(base, base[c], pool, l2sq(base[pool], base[c]).astype(np.float64), limit)
            if levels[i] > max_level:
                entry, max_level = i, int(levels[i])
        done += len(batch)
    graph = v2format.Graph("l2sq", "f32", d, m, m0, base.view(np.uint8).reshape(n, -1), np.arange(n, dtype=np.uint64),
                           levels.astype(np.int16), nbrs, max_level, entry)
    return v2format.dumps(graph), searches


def evaluate(name, blob, queries, truth, ef, k=10):
    ref = bindings.RefIndex("parity")
    ref.load(blob)
    ref.change_expansion_search(ef)
    keys, _, counts, computed, visited = ref.search(queries, k, threads=8)
    recall = np.mean([len(set(keys[i, :int(counts[i])].tolist()) & set(truth[i].tolist())) / k for i in range(len(queries))])
    print(f"{name:34s} ef={ef:4d} recall@{k} {recall:.4f}  computed_distances/query {computed.mean():8.1f}  hops/query {visited.mean():6.1f}")
    return recall


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    m = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    fraction = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
    intra = (sys.argv[5] != "0") if len(sys.argv) > 5 else True
    rank = int(sys.argv[6]) if len(sys.argv) > 6 else min(16, d)
    ef_add = 128
    base = datagen.latent(n, d, seed=42, rank=rank).astype(np.float32)
    queries = datagen.latent(500, d, seed=43, rank=rank).astype(np.float32)
    dist = ((queries ** 2).sum(1)[:, None] - 2 * queries @ base.T + (base ** 2).sum(1)[None, :])
    truth = np.argsort(dist, axis=1)[:, :10]

    t = time.time()
    ref = bindings.RefIndex("parity", metric="l2sq", scalar="f32", dims=d, connectivity=m, expansion_add=ef_add, expansion_search=64)
    ref.add(np.arange(n, dtype=np.uint64), base, threads=1)
    sequential = ref.save()
    print(f"reference, sequential build: {time.time() - t:.1f} s")
    t = time.time()
    batched, searches = build_batched(base, m, ef_add, fraction, intra_batch=intra)
    print(f"batched prototype (batch <= {fraction:.0%} of the current size, intra-batch candidates {'on' if intra else 'OFF'}): "
          f"{time.time() - t:.1f} s, {searches} candidate searches")
    for ef in (16, 64, 128):
        evaluate("reference-built graph", sequential, queries, truth, ef)
        evaluate(f"batched-built graph ({fraction:.0%})", batched, queries, truth, ef)


if __name__ == "__main__":
    main()
