"""Golden vectors for the "next" rows — exact search and cluster — produced by RUNNING THE
UNMODIFIED REFERENCE (oracle/_ref, metric pinned) on the graphs of the committed search fixtures.

Run in the development container (needs /root/reference):  python tests/golden/make_golden_next_rows.py
``next_rows.npz`` holds, per fixture `<name>`:
  <name>/exact_keys, exact_distances, exact_counts     index_dense_gt::search(exact = true), k = fixture k
  <name>/cluster_levels, cluster_keys[L], cluster_distances[L], cluster_computed[L], cluster_visited[L]
  <name>/free_keys, free_distances                      exact_search_t over (vectors of the graph, queries), k + 1
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings  # noqa: E402
from usearch_b200 import v2format  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(HERE, "*_n*.npz"))):
        name = os.path.basename(path)[:-4]
        g = np.load(path)
        blob, queries, k = g["blob"], g["queries"], int(g["k"])
        ref = bindings.RefIndex("parity")
        ref.load(blob)
        ref.pin_metric(True)
        keys, dist, counts = ref.search(queries, k, threads=1, exact=True)[:3]
        out[f"{name}/exact_keys"], out[f"{name}/exact_distances"], out[f"{name}/exact_counts"] = keys, dist, counts
        levels = list(range(0, ref.max_level + 2))
        out[f"{name}/cluster_levels"] = np.array(levels)
        res = [ref.cluster(queries, level) for level in levels]
        for i, tag in enumerate(("keys", "distances", "computed", "visited")):
            out[f"{name}/cluster_{tag}"] = np.stack([r[i] for r in res])
        graph = v2format.loads(blob)
        scalar = graph.scalar
        vectors = graph.vectors.view(bindings.SCALAR_NP[scalar]).reshape(graph.size, -1)
        fk, fd = bindings.ref_exact_search(vectors, queries, k + 1, metric=graph.metric, scalar=scalar,
                                           dims=graph.dimensions, pinned=True)
        out[f"{name}/free_keys"], out[f"{name}/free_distances"] = fk, fd
        print(f"{name}: exact {keys.shape}, cluster levels {levels}, free {fk.shape}")
    np.savez_compressed(os.path.join(HERE, "next_rows.npz"), **out)


if __name__ == "__main__":
    main()
