"""Step-by-step replay of tests/test_zz_gpu_golden.py::test_next_rows_match_golden with timestamps and a watchdog
(faulthandler dumps the Python stack if a step takes more than 60 s). Round 1 ended with one unexplained time-out of
that test; run this FIRST in round 2, under `timeout 300`:

    gpurun --timeout 400 -- 'timeout 300 python tools/repro_golden_next_rows.py 2>&1 | tail -40'
"""
import faulthandler
import glob
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from oracle import bindings  # noqa: E402
from usearch_b200 import v2format  # noqa: E402
from usearch_b200.index import Index, exact_search  # noqa: E402

T0 = time.time()


def step(what):
    print(f"[{time.time() - T0:7.2f}s] {what}", flush=True)
    faulthandler.cancel_dump_traceback_later()
    faulthandler.dump_traceback_later(60, exit=True)


nr = np.load(os.path.join(common.GOLDEN, "next_rows.npz"))
for path in sorted(glob.glob(os.path.join(common.GOLDEN, "*_n*.npz"))):
    name = os.path.basename(path)[:-4]
    g = np.load(path)
    q, k = g["queries"], int(g["k"])
    step(f"{name}: restore")
    index = Index.restore(g["blob"])
    step(f"{name}: exact search (index mode), k={k}, nq={len(q)}")
    got = index.search(q, k, exact=True)
    ok = np.array_equal(got.keys, nr[f"{name}/exact_keys"]) and np.array_equal(
        got.distances.view(np.uint32), nr[f"{name}/exact_distances"].view(np.uint32))
    print("    exact matches golden:", ok, flush=True)
    for i, level in enumerate(nr[f"{name}/cluster_levels"]):
        step(f"{name}: cluster level {int(level)}")
        ck, cd = index.cluster(q, int(level), stats=True)
        ok = (np.array_equal(ck, nr[f"{name}/cluster_keys"][i]) and np.array_equal(index.last_computed, nr[f"{name}/cluster_computed"][i])
              and np.array_equal(cd.view(np.uint32), nr[f"{name}/cluster_distances"][i].view(np.uint32)))
        print("    cluster matches golden:", ok, flush=True)
    step(f"{name}: v2format.loads")
    graph = v2format.loads(g["blob"])
    vectors = graph.vectors.view(bindings.SCALAR_NP[graph.scalar]).reshape(graph.size, -1)
    step(f"{name}: free exact_search over {vectors.shape}")
    free = exact_search(vectors, q, k, metric=graph.metric, dtype=graph.scalar)
    ok = np.array_equal(free.distances.view(np.uint32), nr[f"{name}/free_distances"][:, :k].view(np.uint32))
    print("    free exact distances match golden:", ok, flush=True)
    step(f"{name}: graph search vs golden")
    index.expansion_search = int(g["ef"])
    res = index.search(q, k, stats=True)
    print("    graph search matches golden:", np.array_equal(res.keys, g["keys_pinned"]), flush=True)
    del index
faulthandler.cancel_dump_traceback_later()
print(f"[{time.time() - T0:7.2f}s] REPLAY_DONE", flush=True)
