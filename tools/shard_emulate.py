"""One GPU, one process: build shard S of G of a bench workload exactly as bench.py would at N = G, search it, and measure recall
against exact ground truth restricted to that shard's rows (debug aid for the sharded bench)."""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch

G, S = int(sys.argv[1]), int(sys.argv[2])
sys.argv = ["bench.py", "--workload", sys.argv[3]] + sys.argv[4:]
a = bench.parse_args()
dev = torch.device("cuda", 0)
coll = bench.Collection(a, dev)
index, dt = bench.build_index_gpu(a, coll, S, G)
index.expansion_search = a.ef
q = coll.queries(2048)
qh = bench.to_numpy_queries(a, q)
res = index.search(qh, 10)
# truth over the shard only
qf = coll.as_float(q)
best_d = best_i = None
for ids, xq in coll.base_chunks(S, G):
    x = coll.as_float(xq)
    dist = (x.shape[1] - qf @ x.T) * 0.5 if a.metric == "hamming" else 1.0 - qf @ x.T
    d, i = torch.topk(dist, 10, dim=1, largest=False)
    i = ids[i]
    if best_d is None: best_d, best_i = d, i
    else:
        cd, ci = torch.cat([best_d, d], 1), torch.cat([best_i, i], 1)
        best_d, sel = torch.topk(cd, 10, dim=1, largest=False)
        best_i = torch.gather(ci, 1, sel)
rec = bench.recall_at_k(res.keys, res.counts, best_i.cpu().numpy().astype(np.uint64))
print(json.dumps({"shards": G, "shard": S, "size": len(index), "build_s": round(dt, 1), "recall_vs_shard_truth": round(rec, 4),
                  "counts_min": int(res.counts.min()), "keys_mod": sorted(set((res.keys[:50] % G).ravel().tolist()))[:4],
                  "row0": res.keys[0].tolist(), "truth0": best_i[0].tolist()}))
