#!/usr/bin/env python
"""Development helper (GPU box): time the search kernel on the bench workload and print the
per-phase cycle breakdown. Not part of the product or of the tests.

    python tools/experiment.py [--n 1000000] [--batches 6] [--env KEY=VALUE ...]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--n", type=int, default=1_000_000)
    p.add_argument("--dim", type=int, default=768)
    p.add_argument("--batch", type=int, default=4096)
    p.add_argument("--batches", type=int, default=6)
    p.add_argument("--ef", type=int, default=128)
    p.add_argument("--connectivity", type=int, default=32)
    p.add_argument("--dtype", default="f32")
    p.add_argument("--metric", default="cos")
    p.add_argument("--phases", action="store_true")
    args = p.parse_args()
    a = argparse.Namespace(n=args.n, dim=args.dim, metric=args.metric, dtype=args.dtype, connectivity=args.connectivity,
                           expansion_add=128, ef=args.ef, batch=args.batch, k=10, rank_latent=16)
    import torch
    from usearch_b200.index import Index
    path = os.path.join(bench.CACHE, f"index_{bench.shard_key(a, 0, 1)}.usearch")
    if not os.path.exists(path):  # build once per gpurun call; later invocations skip the 12 s of data generation
        bench.get_index_blob(a, 0, 1, bench.host_threads())
    index = Index.restore(path)
    index.expansion_search = a.ef
    B, k = a.batch, a.k
    total = args.batches * B
    queries = bench.make_queries(a, total)
    bpv = queries.strides[0]
    vs = (bpv + 15) // 16 * 16
    dev = torch.device("cuda:0")
    q_dev = torch.zeros((total, vs), dtype=torch.uint8, device=dev)
    q_dev[:, :bpv] = torch.from_numpy(queries.view(np.uint8).reshape(total, bpv)).to(dev)
    keys_dev = torch.zeros((B, k), dtype=torch.int64, device=dev)
    dist_dev = torch.zeros((B, k), dtype=torch.float32, device=dev)
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    comp = torch.zeros(B, dtype=torch.int32, device=dev)
    vis = torch.zeros(B, dtype=torch.int32, device=dev)
    ms = []
    for s in range(args.batches):
        if args.phases and s == 2:
            index.profile_phases(True)
        index.search_device(q_dev[s * B:(s + 1) * B].data_ptr(), B, vs, k, keys_dev.data_ptr(), dist_dev.data_ptr(),
                            cnt.data_ptr(), comp.data_ptr(), vis.data_ptr(), 0)
        ms.append(index.last_kernel_ms)
    D, H = float(comp.sum()) / B, float(vis.sum()) / B
    alg = (D * index.bytes_per_vector + H * (4 + 8 * index.connectivity)) * B
    best = min(ms[2:]) if len(ms) > 2 else min(ms)
    out = {"kernel_ms": [round(x, 3) for x in ms], "best_ms": round(best, 3), "qps": round(B / best * 1e3),
           "alg_GBps": round(alg / best / 1e6, 1), "D": round(D, 1), "H": round(H, 1), "launches": index.kernel_launches,
           "env": {k: v for k, v in os.environ.items() if k.startswith("USEARCH_B200_")}}
    if args.phases:
        ph = index.profile_phases(False)
        extra = ("queries", "pushes", "avg_max_heap", "max_heap")
        tot = sum(v for kk, v in ph.items() if kk not in extra)
        out["phase_cycles_per_query"] = {kk: round(v) for kk, v in ph.items()}
        out["phase_share"] = {kk: round(v / tot, 3) for kk, v in ph.items() if kk not in extra}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
