"""GPU-assisted add against the reference's own build (DESIGN.md §9 bar), on one GPU.

For each case: build the same collection (a) with the unmodified reference on the host and (b) with
`Index.add` on the GPU; save (b) in the v2 format; check its structure; then let the REFERENCE search both
files and compare recall@10 and computed_distances per query at several ef.

    python tools/build_check.py [--big]      # --big adds the 1M x 768 timing case
"""
import argparse
import json
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
from usearch_b200.index import Index  # noqa: E402


def structure_report(blob) -> dict:
    g = v2format.loads(blob)
    n = g.size
    problems = []
    deg0 = np.zeros(n, dtype=np.int64)
    indeg0 = np.zeros(n, dtype=np.int64)
    for s in range(n):
        for level, lst in enumerate(g.neighbors[s]):
            cap = g.connectivity_base if level == 0 else g.connectivity
            if len(lst) > cap:
                problems.append(f"slot {s} level {level}: degree {len(lst)} > {cap}")
            if s in lst:
                problems.append(f"slot {s} level {level}: self link")
            if len(set(lst)) != len(lst):
                problems.append(f"slot {s} level {level}: repeated neighbour")
            for t in lst:
                if t >= n:
                    problems.append(f"slot {s} level {level}: neighbour {t} out of range")
                elif g.levels[t] < level:
                    problems.append(f"slot {s} level {level}: neighbour {t} lives below that level")
            if level == 0:
                deg0[s] = len(lst)
                for t in lst:
                    if t < n:
                        indeg0[t] += 1
    if n and g.levels[g.entry_slot] != g.max_level:
        problems.append("entry point is not on the top level")
    return {"n": n, "max_level": int(g.max_level), "mean_degree0": float(deg0.mean()), "min_degree0": int(deg0.min()),
            "isolated_in0": int((indeg0 == 0).sum()), "problems": problems[:10], "n_problems": len(problems)}


def exact_truth(base, queries, metric, k):
    x = base.astype(np.float32)
    q = queries.astype(np.float32)
    if metric == "cos":
        x = x / np.linalg.norm(x, axis=1, keepdims=True)
        q = q / np.linalg.norm(q, axis=1, keepdims=True)
        d = 1.0 - q @ x.T
    elif metric == "ip":
        d = 1.0 - q @ x.T
    else:
        d = (q * q).sum(1)[:, None] - 2 * q @ x.T + (x * x).sum(1)[None, :]
    return np.argsort(d, axis=1, kind="stable")[:, :k].astype(np.uint64)


def hamming_truth(base, queries, k):
    d = np.unpackbits(queries[:, None, :] ^ base[None, :, :], axis=2).sum(axis=2)
    return np.argsort(d, axis=1, kind="stable")[:, :k].astype(np.uint64)


def recall(found, truth):
    return float(np.mean([len(set(f.tolist()) & set(t.tolist())) / len(t) for f, t in zip(found, truth)]))


def case(name, n, d, metric, scalar, m, nq=1000, efs=(16, 64, 128), threads=8):
    base, queries = common.make_collection(n, d, scalar, nq)
    keys = np.arange(n, dtype=np.uint64)
    t0 = time.time()
    ref, ref_blob = common.build_reference_blob(base, metric, scalar, d, m, threads=threads)
    t_ref = time.time() - t0
    t0 = time.time()
    index = Index(ndim=d, metric=metric, dtype=scalar, connectivity=m, expansion_add=128)
    index.add(keys, base)
    t_gpu = time.time() - t0
    gpu_blob = index.save()
    rep = structure_report(gpu_blob)
    rep["reference_mean_degree0"] = structure_report(ref_blob)["mean_degree0"]
    truth = hamming_truth(base, queries, 10) if scalar == "b1" else exact_truth(
        base if scalar != "bf16" else (base.astype(np.uint32) << 16).view(np.float32),
        queries if scalar != "bf16" else (queries.astype(np.uint32) << 16).view(np.float32), metric, 10)
    out = {"case": name, "n": n, "d": d, "metric": metric, "scalar": scalar, "M": m, "build_s_reference": round(t_ref, 2),
           "reference_threads": threads, "build_s_gpu": round(t_gpu, 2), "structure": rep, "ef": {}}
    searcher = bindings.RefIndex("parity")
    for label, blob in (("reference_built", ref_blob), ("gpu_built", gpu_blob)):
        searcher.load(blob)
        for ef in efs:
            searcher.change_expansion_search(ef)
            k, dist, cnt, comp, vis = searcher.search(queries, 10, threads=threads)
            out["ef"].setdefault(str(ef), {})[label] = {"recall_at_10": round(recall(k, truth), 4),
                                                         "computed_distances": round(float(comp.mean()), 1)}
    # our own search on our own graph must agree with the reference search on it (same graph => same labels)
    index.expansion_search = 64
    got = index.search(queries, 10, stats=True)
    searcher.load(gpu_blob)
    searcher.pin_metric(True)
    searcher.change_expansion_search(64)
    want = searcher.search(queries, 10, threads=threads)
    out["gpu_search_on_gpu_graph_matches_reference"] = bool(np.array_equal(got.keys, want[0]) and
                                                            np.array_equal(index.last_computed, want[3]))
    # lookups and edits by key
    some = int(keys[n // 3])
    vec = index.get(some)
    edits = {"contains": bool(index.contains(some)), "count": index.count(some), "missing": bool(index.contains(n + 5)),
             "get_matches_input": bool(vec is not None and np.array_equal(np.asarray(vec).view(np.uint8).ravel(),
                                                                         base[n // 3].view(np.uint8).ravel()))}
    edits["removed"] = index.remove(some)
    edits["size_after_remove"] = len(index)
    edits["contains_after_remove"] = bool(index.contains(some))
    again = index.search(base[n // 3], 10)
    edits["removed_key_absent_from_results"] = bool(some not in again.keys.tolist())
    edits["renamed"] = index.rename(int(keys[n // 3 + 1]), n + 100)
    edits["renamed_found"] = bool(index.contains(n + 100))
    out["edits"] = edits
    print(json.dumps(out), flush=True)
    return out


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--big", action="store_true")
    p.add_argument("--only-tiny", action="store_true")
    p.add_argument("--cases", default="", help="comma-separated case names (default: all)")
    a = p.parse_args()
    cases = {
        "tiny": (300, 32, "cos", "f32", 8, 100),
        "latent32": (20000, 32, "cos", "f32", 16, 1000),
        "l2_128": (20000, 128, "l2sq", "f32", 16, 1000),
        "f16": (20000, 96, "cos", "f16", 16, 1000),
        "bf16": (10000, 64, "ip", "bf16", 16, 1000),
        "i8": (20000, 128, "ip", "i8", 16, 1000),
        "b1": (20000, 256, "hamming", "b1", 32, 1000),
        "wide": (20000, 768, "cos", "f32", 32, 500),
        "l2_100k": (100000, 128, "l2sq", "f32", 16, 1000),
    }
    wanted = ["tiny"] if a.only_tiny else ([c for c in a.cases.split(",") if c] or [c for c in cases if c != "l2_100k"])
    for name in wanted:
        n, d, metric, scalar, m, nq = cases[name]
        case(name, n, d, metric, scalar, m, nq=nq)
    if a.big:
        n, d = 1_000_000, 768
        from usearch_b200 import datagen
        base = datagen.latent(n, d, seed=42)
        t0 = time.time()
        index = Index(ndim=d, metric="cos", dtype="f32", connectivity=32, expansion_add=128)
        index.reserve(n)
        index.add(np.arange(n, dtype=np.uint64), base)
        t_gpu = time.time() - t0
        queries = datagen.latent(4096, d, seed=43)
        index.expansion_search = 128
        got = index.search(queries, 10, stats=True)
        print(json.dumps({"case": "1M x 768 f32 cos M=32", "build_s_gpu": round(t_gpu, 2), "max_level": index.max_level,
                          "computed_distances": float(index.last_computed.mean()), "kernel_ms": index.last_kernel_ms}), flush=True)


if __name__ == "__main__":
    main()
