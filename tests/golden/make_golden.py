"""Generate the committed golden fixtures by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref).

Run in the development container (needs /root/reference):  python tests/golden/make_golden.py
Each ``<name>.npz`` holds the serialised v2 index (`blob`), the queries, and what the reference's
own `index_dense_gt::search` returned for them at the recorded ef / k:
  * ``*_pinned``  with the metric pinned to oracle/metrics_pinned.h (the label-parity target),
  * ``*_native``  with the reference's builtin SimSIMD dispatch on the generating host (`isa`).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings  # noqa: E402
from usearch_b200 import datagen  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, metric, scalar, n, d, M, ef, k, nq, removed
    ("cos_f32_n2000_d64", "cos", "f32", 2000, 64, 16, 64, 10, 64, 0),
    ("l2sq_f32_n2000_d33", "l2sq", "f32", 2000, 33, 8, 32, 10, 64, 0),
    ("ip_f32_n1500_d48_removed", "ip", "f32", 1500, 48, 12, 48, 10, 64, 150),
    ("ip_i8_n2000_d64", "ip", "i8", 2000, 64, 16, 64, 10, 64, 0),
    ("hamming_b1_n4000_d256", "hamming", "b1", 4000, 256, 16, 64, 10, 64, 0),
    ("tanimoto_b1_n2000_d96", "tanimoto", "b1", 2000, 96, 8, 32, 5, 64, 0),
]


def main():
    for name, metric, scalar, n, d, m, ef, k, nq, removed in CASES:
        base = datagen.to_scalar(datagen.latent(n, d, seed=42, rank=min(16, d)), scalar)
        queries = datagen.to_scalar(datagen.latent(nq, d, seed=43, rank=min(16, d)), scalar)
        ref = bindings.RefIndex("parity", metric=metric, scalar=scalar, dims=d, connectivity=m, expansion_add=128,
                                expansion_search=ef)
        ref.add(np.arange(n, dtype=np.uint64), base, threads=1)  # single thread: reproducible graph
        for key in range(0, removed * 3, 3):
            ref.remove(key)
        blob = ref.save()
        native = ref.search(queries, k, threads=1)
        isa = ref.isa_name
        ref.pin_metric(True)
        pinned = ref.search(queries, k, threads=1)
        out = dict(blob=blob, queries=queries, ef=ef, k=k, isa=isa)
        for tag, res in (("native", native), ("pinned", pinned)):
            out[f"keys_{tag}"], out[f"distances_{tag}"], out[f"counts_{tag}"] = res[0], res[1], res[2]
            out[f"computed_{tag}"], out[f"visited_{tag}"] = res[3], res[4]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        same = np.array_equal(native[0], pinned[0])
        print(f"{name}: blob {blob.size} B, isa {isa}, native labels == pinned labels: {same}")


if __name__ == "__main__":
    main()
