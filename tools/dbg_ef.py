import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, common
from oracle import bindings
from usearch_b200.index import Index
base, q = common.make_collection(2000, 64, "f32", 64, iid=True)
_, blob = common.build_reference_blob(base, "l2sq", "f32", 64, 16, threads=8)
index = Index.restore(blob)
e = index.search(q, 10, exact=True)
pe = bindings.PortIndex(blob, 64).search(q, 10, threads=4, exact=True)
print("exact gpu==port", np.array_equal(e.keys, pe[0]))
for ef in (256, 300, 512, 600, 1024, 1100, 2000, 4000):
    index.expansion_search = ef
    g = index.search(q, 10, stats=True)
    p = bindings.PortIndex(blob, ef).search(q, 10, threads=4)
    print(ef, "gpu==port keys", (g.keys == p[0]).mean(), "computed eq", np.array_equal(index.last_computed, p[3]),
          "launches", index.kernel_launches, "gpu vs exact", (g.keys == e.keys).mean(), "port vs exact", (p[0] == pe[0]).mean())
