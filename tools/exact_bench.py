"""Dev tool: wall-clock of the exact-search paths (not a bench line)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from usearch_b200.index import exact_search
from usearch_b200 import datagen

n, d, nq, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 10
scalar = sys.argv[4] if len(sys.argv) > 4 else "f32"
base = datagen.to_scalar(datagen.latent(n, d, seed=1, rank=16), scalar)
q = datagen.to_scalar(datagen.latent(nq, d, seed=2, rank=16), scalar)
for rep in range(3):
    t = time.perf_counter()
    m = exact_search(base, q, k, metric="cos", dtype=scalar)
    dt = time.perf_counter() - t
    print(f"exact_search n={n} d={d} nq={nq} {scalar}: {dt*1e3:.1f} ms wall (incl. H2D of {base.nbytes/1e9:.2f} GB), "
          f"{nq/dt:.0f} q/s, {n*nq*base.shape[1]*base.itemsize/dt/1e12:.2f} TB/s smem-side pair bytes")
