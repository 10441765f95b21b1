"""Shared helpers for the parity tests: build an index with the reference, hand the SAME serialised
graph to the oracle(s) and to the GPU backend, compare bit-for-bit."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import bindings  # noqa: E402
from usearch_b200 import datagen  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def have_reference() -> bool:
    return bindings.ref_lib("parity") is not None


def make_collection(n: int, d: int, scalar: str, nq: int, *, seed: int = 42, rank: int = 16, iid: bool = False):
    if scalar == "b1":
        base = datagen.latent(n, d, seed=seed, rank=min(rank, d))
        queries = datagen.latent(nq, d, seed=seed + 1, rank=min(rank, d))
    elif iid:
        base = np.random.default_rng(seed).standard_normal((n, d), dtype=np.float32)
        queries = np.random.default_rng(seed + 1).standard_normal((nq, d), dtype=np.float32)
    else:
        base = datagen.latent(n, d, seed=seed, rank=min(rank, d))
        queries = datagen.latent(nq, d, seed=seed + 1, rank=min(rank, d))
    return datagen.to_scalar(base, scalar), datagen.to_scalar(queries, scalar)


def build_reference_blob(base: np.ndarray, metric: str, scalar: str, d: int, connectivity: int, expansion_add: int = 128,
                         threads: int = 8, keys: np.ndarray | None = None):
    ref = bindings.RefIndex("parity", metric=metric, scalar=scalar, dims=d, connectivity=connectivity,
                            expansion_add=expansion_add, expansion_search=64)
    keys = np.arange(len(base), dtype=np.uint64) if keys is None else keys
    ref.add(keys, base, threads=threads)
    return ref, ref.save()


def assert_same_results(a, b, what: str = ""):
    """(keys, distances, counts, computed, visited) tuples must agree bit-for-bit."""
    ak, ad, ac = a[0], a[1], a[2]
    bk, bd, bc = b[0], b[1], b[2]
    assert np.array_equal(ac.astype(np.uint64), bc.astype(np.uint64)), f"{what}: counts differ"
    bad = np.nonzero((ak != bk).any(axis=1))[0]
    assert bad.size == 0, f"{what}: labels differ for {bad.size} queries, first {bad[:5]}"
    assert np.array_equal(ad.view(np.uint32), bd.view(np.uint32)), f"{what}: distance bits differ"
    if len(a) > 3 and len(b) > 3 and a[3] is not None and b[3] is not None:
        assert np.array_equal(np.asarray(a[3], dtype=np.uint64), np.asarray(b[3], dtype=np.uint64)), f"{what}: computed_distances differ"
        assert np.array_equal(np.asarray(a[4], dtype=np.uint64), np.asarray(b[4], dtype=np.uint64)), f"{what}: visited_members differ"
