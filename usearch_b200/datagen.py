"""Seeded synthetic collections for the parity tests and bench.py.

Dense data is rank-``r`` latent Gaussian, ``x = z W + sigma * eps``: i.i.d. high-dimensional
Gaussians cannot reach recall@10 >= 0.95 on an HNSW graph (SURVEY.md finding 9), low intrinsic
dimension can. Seeds: 42 base, 43 queries, 44 mixing matrix.
"""
from __future__ import annotations

import numpy as np


def latent(n: int, d: int, *, seed: int, rank: int = 16, sigma: float = 0.1, mix_seed: int = 44,
           chunk: int = 65536) -> np.ndarray:
    """``n x d`` float32 rows of rank-``rank`` latent Gaussian data."""
    w = np.random.default_rng(mix_seed).standard_normal((rank, d)).astype(np.float32)
    rng = np.random.default_rng(seed)
    out = np.empty((n, d), dtype=np.float32)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        z = rng.standard_normal((hi - lo, rank), dtype=np.float32)
        e = rng.standard_normal((hi - lo, d), dtype=np.float32)
        out[lo:hi] = z @ w + sigma * e
    return out


def to_scalar(x: np.ndarray, scalar: str) -> np.ndarray:
    """Quantise f32 rows into the index's scalar kind the way a user would before `add`."""
    if scalar == "f32":
        return np.ascontiguousarray(x, dtype=np.float32)
    if scalar == "f16":
        return x.astype(np.float16)
    if scalar == "bf16":
        bits = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
        rounded = bits + 0x7FFF + ((bits >> 16) & 1)  # round-to-nearest-even
        return (rounded >> 16).astype(np.uint16)
    if scalar == "i8":
        norm = np.linalg.norm(x.astype(np.float64), axis=1, keepdims=True)
        return np.clip(np.trunc(x.astype(np.float64) * 127.0 / norm), -127, 127).astype(np.int8)
    if scalar == "b1":
        return np.packbits(x > 0, axis=1, bitorder="big")
    raise ValueError(scalar)
