"""Reader / writer for the reference's v2 ``.usearch`` serialisation (numpy, host side).

Layout (/root/reference/include/usearch):
  index_dense.hpp:994-1062   [u32 rows, u32 cols] [rows x cols vector bytes, slot order] [64-byte head]
  index_dense.hpp:42-79      head: "usearch" | 3 x u16 version | metric u8 | scalar u8 | key kind u8 |
                             slot kind u8 | count_present u64 | count_deleted u64 | dimensions u64 | multi u8
  index.hpp:3276-3317        [size, connectivity, connectivity_base, max_level, entry_slot : 5 x u64]
                             [int16 level per node] [node tapes]
  index.hpp:2116-2195        tape: key u64 | level i16 | {u32 n, slot[M0]} | level x {u32 n, slot[M]}

Used by tests to hand-craft graphs (ties, empty lists, single node) that no builder would produce,
and by tools that want to inspect an index without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

METRIC_CHAR = {"ip": ord("i"), "cos": ord("c"), "l2sq": ord("e"), "hamming": ord("b"), "tanimoto": ord("t"),
               "sorensen": ord("s"), "jaccard": ord("j")}
SCALAR_CODE = {"b1": 1, "bf16": 4, "f64": 10, "f32": 11, "f16": 12, "i8": 23}
SCALAR_BITS = {"b1": 1, "bf16": 16, "f64": 64, "f32": 32, "f16": 16, "i8": 8}
FREE_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)


@dataclass
class Graph:
    metric: str
    scalar: str
    dimensions: int
    connectivity: int
    connectivity_base: int
    vectors: np.ndarray                      # [n, bytes_per_vector] uint8
    keys: np.ndarray                         # [n] uint64
    levels: np.ndarray                       # [n] int16
    neighbors: List[List[List[int]]] = field(default_factory=list)  # [slot][level] -> list of slots
    max_level: int = 0
    entry_slot: int = 0
    multi: bool = False

    @property
    def size(self) -> int:
        return len(self.keys)

    @property
    def bytes_per_vector(self) -> int:
        return (self.dimensions * SCALAR_BITS[self.scalar] + 7) // 8


def dumps(g: Graph) -> np.ndarray:
    n, bpv = g.size, g.bytes_per_vector
    vectors = np.ascontiguousarray(g.vectors).view(np.uint8).reshape(n, bpv) if n else np.zeros((0, bpv), np.uint8)
    parts = [np.array([n, bpv], dtype=np.uint32).tobytes(), vectors.tobytes()]
    head = bytearray(64)
    head[0:7] = b"usearch"
    head[7:13] = np.array([2, 21, 0], dtype=np.uint16).tobytes()
    head[13], head[14], head[15], head[16] = METRIC_CHAR[g.metric], SCALAR_CODE[g.scalar], 14, 15
    deleted = int((np.asarray(g.keys, dtype=np.uint64) == FREE_KEY).sum())
    head[17:25] = np.uint64(n - deleted).tobytes()
    head[25:33] = np.uint64(deleted).tobytes()
    head[33:41] = np.uint64(g.dimensions).tobytes()
    head[41] = 1 if g.multi else 0
    parts.append(bytes(head))
    parts.append(np.array([n, g.connectivity, g.connectivity_base, g.max_level, g.entry_slot], dtype=np.uint64).tobytes())
    parts.append(np.asarray(g.levels, dtype=np.int16).tobytes())
    for slot in range(n):
        tape = bytearray()
        tape += np.uint64(g.keys[slot]).tobytes()
        tape += np.int16(g.levels[slot]).tobytes()
        for level in range(int(g.levels[slot]) + 1):
            cap = g.connectivity_base if level == 0 else g.connectivity
            lst = list(g.neighbors[slot][level])
            assert len(lst) <= cap
            row = np.zeros(cap + 1, dtype=np.uint32)
            row[0] = len(lst)
            row[1:1 + len(lst)] = lst
            tape += row.tobytes()
        parts.append(bytes(tape))
    return np.frombuffer(b"".join(parts), dtype=np.uint8).copy()


def loads(blob) -> Graph:
    b = np.ascontiguousarray(blob, dtype=np.uint8)
    raw = b.tobytes()
    rows, cols = np.frombuffer(raw, dtype=np.uint32, count=2)
    rows, cols = int(rows), int(cols)
    off = 8
    vectors = np.frombuffer(raw, dtype=np.uint8, count=rows * cols, offset=off).reshape(rows, cols).copy()
    off += rows * cols
    head = raw[off:off + 64]
    if head[0:7] != b"usearch":
        raise ValueError("Magic header mismatch - the file isn't an index")
    metric = {v: k for k, v in METRIC_CHAR.items()}[head[13]]
    scalar = {v: k for k, v in SCALAR_CODE.items()}[head[14]]
    dims = int(np.frombuffer(head, dtype=np.uint64, count=1, offset=33)[0]) if False else int.from_bytes(head[33:41], "little")
    multi = bool(head[41])
    off += 64
    n, m, m0, max_level, entry = (int(x) for x in np.frombuffer(raw, dtype=np.uint64, count=5, offset=off))
    off += 40
    levels = np.frombuffer(raw, dtype=np.int16, count=n, offset=off).copy()
    off += 2 * n
    keys = np.zeros(n, dtype=np.uint64)
    neighbors: List[List[List[int]]] = []
    for slot in range(n):
        keys[slot] = int.from_bytes(raw[off:off + 8], "little")
        off += 10
        per_level = []
        for level in range(int(levels[slot]) + 1):
            cap = m0 if level == 0 else m
            row = np.frombuffer(raw, dtype=np.uint32, count=cap + 1, offset=off) if off % 4 == 0 else \
                np.frombuffer(raw[off:off + 4 * (cap + 1)], dtype=np.uint32)
            per_level.append([int(x) for x in row[1:1 + int(row[0])]])
            off += 4 * (cap + 1)
        neighbors.append(per_level)
    return Graph(metric, scalar, dims, m, m0, vectors, keys, levels, neighbors, max_level, entry, multi)
