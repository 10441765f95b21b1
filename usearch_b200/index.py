"""Host-side mirror of the reference's Python search API over the C ABI.

Mirrors, for the search path only, ``usearch.index.Index`` (/root/reference/python/usearch/index.py):
``Index.search`` (index.py:700-748) → ``_search_in_compiled`` (:191-231) → ``search_many``
(python/lib.cpp:415-461), and the result containers ``Matches`` / ``BatchMatches`` (:300-396).
Same argument names and meaning, same shapes and dtypes of the results, same padding of short
rows. Everything below goes through ``libusearch_b200.so`` with plain pointers (ctypes); there is
no CPU fallback: without the CUDA library or without a GPU the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Union

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libusearch_b200.so")

# usearch.h:40-62 ordinals
METRIC_KIND = {"cos": 1, "ip": 2, "l2sq": 3, "haversine": 4, "divergence": 5, "pearson": 6, "jaccard": 7,
               "hamming": 8, "tanimoto": 9, "sorensen": 10}
SCALAR_KIND = {"f32": 1, "f64": 2, "f16": 3, "i8": 4, "b1": 5, "bf16": 6}
_NP_TO_SCALAR = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64", np.dtype(np.float16): "f16",
                 np.dtype(np.int8): "i8", np.dtype(np.uint8): "b1"}
_BITS = {"f32": 32, "f64": 64, "f16": 16, "bf16": 16, "i8": 8, "b1": 1}


class _InitOptions(C.Structure):  # usearch.h:64-110
    _fields_ = [("metric_kind", C.c_int), ("metric", C.c_void_p), ("quantization", C.c_int),
                ("dimensions", C.c_size_t), ("connectivity", C.c_size_t), ("expansion_add", C.c_size_t),
                ("expansion_search", C.c_size_t), ("multi", C.c_bool)]


_lib: Optional[C.CDLL] = None

EXPORTED_SYMBOLS = [
    "usearch_version", "usearch_init", "usearch_free", "usearch_memory_usage", "usearch_hardware_acceleration",
    "usearch_serialized_length", "usearch_save", "usearch_load", "usearch_view", "usearch_metadata",
    "usearch_save_buffer", "usearch_load_buffer", "usearch_view_buffer", "usearch_metadata_buffer", "usearch_size",
    "usearch_capacity", "usearch_dimensions", "usearch_connectivity", "usearch_reserve", "usearch_expansion_add",
    "usearch_expansion_search", "usearch_change_expansion_add", "usearch_change_expansion_search",
    "usearch_change_threads_add", "usearch_change_threads_search", "usearch_change_metric_kind",
    "usearch_change_metric", "usearch_add", "usearch_contains", "usearch_count", "usearch_search",
    "usearch_filtered_search", "usearch_get", "usearch_remove", "usearch_rename", "usearch_distance",
    "usearch_exact_search", "usearch_clear",
    # additive
    "usearch_search_many", "usearch_b200_search_many_device", "usearch_b200_search_many_stats",
    "usearch_b200_filtered_search_many", "usearch_b200_exact_search_many", "usearch_b200_cluster_many",
    "usearch_b200_profile_phases", "usearch_b200_device", "usearch_b200_kernel_launches", "usearch_b200_last_kernel_ms",
    "usearch_b200_bytes_per_vector", "usearch_b200_max_level", "usearch_b200_add_many", "usearch_b200_add_many_device",
    "usearch_b200_shards_unique_id", "usearch_b200_shards_join", "usearch_b200_sharded_search_many",
    "usearch_b200_sharded_search_many_device", "usearch_b200_shards_payload_bytes", "usearch_b200_merge_topk",
    "usearch_b200_search_many_enqueue", "usearch_b200_search_many_finish", "usearch_b200_tune",
]


def load_library() -> C.CDLL:
    """Load the CUDA library. Fails loudly: there is no pure-Python or CPU implementation behind it."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). The B200 backend has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    err = C.POINTER(C.c_char_p)
    lib.usearch_version.restype = C.c_char_p
    lib.usearch_init.restype = C.c_void_p
    lib.usearch_init.argtypes = [C.POINTER(_InitOptions), err]
    lib.usearch_free.argtypes = [C.c_void_p, err]
    lib.usearch_hardware_acceleration.restype = C.c_char_p
    lib.usearch_hardware_acceleration.argtypes = [C.c_void_p, err]
    for name in ("usearch_memory_usage", "usearch_serialized_length", "usearch_size", "usearch_capacity",
                 "usearch_dimensions", "usearch_connectivity", "usearch_expansion_add", "usearch_expansion_search"):
        getattr(lib, name).restype = C.c_size_t
        getattr(lib, name).argtypes = [C.c_void_p, err]
    for name in ("usearch_change_expansion_add", "usearch_change_expansion_search"):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_size_t, err]
    for name in ("usearch_save", "usearch_load", "usearch_view"):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_char_p, err]
    for name in ("usearch_save_buffer", "usearch_load_buffer", "usearch_view_buffer"):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, err]
    lib.usearch_metadata_buffer.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_InitOptions), err]
    lib.usearch_metadata.argtypes = [C.c_char_p, C.POINTER(_InitOptions), err]
    lib.usearch_clear.argtypes = [C.c_void_p, err]
    lib.usearch_add.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, err]
    lib.usearch_search.restype = C.c_size_t
    lib.usearch_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, err]
    lib.usearch_search_many.restype = C.c_size_t
    lib.usearch_search_many.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t,
                                        C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, err]
    lib.usearch_b200_search_many_stats.restype = C.c_size_t
    lib.usearch_b200_search_many_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int,
                                                   C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, err]
    lib.usearch_b200_search_many_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, err]
    lib.usearch_b200_search_many_enqueue.argtypes = lib.usearch_b200_search_many_device.argtypes
    lib.usearch_b200_search_many_finish.argtypes = [C.c_void_p, err]
    lib.usearch_b200_tune.restype = C.c_int
    lib.usearch_b200_tune.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.usearch_b200_filtered_search_many.restype = C.c_size_t
    lib.usearch_b200_filtered_search_many.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int,
                                                      C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                                      C.c_void_p, C.c_void_p, C.c_void_p, err]
    lib.usearch_b200_cluster_many.restype = None
    lib.usearch_b200_cluster_many.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, err]
    lib.usearch_b200_exact_search_many.restype = C.c_size_t
    lib.usearch_b200_exact_search_many.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t,
                                                   C.c_void_p, C.c_void_p, C.c_void_p, err]
    lib.usearch_exact_search.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int,
                                         C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                         C.c_size_t, err]
    lib.usearch_b200_profile_phases.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.usearch_b200_device.argtypes = [C.c_void_p]
    lib.usearch_b200_kernel_launches.restype = C.c_uint64
    lib.usearch_b200_kernel_launches.argtypes = [C.c_void_p]
    lib.usearch_b200_last_kernel_ms.restype = C.c_float
    lib.usearch_b200_last_kernel_ms.argtypes = [C.c_void_p]
    lib.usearch_b200_bytes_per_vector.restype = C.c_size_t
    lib.usearch_b200_bytes_per_vector.argtypes = [C.c_void_p]
    lib.usearch_b200_max_level.restype = C.c_size_t
    lib.usearch_b200_max_level.argtypes = [C.c_void_p]
    lib.usearch_reserve.argtypes = [C.c_void_p, C.c_size_t, err]
    for name in ("usearch_b200_add_many", "usearch_b200_add_many_device"):
        getattr(lib, name).restype = None
        getattr(lib, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, err]
    lib.usearch_contains.restype = C.c_bool
    lib.usearch_contains.argtypes = [C.c_void_p, C.c_uint64, err]
    lib.usearch_count.restype = C.c_size_t
    lib.usearch_count.argtypes = [C.c_void_p, C.c_uint64, err]
    lib.usearch_get.restype = C.c_size_t
    lib.usearch_get.argtypes = [C.c_void_p, C.c_uint64, C.c_size_t, C.c_void_p, C.c_int, err]
    lib.usearch_remove.restype = C.c_size_t
    lib.usearch_remove.argtypes = [C.c_void_p, C.c_uint64, err]
    lib.usearch_rename.restype = C.c_size_t
    lib.usearch_rename.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, err]
    lib.usearch_b200_shards_unique_id.argtypes = [C.c_void_p, err]
    lib.usearch_b200_shards_join.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, err]
    lib.usearch_b200_sharded_search_many.restype = C.c_size_t
    lib.usearch_b200_sharded_search_many.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t,
                                                     C.c_void_p, C.c_void_p, C.c_void_p, err]
    lib.usearch_b200_sharded_search_many_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                            C.c_void_p, err]
    lib.usearch_b200_shards_payload_bytes.restype = C.c_size_t
    lib.usearch_b200_shards_payload_bytes.argtypes = [C.c_size_t, C.c_size_t]
    lib.usearch_b200_merge_topk.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, err]
    lib.usearch_distance.restype = C.c_float
    lib.usearch_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int, err]
    _lib = lib
    return lib


def _raise(err: C.c_char_p) -> None:
    if err.value:
        raise RuntimeError(err.value.decode())


@dataclass
class Matches:
    """Single-query result (index.py:300-330): ``keys`` and ``distances`` trimmed to the found count."""
    keys: np.ndarray
    distances: np.ndarray
    visited_members: int = 0
    computed_distances: int = 0

    def __len__(self) -> int:
        return len(self.keys)

    def to_list(self):
        return [(int(k), float(d)) for k, d in zip(self.keys, self.distances)]


@dataclass
class BatchMatches:
    """Batch result (index.py:333-396): dense ``[nq, count]`` matrices + per-row ``counts``."""
    keys: np.ndarray
    distances: np.ndarray
    counts: np.ndarray
    visited_members: int = 0
    computed_distances: int = 0

    def __len__(self) -> int:
        return len(self.counts)

    def __getitem__(self, i: int) -> Matches:
        n = int(self.counts[i])
        return Matches(self.keys[i, :n], self.distances[i, :n])

    def to_list(self):
        return [self[i].to_list() for i in range(len(self))]

    def mean_recall(self, expected: np.ndarray, count: Optional[int] = None) -> float:
        return float(self.count_matches(expected, count)) / len(expected)

    def count_matches(self, expected: np.ndarray, count: Optional[int] = None) -> int:
        """index.py:379-393: is ``expected[i]`` anywhere among the first ``count`` results of row i."""
        hits = 0
        for i in range(len(expected)):
            n = int(self.counts[i]) if count is None else min(count, int(self.counts[i]))
            hits += int(expected[i] in self.keys[i, :n])
        return hits


class Index:
    """Drop-in for ``usearch.index.Index`` whose graph lives in B200 HBM.

    ``add`` links batches of new members into the graph on the GPU; ``load``/``view``/``restore`` take a ``.usearch``
    file built anywhere; ``search`` runs the whole batch as one persistent-kernel launch.
    """

    def __init__(self, *, ndim: int = 0, metric: str = "cos", dtype: str = "f32", connectivity: int = 16,
                 expansion_add: int = 128, expansion_search: int = 64, multi: bool = False,
                 path: Optional[str] = None, view: bool = False):
        self._lib = load_library()
        err = C.c_char_p()
        if ndim:
            opts = _InitOptions(METRIC_KIND[metric], None, SCALAR_KIND[dtype], ndim, connectivity, expansion_add,
                                expansion_search, multi)
            self._h = C.c_void_p(self._lib.usearch_init(C.byref(opts), C.byref(err)))
        else:
            self._h = C.c_void_p(self._lib.usearch_init(None, C.byref(err)))
        _raise(err)
        self._dtype = dtype
        self._expansion_search = expansion_search
        self._keepalive = None
        if path is not None:
            (self.view if view else self.load)(path)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.usearch_free(self._h, None)
            self._h = None

    # ---- loading (index.py:1100-1200 `load`/`view`/`restore`) ------------------------------------
    @staticmethod
    def restore(path_or_buffer, view: bool = False) -> "Index":
        index = Index()
        (index.view if view else index.load)(path_or_buffer)
        return index

    def load(self, path_or_buffer: Union[str, os.PathLike, bytes, bytearray, np.ndarray]) -> "Index":
        err = C.c_char_p()
        if isinstance(path_or_buffer, (str, os.PathLike)):
            self._lib.usearch_load(self._h, os.fspath(path_or_buffer).encode(), C.byref(err))
            meta = self.metadata(path_or_buffer)
        else:
            buf = np.frombuffer(path_or_buffer, dtype=np.uint8) if not isinstance(path_or_buffer, np.ndarray) \
                else np.ascontiguousarray(path_or_buffer, dtype=np.uint8)
            self._lib.usearch_load_buffer(self._h, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(err))
            meta = self.metadata(buf) if not err.value else None
        _raise(err)
        self._dtype = meta["dtype"]
        self._lib.usearch_change_expansion_search(self._h, self._expansion_search, None)
        return self

    view = load  # the device copy never aliases the file: `view` == `load`

    @staticmethod
    def metadata(path_or_buffer) -> dict:
        lib = load_library()
        opts = _InitOptions()
        err = C.c_char_p()
        if isinstance(path_or_buffer, (str, os.PathLike)):
            lib.usearch_metadata(os.fspath(path_or_buffer).encode(), C.byref(opts), C.byref(err))
        else:
            buf = np.ascontiguousarray(path_or_buffer, dtype=np.uint8)
            lib.usearch_metadata_buffer(buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(opts), C.byref(err))
        _raise(err)
        inv_m = {v: k for k, v in METRIC_KIND.items()}
        inv_s = {v: k for k, v in SCALAR_KIND.items()}
        return {"metric": inv_m.get(opts.metric_kind), "dtype": inv_s.get(opts.quantization),
                "ndim": opts.dimensions, "multi": bool(opts.multi)}

    def save(self, path: Optional[str] = None) -> Optional[np.ndarray]:
        err = C.c_char_p()
        if path is not None:
            self._lib.usearch_save(self._h, os.fspath(path).encode(), C.byref(err))
            _raise(err)
            return None
        n = self._lib.usearch_serialized_length(self._h, None)
        buf = np.empty(n, dtype=np.uint8)
        self._lib.usearch_save_buffer(self._h, buf.ctypes.data_as(C.c_void_p), n, C.byref(err))
        _raise(err)
        return buf

    # ---- properties (index.py:1377-1470) -----------------------------------------------------------
    size = property(lambda s: s._lib.usearch_size(s._h, None))
    ndim = property(lambda s: s._lib.usearch_dimensions(s._h, None))
    connectivity = property(lambda s: s._lib.usearch_connectivity(s._h, None))
    capacity = property(lambda s: s._lib.usearch_capacity(s._h, None))
    memory_usage = property(lambda s: s._lib.usearch_memory_usage(s._h, None))
    serialized_length = property(lambda s: s._lib.usearch_serialized_length(s._h, None))
    hardware_acceleration = property(lambda s: s._lib.usearch_hardware_acceleration(s._h, None).decode())
    dtype = property(lambda s: s._dtype)
    max_level = property(lambda s: s._lib.usearch_b200_max_level(s._h))
    kernel_launches = property(lambda s: s._lib.usearch_b200_kernel_launches(s._h))
    last_kernel_ms = property(lambda s: s._lib.usearch_b200_last_kernel_ms(s._h))
    bytes_per_vector = property(lambda s: s._lib.usearch_b200_bytes_per_vector(s._h))

    def __len__(self) -> int:
        return self.size

    @property
    def expansion_search(self) -> int:
        return self._lib.usearch_expansion_search(self._h, None)

    @expansion_search.setter
    def expansion_search(self, v: int) -> None:
        self._expansion_search = v
        self._lib.usearch_change_expansion_search(self._h, v, None)

    @property
    def expansion_add(self) -> int:
        return self._lib.usearch_expansion_add(self._h, None)

    @expansion_add.setter
    def expansion_add(self, v: int) -> None:
        self._lib.usearch_change_expansion_add(self._h, v, None)

    # ---- mutation (index.py:560-700 `add`, :800-900 `remove`/`rename`/`get`/`contains`/`count`) -------------
    def reserve(self, capacity: int) -> None:
        err = C.c_char_p()
        self._lib.usearch_reserve(self._h, int(capacity), C.byref(err))
        _raise(err)

    def add(self, keys, vectors: np.ndarray, *, copy: bool = True, threads: int = 0, log=False, progress=None):
        """`Index.add` (index.py:560-640): one key + vector, or a batch. The batch is linked into the graph on the
        GPU (builder.cu); `keys=None` numbers the rows from the current size, as the reference does. Returns the keys."""
        vectors = np.asarray(vectors)
        if vectors.ndim == 1:
            vectors = vectors[None, :]
        if not vectors.flags.c_contiguous and vectors.strides[1] != vectors.itemsize:
            vectors = np.ascontiguousarray(vectors)
        n = vectors.shape[0]
        if keys is None:
            start = len(self)
            keys = np.arange(start, start + n, dtype=np.uint64)
        keys = np.ascontiguousarray(np.atleast_1d(np.asarray(keys)), dtype=np.uint64)
        if keys.shape[0] != n:
            raise ValueError("The number of keys must match the number of vectors")
        kind = self._kind_of(vectors)
        err = C.c_char_p()
        self._lib.usearch_b200_add_many(self._h, keys.ctypes.data_as(C.c_void_p), vectors.ctypes.data_as(C.c_void_p), n,
                                        vectors.strides[0], SCALAR_KIND[kind], C.byref(err))
        _raise(err)
        return keys

    def add_device(self, keys_ptr: int, vectors_ptr: int, n: int, stride: int, kind: Optional[str] = None) -> None:
        """Batch add from DEVICE memory (raw pointers): no host round trip for the vectors."""
        err = C.c_char_p()
        self._lib.usearch_b200_add_many_device(self._h, keys_ptr, vectors_ptr, n, stride,
                                               SCALAR_KIND[kind or self._dtype], C.byref(err))
        _raise(err)

    def contains(self, key: int) -> bool:
        return bool(self._lib.usearch_contains(self._h, int(key), None))

    __contains__ = contains

    def count(self, key: int) -> int:
        return int(self._lib.usearch_count(self._h, int(key), None))

    def get(self, key: int, dtype: Optional[str] = None, count: int = 1) -> Optional[np.ndarray]:
        """`Index.get` (index.py:820-870): the vector(s) stored under `key`, or None."""
        kind = dtype or self._dtype
        np_t = {"f32": np.float32, "f64": np.float64, "f16": np.float16, "bf16": np.uint16, "i8": np.int8, "b1": np.uint8}[kind]
        cols = (self.ndim + 7) // 8 if kind == "b1" else self.ndim
        out = np.zeros((count, cols), dtype=np_t)
        err = C.c_char_p()
        found = self._lib.usearch_get(self._h, int(key), count, out.ctypes.data_as(C.c_void_p), SCALAR_KIND[kind], C.byref(err))
        _raise(err)
        if not found:
            return None
        return out[0] if count == 1 else out[:found]

    def remove(self, key: int) -> int:
        err = C.c_char_p()
        n = self._lib.usearch_remove(self._h, int(key), C.byref(err))
        _raise(err)
        return int(n)

    def rename(self, key_from: int, key_to: int) -> int:
        err = C.c_char_p()
        n = self._lib.usearch_rename(self._h, int(key_from), int(key_to), C.byref(err))
        _raise(err)
        return int(n)

    def clear(self) -> None:
        self._lib.usearch_clear(self._h, None)

    # ---- search (index.py:700-748) ---------------------------------------------------------------
    def _kind_of(self, vectors: np.ndarray) -> str:
        if vectors.dtype == np.uint16:
            return "bf16"
        kind = _NP_TO_SCALAR.get(vectors.dtype)
        if kind is None:
            raise TypeError(f"Unsupported query dtype {vectors.dtype}")
        if kind == "b1" and self._dtype == "i8":
            return "i8"
        return kind

    def filtered_search(self, vectors: np.ndarray, count: int, allowed_keys) -> Union[Matches, BatchMatches]:
        """`filtered_search` (index_dense.hpp:774-779) for the predicate "key in allowed_keys"."""
        return self.search(vectors, count, stats=True, _allowed=np.ascontiguousarray(allowed_keys, dtype=np.uint64))

    def cluster(self, vectors: np.ndarray, level: int = 1, *, stats: bool = False):
        """`index_dense_gt::cluster(vector, level)` (index_dense.hpp:788-793; index.hpp:3092-3125) for every row: the
        closest member on graph level `level` (levels above the top return the entry point; 0 behaves like 1).
        Returns `(keys [nq] u64, distances [nq] f32)`; with `stats=True` the counters land in `last_computed` /
        `last_visited`."""
        vectors = np.asarray(vectors)
        if vectors.ndim == 1:
            vectors = vectors[None, :]
        if not vectors.flags.c_contiguous and vectors.strides[1] != vectors.itemsize:
            vectors = np.ascontiguousarray(vectors)
        kind = self._kind_of(vectors)
        nq = vectors.shape[0]
        keys = np.zeros(nq, dtype=np.uint64)
        distances = np.zeros(nq, dtype=np.float32)
        computed = np.zeros(nq, dtype=np.uint64)
        visited = np.zeros(nq, dtype=np.uint64)
        err = C.c_char_p()
        self._lib.usearch_b200_cluster_many(
            self._h, vectors.ctypes.data_as(C.c_void_p), nq, vectors.strides[0], SCALAR_KIND[kind], int(level),
            keys.ctypes.data_as(C.c_void_p), distances.ctypes.data_as(C.c_void_p),
            computed.ctypes.data_as(C.c_void_p) if stats else None, visited.ctypes.data_as(C.c_void_p) if stats else None,
            C.byref(err))
        _raise(err)
        if stats:
            self.last_computed, self.last_visited = computed, visited
        return keys, distances

    def search(self, vectors: np.ndarray, count: int = 10, *, stats: bool = False, threads: int = 0, exact: bool = False,
               log=False, progress=None, _allowed: Optional[np.ndarray] = None) -> Union[Matches, BatchMatches]:
        """1-D input → :class:`Matches`; 2-D input → :class:`BatchMatches` (index.py:191-231).

        `threads`, `log` and `progress` are accepted for signature compatibility with index.py:700-748 and
        ignored (one kernel launch serves the whole batch); `exact=True` brute-forces every member on the GPU."""
        vectors = np.asarray(vectors)
        single = vectors.ndim == 1
        if single:
            vectors = vectors[None, :]
        if vectors.ndim != 2:
            raise ValueError("Expects a matrix or a vector")
        if not vectors.flags.c_contiguous and vectors.strides[1] != vectors.itemsize:
            vectors = np.ascontiguousarray(vectors)  # rows must be contiguous (python/lib.cpp:426-428)
        kind = self._kind_of(vectors)
        expect_cols = (self.ndim * _BITS[kind] + 7) // 8 // vectors.itemsize if kind != "b1" else (self.ndim + 7) // 8
        if vectors.shape[1] != expect_cols:
            raise ValueError("The number of columns must match the dimensionality of the index")
        nq = vectors.shape[0]
        keys = np.zeros((nq, count), dtype=np.uint64)
        distances = np.zeros((nq, count), dtype=np.float32)
        counts = np.zeros(nq, dtype=np.uint64)
        err = C.c_char_p()
        vm = cd = 0
        if exact:
            self._lib.usearch_b200_exact_search_many(
                self._h, vectors.ctypes.data_as(C.c_void_p), nq, vectors.strides[0], SCALAR_KIND[kind], count,
                keys.ctypes.data_as(C.c_void_p), distances.ctypes.data_as(C.c_void_p),
                counts.ctypes.data_as(C.c_void_p), C.byref(err))
            _raise(err)
        elif _allowed is not None:
            computed = np.zeros(nq, dtype=np.uint64)
            visited = np.zeros(nq, dtype=np.uint64)
            self._lib.usearch_b200_filtered_search_many(
                self._h, vectors.ctypes.data_as(C.c_void_p), nq, vectors.strides[0], SCALAR_KIND[kind], count,
                _allowed.ctypes.data_as(C.c_void_p), _allowed.size, keys.ctypes.data_as(C.c_void_p),
                distances.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p),
                computed.ctypes.data_as(C.c_void_p), visited.ctypes.data_as(C.c_void_p), C.byref(err))
            _raise(err)
            self.last_computed, self.last_visited = computed, visited
            vm, cd = int(visited.sum()), int(computed.sum())
        elif stats:
            computed = np.zeros(nq, dtype=np.uint64)
            visited = np.zeros(nq, dtype=np.uint64)
            self._lib.usearch_b200_search_many_stats(
                self._h, vectors.ctypes.data_as(C.c_void_p), nq, vectors.strides[0], SCALAR_KIND[kind], count,
                keys.ctypes.data_as(C.c_void_p), distances.ctypes.data_as(C.c_void_p),
                counts.ctypes.data_as(C.c_void_p), computed.ctypes.data_as(C.c_void_p),
                visited.ctypes.data_as(C.c_void_p), C.byref(err))
            _raise(err)
            self.last_computed, self.last_visited = computed, visited
            vm, cd = int(visited.sum()), int(computed.sum())
        else:
            self._lib.usearch_search_many(
                self._h, vectors.ctypes.data_as(C.c_void_p), nq, vectors.strides[0], SCALAR_KIND[kind], count,
                keys.ctypes.data_as(C.c_void_p), keys.strides[0], distances.ctypes.data_as(C.c_void_p),
                distances.strides[0], counts.ctypes.data_as(C.c_void_p), C.byref(err))
            _raise(err)
        if single:
            n = int(counts[0])
            return Matches(keys[0, :n], distances[0, :n], vm, cd)
        return BatchMatches(keys, distances, counts, vm, cd)

    def tune(self, **knobs: int) -> None:
        """Launch tuning knobs of this handle (stage_sets, warps_per_sm); results never change."""
        for name, value in knobs.items():
            if self._lib.usearch_b200_tune(self._h, name.encode(), int(value)) != 0:
                raise ValueError(f"unknown knob {name}")

    def profile_phases(self, enable: bool = True) -> dict:
        """Read (then reset) the kernel's per-phase cycle counters; see include/usearch_b200.h."""
        out = np.zeros(16, dtype=np.uint64)
        self._lib.usearch_b200_profile_phases(self._h, int(enable), out.ctypes.data_as(C.c_void_p))
        names = ["setup_descent", "heap_pop", "row_visited", "vector_wait", "distance_math", "accept", "output"]
        q = max(int(out[7]), 1)
        return {"queries": int(out[7]), **{n: float(out[i]) / q for i, n in enumerate(names)},
                "pushes": float(out[8]) / q, "avg_max_heap": float(out[9]) / q, "max_heap": int(out[10])}

    # ---- sharded search: this index is one shard of a group of processes (shards.cu) ------------------------
    def join_shards(self, rank: int, world: int, unique_id: bytes) -> None:
        """Collective. `unique_id` = the 128 bytes rank 0 got from :func:`shards_unique_id`, handed to every rank."""
        err = C.c_char_p()
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        self._lib.usearch_b200_shards_join(self._h, rank, world, buf, C.byref(err))
        _raise(err)

    def sharded_search(self, vectors: np.ndarray, count: int = 10) -> BatchMatches:
        """Collective: every rank passes the same queries and receives the merged top-`count` of all shards."""
        vectors = np.ascontiguousarray(vectors)
        if vectors.ndim == 1:
            vectors = vectors[None, :]
        kind = self._kind_of(vectors)
        nq = vectors.shape[0]
        keys = np.zeros((nq, count), dtype=np.uint64)
        distances = np.zeros((nq, count), dtype=np.float32)
        counts = np.zeros(nq, dtype=np.uint64)
        err = C.c_char_p()
        self._lib.usearch_b200_sharded_search_many(self._h, vectors.ctypes.data_as(C.c_void_p), nq, vectors.strides[0],
                                                   SCALAR_KIND[kind], count, keys.ctypes.data_as(C.c_void_p),
                                                   distances.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p),
                                                   C.byref(err))
        _raise(err)
        return BatchMatches(keys, distances, counts)

    def sharded_search_device(self, queries_ptr: int, nq: int, stride: int, count: int, keys_ptr: int, distances_ptr: int,
                              counts_ptr: int, computed_ptr: int = 0, visited_ptr: int = 0, stream: int = 0) -> None:
        err = C.c_char_p()
        self._lib.usearch_b200_sharded_search_many_device(self._h, queries_ptr, nq, stride, count, keys_ptr, distances_ptr,
                                                          counts_ptr, computed_ptr or None, visited_ptr or None,
                                                          stream or None, C.byref(err))
        _raise(err)

    def search_enqueue(self, queries_ptr: int, nq: int, stride: int, count: int, keys_ptr: int, distances_ptr: int,
                       counts_ptr: int, computed_ptr: int = 0, visited_ptr: int = 0, stream: int = 0) -> None:
        """Like :meth:`search_device` but returns as soon as the kernel is enqueued; call :meth:`search_finish` once."""
        err = C.c_char_p()
        self._lib.usearch_b200_search_many_enqueue(self._h, queries_ptr, nq, stride, count, keys_ptr, distances_ptr,
                                                   counts_ptr, computed_ptr or None, visited_ptr or None,
                                                   stream or None, C.byref(err))
        _raise(err)

    def search_finish(self) -> None:
        err = C.c_char_p()
        self._lib.usearch_b200_search_many_finish(self._h, C.byref(err))
        _raise(err)

    def search_device(self, queries_ptr: int, nq: int, stride: int, count: int, keys_ptr: int, distances_ptr: int,
                      counts_ptr: int, computed_ptr: int = 0, visited_ptr: int = 0, stream: int = 0) -> None:
        """Device-resident batch: raw device pointers (e.g. ``tensor.data_ptr()``) and a CUDA stream handle."""
        err = C.c_char_p()
        self._lib.usearch_b200_search_many_device(self._h, queries_ptr, nq, stride, count, keys_ptr, distances_ptr,
                                                  counts_ptr, computed_ptr or None, visited_ptr or None,
                                                  stream or None, C.byref(err))
        _raise(err)


def exact_search(dataset: np.ndarray, queries: np.ndarray, count: int = 10, *, metric: str = "cos",
                 dtype: Optional[str] = None) -> BatchMatches:
    """Brute-force many-to-many search over raw matrices: `usearch.index.search(dataset, query, count, metric,
    exact=True)` (python/usearch/index.py) -> `usearch_exact_search` (c/lib.cpp:468-501). Keys are dataset rows."""
    lib = load_library()
    dataset = np.ascontiguousarray(dataset)
    queries = np.ascontiguousarray(queries)
    if queries.ndim == 1:
        queries = queries[None, :]
    kind = dtype or ("bf16" if dataset.dtype == np.uint16 else _NP_TO_SCALAR[dataset.dtype])
    dims = dataset.shape[1] * 8 if kind == "b1" else dataset.shape[1]
    nq = queries.shape[0]
    keys = np.zeros((nq, count), dtype=np.uint64)
    distances = np.zeros((nq, count), dtype=np.float32)
    err = C.c_char_p()
    lib.usearch_exact_search(dataset.ctypes.data_as(C.c_void_p), dataset.shape[0], dataset.strides[0],
                             queries.ctypes.data_as(C.c_void_p), nq, queries.strides[0], SCALAR_KIND[kind], dims,
                             METRIC_KIND[metric], count, 0, keys.ctypes.data_as(C.c_void_p), keys.strides[0],
                             distances.ctypes.data_as(C.c_void_p), distances.strides[0], C.byref(err))
    _raise(err)
    return BatchMatches(keys, distances, np.full(nq, count, dtype=np.uint64))


def shards_unique_id() -> bytes:
    """The 128-byte group id (an ncclUniqueId) rank 0 creates; every rank passes it to `Index.join_shards`."""
    lib = load_library()
    buf = (C.c_char * 128)()
    err = C.c_char_p()
    lib.usearch_b200_shards_unique_id(buf, C.byref(err))
    _raise(err)
    return bytes(buf)


def merge_topk(shard_results, count: int) -> BatchMatches:
    """The merge kernel on its own: `shard_results` = per-shard (keys [nq,count] u64, distances [nq,count] f32, counts [nq]),
    as `Indexes` would merge them (python/lib.cpp:350-391) but ordered by (distance, shard, position)."""
    lib = load_library()
    world = len(shard_results)
    nq = shard_results[0][0].shape[0]
    size = lib.usearch_b200_shards_payload_bytes(nq, count)
    blob = np.zeros(world * size, dtype=np.uint8)
    for r, (k, d, c) in enumerate(shard_results):
        base = r * size
        blob[base:base + nq * count * 8] = np.ascontiguousarray(k, dtype=np.uint64).view(np.uint8).ravel()
        blob[base + nq * count * 8:base + nq * count * 12] = np.ascontiguousarray(d, dtype=np.float32).view(np.uint8).ravel()
        blob[base + nq * count * 12:base + nq * count * 12 + nq * 4] = np.ascontiguousarray(c).astype(np.uint32).view(np.uint8)
    keys = np.zeros((nq, count), dtype=np.uint64)
    distances = np.zeros((nq, count), dtype=np.float32)
    counts = np.zeros(nq, dtype=np.uint32)
    err = C.c_char_p()
    lib.usearch_b200_merge_topk(blob.ctypes.data_as(C.c_void_p), world, nq, count, keys.ctypes.data_as(C.c_void_p),
                                distances.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p), C.byref(err))
    _raise(err)
    return BatchMatches(keys, distances, counts.astype(np.uint64))
