"""GPU parity: the CUDA search (through the C ABI) against the reference CPU search on the SAME
serialised graph, bit-for-bit: labels, distance bits, counts and the reference's two counters."""
import numpy as np
import pytest

import common
from oracle import bindings

pytestmark = pytest.mark.gpu

CASES = [
    # metric, scalar, n, d, M, ef, k, nq
    ("l2sq", "f32", 20000, 128, 16, 64, 10, 512),
    ("cos", "f32", 8000, 768, 32, 128, 10, 256),
    ("ip", "f32", 6000, 96, 16, 64, 10, 256),
    ("cos", "f32", 3000, 97, 13, 32, 7, 128),     # ragged dimension, odd connectivity
    ("ip", "i8", 8000, 1024, 16, 128, 10, 256),
    ("l2sq", "i8", 4000, 100, 16, 64, 10, 128),
    ("cos", "i8", 4000, 256, 16, 64, 10, 128),
    ("hamming", "b1", 20000, 256, 64, 64, 10, 256),
    ("tanimoto", "b1", 6000, 200, 16, 64, 10, 128),
    ("sorensen", "b1", 6000, 256, 16, 64, 10, 128),
]


def _cpu(blob, ref, q, k, ef):
    if ref is not None:
        ref.pin_metric(True)
        ref.change_expansion_search(ef)
        return ref.search(q, k, threads=16)
    port = bindings.PortIndex(blob, ef)
    return port.search(q, k, threads=16)


@pytest.mark.parametrize("metric,scalar,n,d,m,ef,k,nq", CASES)
def test_search_matches_reference(metric, scalar, n, d, m, ef, k, nq):
    from usearch_b200.index import Index
    base, q = common.make_collection(n, d, scalar, nq)
    ref, blob = common.build_reference_blob(base, metric, scalar, d, m, threads=16)
    want = _cpu(blob, ref, q, k, ef)
    port = bindings.PortIndex(blob, ef).search(q, k, threads=16)
    common.assert_same_results(want, port, "port vs reference")
    index = Index.restore(blob)
    index.expansion_search = ef
    assert index.size == n and index.ndim == d and index.connectivity == m
    got = index.search(q, k, stats=True)
    common.assert_same_results(want, (got.keys, got.distances, got.counts, index.last_computed, index.last_visited),
                               f"gpu vs reference [{metric}/{scalar}]")
    assert index.kernel_launches >= 1
