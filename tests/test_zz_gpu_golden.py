"""GPU parity against the COMMITTED golden vectors (tests/golden/*.npz: outputs of the unmodified reference, made by
make_golden.py / make_golden_next_rows.py): nothing of the reference or the oracle is involved at run time.
Collected last (file name) so that every other GPU test has reported before these start."""
import glob
import os

import numpy as np
import pytest

import common
from oracle import bindings

pytestmark = pytest.mark.gpu

GOLDEN_FIXTURES = sorted(glob.glob(os.path.join(common.GOLDEN, "*_n*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLDEN_FIXTURES]


# Opt-in like the rest of this file: none of these has been seen green on hardware yet (see the note below). The same
# comparison - GPU results of `usearch_search` against the fixture's reference outputs - is part of the default suite
# through the C client (tests/test_native_clients.py, fixture cos_f32_n2000_d64).
@pytest.mark.skipif(os.environ.get("USEARCH_B200_TEST_EXPERIMENTAL") != "1", reason="opt-in: USEARCH_B200_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("path", GOLDEN_FIXTURES, ids=IDS)
def test_search_matches_golden(path):
    """Labels, distance bits, counts and both counters of the graph search."""
    from usearch_b200.index import Index
    g = np.load(path)
    index = Index.restore(g["blob"])
    index.expansion_search = int(g["ef"])
    got = index.search(g["queries"], int(g["k"]), stats=True)
    want = (g["keys_pinned"], g["distances_pinned"], g["counts_pinned"], g["computed_pinned"], g["visited_pinned"])
    common.assert_same_results(want, (got.keys, got.distances, got.counts, index.last_computed, index.last_visited), "gpu vs golden")
    # the reference's native SimSIMD dispatch on the generating host: same labels
    assert np.array_equal(got.keys, g["keys_native"])


# Written after the round's GPU budget was spent: the one attempt to run it ended in the job's time limit before this
# test reported (cause not established - see DESIGN.md §8 "open items"), so it stays opt-in until it has been seen green.
@pytest.mark.skipif(os.environ.get("USEARCH_B200_TEST_EXPERIMENTAL") != "1", reason="opt-in: USEARCH_B200_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("path", GOLDEN_FIXTURES, ids=IDS)
def test_next_rows_match_golden(path):
    """Exact search (index mode and free function) and cluster against tests/golden/next_rows.npz."""
    from usearch_b200 import v2format
    from usearch_b200.index import Index, exact_search
    name = os.path.basename(path)[:-4]
    g, nr = np.load(path), np.load(os.path.join(common.GOLDEN, "next_rows.npz"))
    q, k = g["queries"], int(g["k"])
    index = Index.restore(g["blob"])
    got = index.search(q, k, exact=True)
    assert np.array_equal(got.keys, nr[f"{name}/exact_keys"]) and np.array_equal(got.counts, nr[f"{name}/exact_counts"])
    assert np.array_equal(got.distances.view(np.uint32), nr[f"{name}/exact_distances"].view(np.uint32))
    for i, level in enumerate(nr[f"{name}/cluster_levels"]):
        ck, cd = index.cluster(q, int(level), stats=True)
        assert np.array_equal(ck, nr[f"{name}/cluster_keys"][i]), f"level {level}"
        assert np.array_equal(cd.view(np.uint32), nr[f"{name}/cluster_distances"][i].view(np.uint32))
        assert np.array_equal(index.last_computed, nr[f"{name}/cluster_computed"][i])
        assert np.array_equal(index.last_visited, nr[f"{name}/cluster_visited"][i])
    graph = v2format.loads(g["blob"])
    vectors = graph.vectors.view(bindings.SCALAR_NP[graph.scalar]).reshape(graph.size, -1)
    free = exact_search(vectors, q, k, metric=graph.metric, dtype=graph.scalar)
    wd, wk = nr[f"{name}/free_distances"], nr[f"{name}/free_keys"]
    assert np.array_equal(free.distances.view(np.uint32), wd[:, :k].view(np.uint32))
    unique = wd[:, :k] != wd[:, 1:k + 1]
    unique[:, 1:] &= wd[:, 1:k] != wd[:, :k - 1]
    assert np.array_equal(free.keys[unique], wk[:, :k][unique])


@pytest.mark.parametrize("flags,metric,scalar,n,d,m,ef,k,nq", [
    ("HALF_WORDS", "cos", "f16", 6000, 768, 32, 128, 10, 128),
    ("HALF_WORDS", "l2sq", "f16", 4000, 200, 16, 64, 10, 128),     # 400-byte vectors: ragged last chunk group
    ("HALF_WORDS", "ip", "bf16", 4000, 256, 16, 64, 10, 128),
    ("HALF_WORDS", "cos", "bf16", 3000, 136, 16, 300, 20, 64),     # ef > 256: shared-memory `top`
    ("STAGED_DENSE", "ip", "i8", 8000, 1024, 16, 128, 10, 256),    # 16 resident warps per SM
    ("STAGED_DENSE", "cos", "i8", 4000, 256, 16, 64, 10, 128),
    ("STAGED_DENSE,STAGE_SETS=1", "l2sq", "i8", 4000, 512, 16, 64, 10, 128),
    ("HALF_WORDS,STAGED_DENSE", "cos", "f16", 6000, 768, 32, 256, 10, 128),
])
@pytest.mark.skipif(os.environ.get("USEARCH_B200_TEST_EXPERIMENTAL") != "1",
                    reason="opt-in (USEARCH_B200_TEST_EXPERIMENTAL=1): these variants have not run on hardware yet")
def test_experimental_variants_match_reference(flags, metric, scalar, n, d, m, ef, k, nq):
    """The off-by-default kernel variants (USEARCH_B200_HALF_WORDS: f16/bf16 with 4 lanes per vector split by
    accumulator; USEARCH_B200_STAGED_DENSE: 16 resident warps per SM) against the reference. The switches are read once
    per process, hence the subprocess."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np, common\n"
        "from oracle import bindings\n"
        "from usearch_b200.index import Index\n"
        "base, q = common.make_collection(%d, %d, %r, %d)\n"
        "ref, blob = common.build_reference_blob(base, %r, %r, %d, %d, threads=16)\n"
        "want = bindings.PortIndex(blob, %d).search(q, %d, threads=16)\n"
        "index = Index.restore(blob); index.expansion_search = %d\n"
        "got = index.search(q, %d, stats=True)\n"
        "common.assert_same_results(want, (got.keys, got.distances, got.counts, index.last_computed, index.last_visited), 'variant')\n"
        "print('VARIANT_OK')\n"
    ) % (common.ROOT, os.path.join(common.ROOT, "tests"), n, d, scalar, nq, metric, scalar, d, m, ef, k, ef, k)
    env = dict(os.environ)
    for flag in flags.split(","):
        name, _, value = flag.partition("=")
        env["USEARCH_B200_" + name] = value or "1"
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "VARIANT_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
