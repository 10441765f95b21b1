"""Host logic of bench.py on the CPU (torch CPU tensors stand in for the device): the synthetic collection is a pure function
of (seed, chunk), shards partition it, the ground truth and the two recall definitions do what they say."""
import os
import sys

import numpy as np
import pytest
import torch

import common

sys.path.insert(0, common.ROOT)
import bench  # noqa: E402


def _args(monkeypatch, *extra):
    monkeypatch.setattr(sys, "argv", ["bench.py", *extra])
    return bench.parse_args()


def test_named_workloads_are_the_baseline_configs(monkeypatch):
    a = _args(monkeypatch)
    assert (a.n, a.dim, a.dtype, a.metric, a.connectivity, a.ef, a.batch) == (10_000_000, 768, "f32", "cos", 32, 128, 4096)
    assert "10000000x768 f32 cos" in bench.workload_name(a) and a.parallelism == "replica"
    c5 = _args(monkeypatch, "--workload", "C5")
    assert (c5.n, c5.dim, c5.dtype, c5.metric, c5.connectivity, c5.batch, c5.parallelism) == (100_000_000, 256, "b1", "hamming", 64, 32768, "shard")
    dev = _args(monkeypatch, "--workload", "C3", "--n", "5000", "--batch", "64")
    assert (dev.n, dev.dtype, dev.ef, dev.batch) == (5000, "f16", 256, 64)


@pytest.mark.parametrize("dtype,metric", [("f32", "cos"), ("f16", "cos"), ("bf16", "ip"), ("i8", "ip"), ("b1", "hamming")])
def test_collection_is_deterministic_and_shards_partition_it(monkeypatch, dtype, metric):
    monkeypatch.setattr(bench, "CHUNK", 256)
    a = _args(monkeypatch, "--n", "1000", "--dim", "64", "--dtype", dtype, "--metric", metric)
    coll = bench.Collection(a, torch.device("cpu"))
    whole = [(ids.clone(), x.clone()) for ids, x in coll.base_chunks()]
    again = [(ids, x) for ids, x in bench.Collection(a, torch.device("cpu")).base_chunks()]
    assert all(torch.equal(i0, i1) and torch.equal(x0.view(torch.uint8), x1.view(torch.uint8)) for (i0, x0), (i1, x1) in zip(whole, again))
    all_ids = torch.cat([i for i, _ in whole])
    assert torch.equal(all_ids, torch.arange(1000))
    rows = {int(i): x[j].view(torch.uint8).numpy().tobytes() for i_, x in whole for j, i in enumerate(i_)}
    for shards in (2, 3, 8):
        seen = {}
        for shard in range(shards):
            for ids, x in coll.base_chunks(shard, shards):
                assert (ids % shards == shard).all()
                for j, i in enumerate(ids.tolist()):
                    assert i not in seen
                    seen[i] = x[j].view(torch.uint8).numpy().tobytes()
        assert seen == rows
    q = coll.queries(300)
    assert q.shape[0] == 300 and torch.equal(q.view(torch.uint8), coll.queries(300).view(torch.uint8))
    assert not torch.equal(q.view(torch.uint8), coll.queries(300, stream=1).view(torch.uint8))
    host = bench.to_numpy_queries(a, q)
    assert host.dtype == {"f32": np.float32, "f16": np.float16, "bf16": np.uint16, "i8": np.int8, "b1": np.uint8}[dtype]
    assert host.shape[0] == 300 and host.strides[0] == (64 * {"f32": 32, "f16": 16, "bf16": 16, "i8": 8, "b1": 1}[dtype]) // 8


def test_quantisation_matches_the_host_recipe(monkeypatch):
    """The device-side quantisation of bench.py and usearch_b200.datagen.to_scalar (what the tests feed the reference) agree."""
    from usearch_b200 import datagen
    a = _args(monkeypatch, "--n", "100", "--dim", "64")
    x = torch.randn((50, 64), generator=torch.Generator().manual_seed(3))
    for dtype in ("f16", "bf16", "i8", "b1"):
        a.dtype = dtype
        got = bench.Collection(a, torch.device("cpu")).quantise(x)
        want = datagen.to_scalar(x.numpy(), dtype)
        got_np = got.view(torch.uint16).numpy() if dtype == "bf16" else got.numpy()
        assert np.array_equal(got_np.view(np.uint8), want.view(np.uint8)), dtype


@pytest.mark.parametrize("dtype,metric", [("f32", "cos"), ("f32", "l2sq"), ("i8", "ip"), ("b1", "hamming")])
def test_ground_truth_is_the_brute_force_answer(monkeypatch, dtype, metric):
    monkeypatch.setattr(bench, "CHUNK", 128)
    a = _args(monkeypatch, "--n", "700", "--dim", "64", "--dtype", dtype, "--metric", metric)
    coll = bench.Collection(a, torch.device("cpu"))
    q = coll.queries(20)
    gt_k, gt_d = bench.exact_topk_gpu(a, coll, q, 10)
    base = torch.cat([x for _, x in coll.base_chunks()])
    xf, qf = coll.as_float(base).double().numpy(), coll.as_float(q).double().numpy()
    if metric == "cos":
        d = 1 - (qf / np.linalg.norm(qf, axis=1, keepdims=True)) @ (xf / np.linalg.norm(xf, axis=1, keepdims=True)).T
    elif metric == "ip":
        d = 1 - qf @ xf.T
    elif metric == "hamming":
        d = (64 - qf @ xf.T) / 2
    else:
        d = ((qf[:, None, :] - xf[None, :, :]) ** 2).sum(2)
    want = np.sort(d, axis=1)[:, :10]
    got = np.take_along_axis(d, gt_k.numpy(), axis=1)            # the true distances of the reported ids
    assert np.allclose(np.sort(got, axis=1), want, rtol=1e-4, atol=1e-4)
    if metric == "hamming":                                      # integers: the reported distances are exact
        assert np.array_equal(np.sort(gt_d.numpy(), axis=1), want.astype(np.float32))


def test_recall_definitions():
    truth = np.array([[1, 2, 3, 4], [5, 6, 7, 8]], dtype=np.uint64)
    found = np.array([[1, 2, 9, 4], [5, 6, 0, 0]], dtype=np.uint64)
    assert bench.recall_at_k(found, np.array([4, 2]), truth) == pytest.approx(5 / 8)
    # ties: the third entry of row 0 is another member at the k-th distance -> counts; the padding of row 1 does not
    truth_d = np.array([[1, 2, 3, 3], [1, 1, 2, 2]], dtype=np.float32)
    found_d = np.array([[1, 2, 3, 3], [1, 1, np.nan, np.nan]], dtype=np.float32)
    assert bench.recall_at_k_with_ties(found_d, np.array([4, 2]), truth_d) == pytest.approx(6 / 8)
    worse = np.array([[1, 2, 3, 4], [1, 1, 2, 3]], dtype=np.float32)
    assert bench.recall_at_k_with_ties(worse, np.array([4, 4]), truth_d) == pytest.approx(6 / 8)
