"""CPU test of the N>1 path: two gloo ranks all-gather their per-shard top-k and merge deterministically."""
import os
import subprocess
import sys

import common


def test_two_rank_gloo_merge():
    worker = os.path.join(common.ROOT, "tests", "dist_merge_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", worker]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MERGE_OK 2" in out.stdout


def test_merge_single_process_semantics():
    import numpy as np
    import torch
    from usearch_b200.sharded import merge_gathered
    # two shards, a tie at distance 1.0: shard 0 wins; padding never surfaces before real entries
    d0 = torch.tensor([[1.0, 2.0, float("nan")]])
    d1 = torch.tensor([[1.0, 1.5, 3.0]])
    k0 = torch.tensor([[10, 12, 0]])
    k1 = torch.tensor([[11, 13, 15]])
    keys, dist_, counts = merge_gathered([k0, k1], [d0, d1], [torch.tensor([2]), torch.tensor([3])], 4)
    assert keys.tolist() == [[10, 11, 13, 12]] and counts.tolist() == [4]
    assert dist_.tolist() == [[1.0, 1.0, 1.5, 2.0]]
    keys, dist_, counts = merge_gathered([k0, k1], [d0, d1], [torch.tensor([1]), torch.tensor([0])], 3)
    assert counts.tolist() == [1] and keys.tolist() == [[10, 0, 0]] and np.isnan(dist_.numpy()[0, 1:]).all()


def test_merge_orders_infinities_nans_and_short_inputs():
    """a valid +inf / NaN distance must not lose its place to padding; k may exceed world * columns."""
    import numpy as np
    import torch
    from usearch_b200.sharded import merge_gathered
    inf, nan = float("inf"), float("nan")
    d0 = torch.tensor([[1.0, nan]])          # one valid entry, one padding
    d1 = torch.tensor([[inf, nan]])          # a valid +inf, then a valid NaN
    k0 = torch.tensor([[10, 0]])
    k1 = torch.tensor([[11, 12]])
    keys, dist_, counts = merge_gathered([k0, k1], [d0, d1], [torch.tensor([1]), torch.tensor([2])], 3)
    assert counts.tolist() == [3] and keys.tolist() == [[10, 11, 12]]
    assert dist_[0, 0] == 1.0 and dist_[0, 1] == inf and np.isnan(dist_[0, 2].item())
    keys, dist_, counts = merge_gathered([k0, k1], [d0, d1], [torch.tensor([1]), torch.tensor([1])], 6)   # k > world * cols
    assert keys.shape == (1, 6) and counts.tolist() == [2] and keys.tolist()[0][:2] == [10, 11] and keys.tolist()[0][2:] == [0] * 4
