"""Sharded search through the library's NCCL path against the CPU `Indexes`-style merge, one process per GPU:

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/shard_check.py

Rank r builds shard r (keys congruent to r) with the REFERENCE on the host, loads it into its GPU, joins the group
(usearch_b200_shards_join) and calls the collective usearch_b200_sharded_search_many. Every rank must receive the rows the
CPU reference produces when it searches the same serialised shards and merges by (distance, shard, position)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from usearch_b200 import sharded  # noqa: E402
from usearch_b200.index import Index  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ["USEARCH_B200_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for metric, scalar, d, m in (("cos", "f32", 96, 16), ("ip", "i8", 256, 16), ("hamming", "b1", 128, 16)):
        n, k, ef, nq = 12000, 10, 64, 300
        base, queries = common.make_collection(n, d, scalar, nq)
        keys = np.arange(n, dtype=np.uint64)
        ref, blob = common.build_reference_blob(base[rank::world], metric, scalar, d, m, threads=8, keys=keys[rank::world])
        ref.pin_metric(True)
        ref.change_expansion_search(ef)
        mine = ref.search(queries, k, threads=8)
        index = Index.restore(blob)
        index.expansion_search = ef
        sharded.join(index)
        got = index.sharded_search(queries, k)
        # device-pointer variant
        q_dev = torch.from_numpy(queries.view(np.uint8).reshape(nq, -1)).cuda()
        vs = (q_dev.shape[1] + 15) // 16 * 16
        q_pad = torch.zeros((nq, vs), dtype=torch.uint8, device="cuda")
        q_pad[:, :q_dev.shape[1]] = q_dev
        kd = torch.zeros((nq, k), dtype=torch.int64, device="cuda")
        dd = torch.zeros((nq, k), dtype=torch.float32, device="cuda")
        cd = torch.zeros(nq, dtype=torch.int32, device="cuda")
        index.sharded_search_device(q_pad.data_ptr(), nq, vs, k, kd.data_ptr(), dd.data_ptr(), cd.data_ptr(),
                                    stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        box = [None] * world
        dist.all_gather_object(box, (mine[0], mine[1], mine[2]))
        want_k = np.zeros((nq, k), np.uint64)
        want_d = np.full((nq, k), np.array(0x7FA00000, dtype=np.uint32).view(np.float32), np.float32)
        for q in range(nq):
            items = []
            for r, (kk, ddd, cc) in enumerate(box):
                items += [(float(ddd[q, i]), r, i, int(kk[q, i])) for i in range(int(cc[q]))]
            items.sort(key=lambda t: (t[0], t[1], t[2]))
            for i, (dv, _, _, kv) in enumerate(items[:k]):
                want_k[q, i], want_d[q, i] = kv, dv
        same = (np.array_equal(got.keys, want_k) and np.array_equal(got.distances.view(np.uint32), want_d.view(np.uint32)) and
                np.array_equal(kd.cpu().numpy().astype(np.uint64), want_k) and
                np.array_equal(dd.cpu().numpy().view(np.uint32), want_d.view(np.uint32)))
        print(f"rank {rank} {metric}/{scalar}: sharded search == CPU merge of the same shards: {same}", flush=True)
        ok = ok and same
        del index
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARD_CHECK_OK" if int(flag.item()) == 1 else "SHARD_CHECK_FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
