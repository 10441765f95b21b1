#!/bin/bash
# Round 2, seventh GPU call (2 GPUs): sharded search through NCCL in the library, bench at N = 2 in both layouts.
O=gpurun_out/r2c7; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29555 tools/shard_check.py > $O/shard_check.log 2>&1; echo "rc=$?" >> $O/shard_check.log
timeout 900 $TR --master-port 29556 bench.py --gpus 2 --workload C2 --steps 10 --warmup 3 > $O/bench_c2_replica_2gpu.json 2> $O/bench_c2_replica_2gpu.err; echo "rc=$?" >> $O/bench_c2_replica_2gpu.err
timeout 900 $TR --master-port 29557 bench.py --gpus 2 --workload C4 --steps 10 --warmup 3 > $O/bench_c4_shard_2gpu.json 2> $O/bench_c4_shard_2gpu.err; echo "rc=$?" >> $O/bench_c4_shard_2gpu.err
timeout 900 $TR --master-port 29558 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_ns_replica_2gpu.json 2> $O/bench_ns_replica_2gpu.err; echo "rc=$?" >> $O/bench_ns_replica_2gpu.err
tail -n 6 $O/shard_check.log; for f in c2_replica c4_shard ns_replica; do tail -n 2 $O/bench_${f}_2gpu.err; cut -c1-1200 $O/bench_${f}_2gpu.json; done
