#!/bin/bash
O=gpurun_out/r2c10; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29571 bench.py --gpus 4 --workload C4 --steps 3 --warmup 3 > $O/bench_c4_shard_4gpu.json 2> $O/bench_c4_shard_4gpu.err; echo "rc=$?" >> $O/bench_c4_shard_4gpu.err
grep "recall@" $O/bench_c4_shard_4gpu.err | cut -c1-400; cut -c1-600 $O/bench_c4_shard_4gpu.json
