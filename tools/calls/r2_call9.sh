#!/bin/bash
# Round 2, GPU call 9 (8 GPUs): the two sharded BASELINE configs at full size, sharded parity at world = 8.
O=gpurun_out/r2c9; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29565 tools/shard_check.py > $O/shard_check.log 2>&1; echo "rc=$?" >> $O/shard_check.log
timeout 600 $TR --master-port 29566 bench.py --gpus 8 --workload C4 --steps 10 --warmup 3 > $O/bench_c4_shard_8gpu.json 2> $O/bench_c4_shard_8gpu.err; echo "rc=$?" >> $O/bench_c4_shard_8gpu.err
timeout 900 $TR --master-port 29567 bench.py --gpus 8 --workload C5 --steps 10 --warmup 3 > $O/bench_c5_shard_8gpu.json 2> $O/bench_c5_shard_8gpu.err; echo "rc=$?" >> $O/bench_c5_shard_8gpu.err
tail -n 3 $O/shard_check.log; for f in c4 c5; do tail -n 3 $O/bench_${f}_shard_8gpu.err; cut -c1-1500 $O/bench_${f}_shard_8gpu.json; done
