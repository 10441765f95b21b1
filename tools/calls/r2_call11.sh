#!/bin/bash
# Round 2, GPU call 11 (1 GPU): shard emulation (debug of the N=8 recall figure), full-size bench lines with parity for C3 / C4 / C5-shaped.
O=gpurun_out/r2c11; mkdir -p $O
timeout 300 python tools/shard_emulate.py 8 3 C4 > $O/emulate_c4_8.json 2> $O/emulate_c4_8.err
timeout 300 python tools/shard_emulate.py 2 1 C4 > $O/emulate_c4_2.json 2> $O/emulate_c4_2.err
timeout 900 python bench.py --workload C3 --steps 5 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "rc=$?" >> $O/bench_c3.err
timeout 600 python bench.py --workload C4 --steps 10 --warmup 3 > $O/bench_c4_1gpu.json 2> $O/bench_c4_1gpu.err; echo "rc=$?" >> $O/bench_c4_1gpu.err
timeout 600 python bench.py --workload C5 --n 12500000 --steps 5 --warmup 3 > $O/bench_c5_12m.json 2> $O/bench_c5_12m.err; echo "rc=$?" >> $O/bench_c5_12m.err
timeout 600 python tools/sweep.py --workload C5 --n 12500000 --ef 128 --steps 3 --configs base > $O/sweep_c5_ef128.jsonl 2> $O/sweep_c5_ef128.err
cat $O/emulate_c4_8.json $O/emulate_c4_2.json; tail -n 2 $O/emulate_c4_8.err; for f in c3 c4_1gpu c5_12m; do tail -n 2 $O/bench_$f.err | cut -c1-300; cut -c1-1800 $O/bench_$f.json; done; cat $O/sweep_c5_ef128.jsonl
