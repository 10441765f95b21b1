#!/bin/bash
# Round 2, GPU call 13 (1 GPU): visited set as an L2-resident hash table instead of a DRAM-resident bitmap at 10M slots.
O=gpurun_out/r2c13; mkdir -p $O
USEARCH_B200_VISITED=hash timeout 600 python tools/sweep.py --workload C4 --phases --configs base > $O/sweep_c4_hash.jsonl 2> $O/sweep_c4_hash.err
USEARCH_B200_VISITED=hash timeout 600 python tools/sweep.py --workload NS --phases --configs base > $O/sweep_ns_hash.jsonl 2> $O/sweep_ns_hash.err
cat $O/sweep_c4_hash.jsonl $O/sweep_ns_hash.jsonl | cut -c1-700; tail -n 2 $O/*.err
