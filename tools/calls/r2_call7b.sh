#!/bin/bash
# Round 2, GPU call 7b (1 GPU): full suite, tcgen05 exact kernel with the branch-free filter.
O=gpurun_out/r2c7b; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_suite.log 2>&1; echo "rc=$?" >> $O/gpu_suite.log
for M in ip cos l2sq; do for K in imma umma; do
  USEARCH_B200_EXACT=$K timeout 300 python tools/exact_bench.py 1000000 1024 4096 i8 $M > $O/exact_bench_${K}_$M.log 2>&1
done; done
USEARCH_B200_EXACT=umma timeout 300 python tools/exact_bench.py 4000000 768 8192 i8 ip > $O/exact_bench_umma_ip_4M.log 2>&1
tail -n 4 $O/gpu_suite.log; head -qn 1 $O/exact_bench_*.log
