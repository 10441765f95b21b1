#!/bin/bash
# Round 2, GPU call 7a (1 GPU): staging loop without divisions, tcgen05 epilogue with shared-memory lists; ncu of the tcgen05 kernel.
O=gpurun_out/r2c7a; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/gpu_suite.log 2>&1; echo "rc=$?" >> $O/gpu_suite.log
for K in imma umma; do
  USEARCH_B200_EXACT=$K timeout 300 python tools/exact_bench.py 1000000 1024 4096 i8 ip > $O/exact_bench_$K.log 2>&1; echo "rc=$?" >> $O/exact_bench_$K.log
done
USEARCH_B200_EXACT=umma timeout 300 python tools/exact_bench.py 1000000 768 4096 i8 cos > $O/exact_bench_umma_cos.log 2>&1
USEARCH_B200_EXACT=imma timeout 300 python tools/exact_bench.py 1000000 768 4096 i8 cos > $O/exact_bench_imma_cos.log 2>&1
USEARCH_B200_EXACT=umma timeout 600 ncu --set full --import-source on --clock-control none -k regex:exact_umma -c 1 -f -o $O/umma_i8 python tools/exact_bench.py 1000000 1024 4096 i8 ip > $O/ncu_umma.log 2>&1; echo "rc=$?" >> $O/ncu_umma.log
timeout 600 python tools/sweep.py --workload NS --phases --configs "base" > $O/sweep_ns.jsonl 2> $O/sweep_ns.err
timeout 600 python tools/sweep.py --workload C3 --steps 4 --configs "base" > $O/sweep_c3.jsonl 2> $O/sweep_c3.err
timeout 600 python tools/sweep.py --workload C4 --configs "base" > $O/sweep_c4.jsonl 2> $O/sweep_c4.err
tail -n 4 $O/gpu_suite.log; tail -n 3 $O/exact_bench_*.log; tail -n 2 $O/ncu_umma.log; for f in ns c3 c4; do cut -c1-500 $O/sweep_$f.jsonl; done
