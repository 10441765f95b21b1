#!/bin/bash
# Round 2, sixth GPU call: tcgen05 exact kernel (first run), its timing against mma.sync, ncu on the headline kernel.
O=gpurun_out/r2c6; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_exact.py -q -x -k "i8_exact_kernels or i8" > $O/exact_i8.log 2>&1; echo "rc=$?" >> $O/exact_i8.log
for K in imma umma; do
  USEARCH_B200_EXACT=$K timeout 300 python tools/exact_bench.py 1000000 1024 4096 i8 ip > $O/exact_bench_$K.log 2>&1; echo "rc=$?" >> $O/exact_bench_$K.log
done
USEARCH_B200_EXACT=umma timeout 300 python tools/exact_bench.py 1000000 768 4096 i8 cos > $O/exact_bench_umma_cos.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > $O/gpu_suite.log 2>&1; echo "rc=$?" >> $O/gpu_suite.log
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -c 1 -f -o $O/ns_f32 python tools/sweep.py --workload NS --steps 2 --ncu --configs base > $O/ncu_ns.log 2>&1; echo "rc=$?" >> $O/ncu_ns.log
USEARCH_B200_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_ns.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1; echo "rc=$?" >> $O/bench_under_ncu.log
tail -n 5 $O/exact_i8.log; tail -n 3 $O/exact_bench_*.log; tail -n 4 $O/gpu_suite.log; tail -n 3 $O/ncu_ns.log; tail -n 5 $O/launches_ns.csv; ls -la $O
