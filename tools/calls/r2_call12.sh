#!/bin/bash
# Round 2, final 1-GPU call: what the driver will run — the GPU suite, smoke, both bench arms on the default workload.
O=gpurun_out/r2c12; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_suite.log 2>&1; echo "rc=$?" >> $O/gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
( time timeout 1200 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err ) 2> $O/time_reference.txt
( time timeout 1200 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err ) 2> $O/time_bench.txt
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -c 1 -f -o $O/c4_i8 python tools/sweep.py --workload C4 --steps 2 --ncu --configs base > $O/ncu_c4.log 2>&1
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -c 1 -f -o $O/c3_f16 python tools/sweep.py --workload C3 --steps 2 --ncu --configs base > $O/ncu_c3.log 2>&1
tail -n 3 $O/gpu_suite.log; tail -n 3 $O/smoke.log; cat $O/time_reference.txt $O/time_bench.txt; cut -c1-700 $O/bench_reference.json; cut -c1-2500 $O/bench.json
