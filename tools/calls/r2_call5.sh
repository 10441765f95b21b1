#!/bin/bash
# Round 2, fifth GPU call: FFMA2 metrics + cp.async staging + unrolled log cleaning; b1 24-warp kernel.
O=gpurun_out/r2c5; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/gpu_suite.log 2>&1; echo "rc=$?" >> $O/gpu_suite.log
USEARCH_B200_STAGE_COPY=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_golden.py -q -x > $O/gpu_suite_ldgsts.log 2>&1; echo "rc=$?" >> $O/gpu_suite_ldgsts.log
timeout 600 python tools/sweep.py --workload NS --phases --configs "base;stage_copy=1;stage_copy=1,warps_per_sm=3" > $O/sweep_ns.jsonl 2> $O/sweep_ns.err
timeout 600 python tools/sweep.py --workload C3 --steps 4 --configs "base;stage_copy=1;stage_copy=1,stage_sets=2" > $O/sweep_c3.jsonl 2> $O/sweep_c3.err
timeout 600 python tools/sweep.py --workload C4 --configs "base;stage_copy=1;stage_copy=1,stage_sets=2" > $O/sweep_c4.jsonl 2> $O/sweep_c4.err
timeout 900 python tools/sweep.py --workload C5 --n 12500000 --steps 4 --phases --configs "base;dense_direct=1" > $O/sweep_c5.jsonl 2> $O/sweep_c5.err
tail -n 6 $O/gpu_suite.log; tail -n 4 $O/gpu_suite_ldgsts.log; for f in ns c3 c4 c5; do echo == $f; tail -n 2 $O/sweep_$f.err; cut -c1-700 $O/sweep_$f.jsonl; done
