#!/bin/bash
# Round 2, second GPU call: the batched builder (correctness, quality vs a reference build, timing), golden tests, GPU suite.
O=gpurun_out/r2c2; mkdir -p $O
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/build_check.py --only-tiny > $O/memcheck.log 2>&1; echo "rc=$?" >> $O/memcheck.log
timeout 1200 python tools/build_check.py --big > $O/build_check.jsonl 2> $O/build_check.err; echo "rc=$?" >> $O/build_check.err
timeout 600 python -m pytest tests/test_zz_gpu_golden.py -x -q > $O/golden.log 2>&1; echo "rc=$?" >> $O/golden.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_zz_gpu_golden.py > $O/gpu_suite.log 2>&1; echo "rc=$?" >> $O/gpu_suite.log
tail -n 5 $O/memcheck.log; tail -n 3 $O/build_check.err; cut -c1-600 $O/build_check.jsonl; tail -n 5 $O/golden.log; tail -n 8 $O/gpu_suite.log
