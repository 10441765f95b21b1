#!/bin/bash
O=gpurun_out/r2c17; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_build.py tests/test_native_clients.py tests/test_zz_gpu_golden.py -m gpu -q > $O/gpu_subset.log 2>&1; echo "rc=$?" >> $O/gpu_subset.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
tail -n 3 $O/gpu_subset.log; tail -n 2 $O/smoke.log
