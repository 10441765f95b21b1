#!/bin/bash
O=gpurun_out/r2c15; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_build.py tests/test_native_clients.py tests/test_gpu_parity.py -m gpu -q > $O/gpu_subset.log 2>&1; echo "rc=$?" >> $O/gpu_subset.log
tail -n 3 $O/gpu_subset.log
