#!/bin/bash
# Round 2, last GPU call: the GPU suite and smoke on the final build.
O=gpurun_out/r2c14; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/gpu_suite.log 2>&1; echo "rc=$?" >> $O/gpu_suite.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
tail -n 3 $O/gpu_suite.log; tail -n 3 $O/smoke.log
