#!/bin/bash
# Round 2, first GPU call: golden-test replay, opt-in tests, experimental kernel variants on f16 / i8.
mkdir -p gpurun_out/r2c1
O=gpurun_out/r2c1
timeout 300 python tools/repro_golden_next_rows.py > $O/repro.log 2>&1; echo "repro rc=$?" >> $O/repro.log
USEARCH_B200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_zz_gpu_golden.py -x -q > $O/golden.log 2>&1; echo "rc=$?" >> $O/golden.log
F16="--dtype f16 --ef 256 --batch 16384 --steps 5 --warmup 3 --no-cpu-baseline"
timeout 400 python bench.py $F16 > $O/f16_default.json 2> $O/f16_default.err
USEARCH_B200_HALF_WORDS=1 timeout 300 python bench.py $F16 > $O/f16_halfw.json 2> $O/f16_halfw.err
USEARCH_B200_HALF_WORDS=1 USEARCH_B200_STAGED_DENSE=1 timeout 300 python bench.py $F16 > $O/f16_halfw_dense.json 2> $O/f16_halfw_dense.err
USEARCH_B200_HALF_WORDS=1 USEARCH_B200_STAGED_DENSE=1 USEARCH_B200_STAGE_SETS=1 timeout 300 python bench.py $F16 > $O/f16_halfw_dense_1set.json 2> $O/f16_halfw_dense_1set.err
I8="--dtype i8 --metric ip --dim 1024 --connectivity 16 --batch 16384 --steps 5 --warmup 3 --no-cpu-baseline"
timeout 400 python bench.py $I8 > $O/i8_default.json 2> $O/i8_default.err
USEARCH_B200_STAGED_DENSE=1 timeout 300 python bench.py $I8 > $O/i8_dense.json 2> $O/i8_dense.err
USEARCH_B200_STAGED_DENSE=1 USEARCH_B200_STAGE_SETS=1 timeout 300 python bench.py $I8 > $O/i8_dense_1set.json 2> $O/i8_dense_1set.err
tail -3 $O/repro.log $O/golden.log; cat $O/*.json | cut -c1-400
