#!/bin/bash
# Round 2, third GPU call: new GPU tests, builder batch-ratio sweep, the metric's own configuration (10M x 768 f32), both arms.
O=gpurun_out/r2c3; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_build.py -q > $O/test_gpu_build.log 2>&1; echo "rc=$?" >> $O/test_gpu_build.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
for R in 16 32 64; do
  USEARCH_B200_BUILD_RATIO=$R timeout 400 python tools/build_check.py --cases latent32,l2_128,l2_100k > $O/ratio_$R.jsonl 2> $O/ratio_$R.err
done
timeout 1500 python bench.py --steps 10 --warmup 3 > $O/bench_ns.json 2> $O/bench_ns.err; echo "rc=$?" >> $O/bench_ns.err
timeout 1200 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ns_ref.json 2> $O/bench_ns_ref.err; echo "rc=$?" >> $O/bench_ns_ref.err
tail -n 6 $O/test_gpu_build.log; tail -n 3 $O/smoke.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c3/ratio_*.jsonl')):
    for l in open(f):
        j=json.loads(l)
        print(f.split('/')[-1], j['case'], 'build', j['build_s_gpu'], 'deg', round(j['structure']['mean_degree0'],1), round(j['structure'].get('reference_mean_degree0',0),1),
              {ef:(v['reference_built']['recall_at_10'],v['gpu_built']['recall_at_10'],v['reference_built']['computed_distances'],v['gpu_built']['computed_distances']) for ef,v in j['ef'].items()})
PY
tail -n 4 $O/bench_ns.err; cut -c1-3000 $O/bench_ns.json; tail -n 3 $O/bench_ns_ref.err; cut -c1-1500 $O/bench_ns_ref.json
