#!/usr/bin/env bash
# round 2, run 3l: final state after the 6-block k_direct and the group records — whole GPU suite, smoke, bench line
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_l3
timeout 1700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r3l.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r3l.log
tail -4 gpurun_out/pytest_gpu_r3l.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_r3l.json 2> gpurun_out/bench_r3l.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r3l.json').read().strip().splitlines()[-1])
print('value %.4g ms %.2f e2e %.4g frac %.4f whole %.4f cpu %.4g launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['cpu_baseline']['value'], d['gpu_launches']))
print({k: round(v['ms_per_step'], 2) for k, v in d['extras']['configs'].items()}, d['clocks'])
PY
rm -rf gpurun_out/spec_cache_l3
