#!/usr/bin/env bash
# GPU iteration: parity tests, then a short bench of the scene given as $1 (default wine_glass)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 --scene ${1:-wine_glass} ${2:-} > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench rc $?"
tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_iter.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'rays/s',d['rays_per_sec'],'launches',d['gpu_launches'])
    print('kernel ms',r['kernel_ms_per_step']); print('frac',r.get('frac'),'whole',r.get('whole_step_frac')); print('cpu',d.get('cpu_baseline'))
except Exception as e: print('parse fail',e)
PY
