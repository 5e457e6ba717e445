#!/usr/bin/env bash
# both bench arms at N=1 (the driver's round-end sequence), then smoke()
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc $?"; tail -c 600 gpurun_out/bench_ref.json
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc $?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e'],'rays/s',d['rays_per_sec'],'launches',d['gpu_launches'])
print('kernel ms',r['kernel_ms_per_step']); print('roof',{k:r.get(k) for k in ('kernel','achieved','peak','frac','whole_step_frac')}); print('cpu',d.get('cpu_baseline')); print('clocks',d['clocks'])
PY
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc $?"
