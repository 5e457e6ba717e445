#!/usr/bin/env bash
# 8-GPU box: the bench line (weak scaling headline + extras: strong scaling, C3/C4/C1/C5 split over the ranks, video frames)
# at N = 8, 4, 2, and two whole images through the multi-GPU pass controller
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_scale
PORT=29511
run() { n=$1; shift; PORT=$((PORT+1)); timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT "$@"; }
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for n in 8 4 2; do
  run $n bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/scale_r2_n${n}.err | tail -1 > gpurun_out/scale_r2_n${n}.json; echo "N=$n rc $?"
done
python bench.py --gpus 1 --steps 5 --warmup 3 2> gpurun_out/scale_r2_n1.err | tail -1 > gpurun_out/scale_r2_n1.json; echo "N=1 rc $?"
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open(f'gpurun_out/scale_r2_n{n}.json').read().strip().splitlines()[-1])
        ex=d.get('extras') or {}
        print('N',d['n_gpus'],'value %.4g'%d['value'],'ms/step %.2f'%d['ms_per_step'],'e2e %.4g'%d['e2e']['value'],'rays/s %.4g'%d['rays_per_sec'], 'strong', (ex.get('strong') or {}).get('ms_per_step'), {k:round(v['ms_per_step'],2) for k,v in (ex.get('configs') or {}).items()}, 'video fps', (ex.get('video') or {}).get('frames_per_sec'))
    except Exception as e: print(n,'parse fail',e)
PY
# whole images (all passes of the scripted adaptive controller), one process: acn_group over 8 GPUs vs 1 GPU
for s in primitives wine_glass; do for g in 8 1; do timeout 300 python tools/render.py --scene $s --gpus $g --out gpurun_out/${s}_g$g.pnm 2>&1 | grep -v Warning | tail -1; done; done
rm -rf gpurun_out/spec_cache_scale gpurun_out/*.pnm
