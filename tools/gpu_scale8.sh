#!/usr/bin/env bash
# 8-GPU box: scaling bench (wine_glass at N=8,4,2; many_spheres and diamond at N=8) and multi-GPU renders
mkdir -p gpurun_out
PORT=29511
run() { n=$1; shift; PORT=$((PORT+1)); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT "$@"; }
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for n in 8 4 2; do
  run $n bench.py --gpus $n --steps 3 --warmup 3 2> gpurun_out/scale_n${n}_wine_glass.err | tail -1 > gpurun_out/scale_n${n}_wine_glass.json; echo "N=$n wine_glass rc $?"
done
for s in many_spheres diamond; do
  run 8 bench.py --gpus 8 --steps 3 --warmup 3 --scene $s 2> gpurun_out/scale_n8_$s.err | tail -1 > gpurun_out/scale_n8_$s.json; echo "N=8 $s rc $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/scale_n*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'N',d['n_gpus'],'value %.3g'%d['value'],'ms/step %.2f'%d['ms_per_step'],'e2e %.3g'%d['e2e']['value'],'rays/s %.3g'%d['rays_per_sec'])
    except Exception as e: print(f,'parse fail',e)
PY
run 8 tools/render.py --scene primitives --out gpurun_out/primitives_n8.pnm 2>&1 | grep -v Warning | tail -2
run 8 tools/render.py --frames diamond_video_000000 diamond_video_000010 diamond_video_000020 diamond_video_000030 diamond_video_000040 diamond_video_000050 diamond_video_000060 diamond_video_000070 --passes 2 --out gpurun_out/vid8 2>&1 | grep -v Warning | tail -9
run 8 tools/render.py --scene hanging_lamps_in_row --passes 1 --out gpurun_out/lamps_n8.pnm 2>&1 | grep -v Warning | tail -2
ls -la gpurun_out/*.pnm | tail -12
