#!/usr/bin/env bash
# ncu full capture only (3 kernels). $1 = tag, $2 = scene, $3 = skip
TAG=${1:-r1d}; SCENE=${2:-wine_glass}; SKIP=${3:-45}
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --scene $SCENE > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'k_direct|k_path|k_rays' -s $SKIP -c 3 \
    -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --scene $SCENE > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc $?"
