#!/usr/bin/env bash
# ncu: launch list + full capture of the tracing kernels. $1 = tag, $2 = scene
TAG=${1:-r1b}; SCENE=${2:-wine_glass}
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --scene $SCENE > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --scene $SCENE > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc $?"
ncu --set full --clock-control none --import-source on -k regex:'k_direct|k_path|k_rays' -s 45 -c 6 \
    -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --scene $SCENE > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc $?"
ls -la gpurun_out | tail -8
