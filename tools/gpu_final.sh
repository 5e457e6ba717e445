#!/usr/bin/env bash
# round-end sequence: both bench arms at N=1, smoke(), ncu launch list + full capture of the tracing kernels, other scenes
TAG=${1:-r1m}
mkdir -p gpurun_out
bash tools/gpu_bench_full.sh
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc $?"
ncu --set full --clock-control none --import-source on -k regex:'k_direct|k_path|k_rays' -s 45 -c 3 \
    -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc $?"
for s in primitives many_spheres diamond; do timeout 200 python tools/quick_bench.py $s 3 2>&1 | tail -1 | cut -c52-; done | tee gpurun_out/scenes_$TAG.log
