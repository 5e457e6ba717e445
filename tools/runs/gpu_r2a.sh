#!/usr/bin/env bash
# round 2, run A: CPU-oracle full renders (background), whole GPU test suite, timings of the config scenes, bench line
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt; nvidia-smi -L >> gpurun_out/nproc.txt
( python tools/parity_report.py --cpu-full primitives diamond many_spheres wine_glass > gpurun_out/cpu_full.log 2>&1 ) &
CPU_PID=$!
timeout 1500 python -m pytest tests -m gpu -x -q -s -p no:cacheprovider > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2a.log
for s in wine_glass diamond many_spheres primitives; do timeout 300 python tools/quick_bench.py $s 5 2>&1 | tail -1; done > gpurun_out/quick_r2a.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
wait $CPU_PID
python tools/parity_report.py --merge >> gpurun_out/cpu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu_r2a.log; cat gpurun_out/quick_r2a.log; tail -3 gpurun_out/cpu_full.log
