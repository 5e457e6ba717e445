#!/usr/bin/env bash
# round 2, run I: HEAD after the revert — full GPU suite, bench line, per-scene timings, launch list
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_i
q() { timeout 900 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
timeout 1700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2i.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2i.log
python bench.py > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; echo "bench rc $?"
{
for s in wine_glass diamond primitives; do echo "spec    $(ACN_SPECIALIZE=1 q $s 3)"; done
echo "generic $(q many_spheres 3)"
echo "generic $(q hanging_lamps_in_row 1 0 640 360)"
} > gpurun_out/quick_r2i.log 2>&1
cat gpurun_out/quick_r2i.log; tail -4 gpurun_out/pytest_gpu_r2i.log; tail -c 1500 gpurun_out/bench_r2i.json
