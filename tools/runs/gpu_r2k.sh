#!/usr/bin/env bash
# round 2, run K: threaded traversal records (stackless walk) — timings, full GPU suite
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_k
q() { timeout 900 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "generic $(q many_spheres 3)"
echo "generic $(q hanging_lamps_in_row 1 0 640 360)"
echo "generic $(ACN_SPECIALIZE=0 q wine_glass 3)"
echo "spec    $(ACN_SPECIALIZE=1 q wine_glass 3)"
echo "generic $(ACN_SPECIALIZE=0 q primitives 3)"
} > gpurun_out/quick_r2k.log 2>&1
cat gpurun_out/quick_r2k.log
timeout 1700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2k.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2k.log
tail -5 gpurun_out/pytest_gpu_r2k.log
rm -rf gpurun_out/spec_cache_k
