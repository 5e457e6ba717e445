#!/usr/bin/env bash
# round 2, run Z: segmented sums of multi-task blocks by match + warp-reduce instead of the shuffle ladder
mkdir -p gpurun_out
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
export ACN_SPECIALIZE=1
{
echo "ladder       $(ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_z0 ACN_SPEC_OPTS=-DACN_NO_MATCH_SUM=1 q wine_glass 3)"
echo "match        $(ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_z1 q wine_glass 3)"
echo "ladder       $(ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_z0 ACN_SPEC_OPTS=-DACN_NO_MATCH_SUM=1 q diamond 3)"
echo "match        $(ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_z1 q diamond 3)"
echo "match        $(ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_z1 q primitives 3)"
echo "match generic $(ACN_SPECIALIZE=0 q hanging_lamps_in_row 1 0 640 360)"
} > gpurun_out/quick_r2z.log 2>&1
cat gpurun_out/quick_r2z.log
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_z1
unset ACN_SPECIALIZE
timeout 1700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2z.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2z.log
tail -5 gpurun_out/pytest_gpu_r2z.log
rm -rf gpurun_out/spec_cache_z0 gpurun_out/spec_cache_z1
