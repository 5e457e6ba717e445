#!/usr/bin/env bash
# round 2, run 3a: k_path of the group kernels — blocks that span several tasks add per lane instead of the shuffle ladder
mkdir -p gpurun_out
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
export ACN_SPECIALIZE=1
{
echo "ladder       $(ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_a0 q wine_glass 3)"
echo "lane atomics $(ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_a1 ACN_SPEC_OPTS=-DACN_PATH_LANE_ATOMICS=1 q wine_glass 3)"
echo "ladder       $(ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_a0 q diamond 3)"
echo "lane atomics $(ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_a1 ACN_SPEC_OPTS=-DACN_PATH_LANE_ATOMICS=1 q diamond 3)"
echo "ladder       $(ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_a0 q primitives 3)"
echo "lane atomics $(ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_a1 ACN_SPEC_OPTS=-DACN_PATH_LANE_ATOMICS=1 q primitives 3)"
} > gpurun_out/quick_r3a.log 2>&1
cat gpurun_out/quick_r3a.log
rm -rf gpurun_out/spec_cache_a0 gpurun_out/spec_cache_a1
