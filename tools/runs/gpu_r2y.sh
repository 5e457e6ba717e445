#!/usr/bin/env bash
# round 2, run Y: experiment — the warps of a block start every group together (instruction-cache sharing), specialised kernels
mkdir -p gpurun_out
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
export ACN_SPECIALIZE=1
{
echo "default      $(ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_y0 q wine_glass 3)"
echo "block phase  $(ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_y1 ACN_SPEC_OPTS=-DACN_BLOCK_PHASE=1 q wine_glass 3)"
echo "default      $(ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_y0 q diamond 3)"
echo "block phase  $(ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_y1 ACN_SPEC_OPTS=-DACN_BLOCK_PHASE=1 q diamond 3)"
} > gpurun_out/quick_r2y.log 2>&1
cat gpurun_out/quick_r2y.log
rm -rf gpurun_out/spec_cache_y0 gpurun_out/spec_cache_y1
