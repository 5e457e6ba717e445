#!/usr/bin/env bash
# round 2, run X: ncu of the scene-specialised kernels on wine_glass, the BIG launches of the pass (the r2c capture had caught a 89 us one)
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_x
export ACN_SPECIALIZE=1
python tools/quick_bench.py wine_glass 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_direct|k_path|k_rays|k_shade' -s 8 -c 8 -o gpurun_out/prof_r2x_wine python tools/quick_bench.py wine_glass 1 > gpurun_out/ncu_r2x.log 2>&1; echo "ncu rc $?"
ls -la gpurun_out/spec_cache_x | head
