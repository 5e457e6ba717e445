#!/usr/bin/env bash
# round 2, run O: branch-light walk steps (sphere leaves by selects, deferred normals)
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_o
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "default      $(q many_spheres 3)"
for v in p6 p6r12 p6r16 p5; do echo "$v         $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/lib$v.so q many_spheres 3)"; done
echo "default      $(q hanging_lamps_in_row 1 0 640 360)"
echo "generic      $(ACN_SPECIALIZE=0 q wine_glass 3)"
} > gpurun_out/quick_r2o.log 2>&1
cat gpurun_out/quick_r2o.log
timeout 1700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2o.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2o.log
tail -5 gpurun_out/pytest_gpu_r2o.log
ncu --set full --clock-control none --import-source on -k regex:'k_direct|k_path' -s 0 -c 4 -o gpurun_out/prof_r2o_spheres python tools/quick_bench.py many_spheres 1 > gpurun_out/ncu_r2o.log 2>&1; echo "ncu rc $?"
rm -rf gpurun_out/spec_cache_o
