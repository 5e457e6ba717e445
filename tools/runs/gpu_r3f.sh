#!/usr/bin/env bash
# round 2, run 3f: shadow rays test the warp's last occluder first (refill k_direct)
mkdir -p gpurun_out
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "no hint      $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libnohint.so q many_spheres 3)"
echo "hint         $(q many_spheres 3)"
} > gpurun_out/quick_r3f.log 2>&1
cat gpurun_out/quick_r3f.log
timeout 1200 python -m pytest tests/test_gpu_scripted.py tests/test_gpu_walk.py tests/test_gpu_configs.py -m gpu -x -q -k "many_spheres or walk or tight" -p no:cacheprovider 2>&1 | tail -3
