#!/usr/bin/env bash
# round 2, run 3e: matter records front to back per octant of ray directions (many_spheres)
mkdir -p gpurun_out
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "scene order  $(ACN_NO_OCTANT_ORDER=1 q many_spheres 3)"
echo "octants      $(q many_spheres 3)"
} > gpurun_out/quick_r3e.log 2>&1
cat gpurun_out/quick_r3e.log
ACN_VERBOSE=1 python tools/quick_bench.py many_spheres 1 0 60 60 2>&1 | grep "traversal records"
timeout 1200 python -m pytest tests/test_gpu_scripted.py tests/test_gpu_walk.py tests/test_gpu_configs.py -m gpu -x -q -k "many_spheres or walk or tight" -p no:cacheprovider 2>&1 | tail -3
