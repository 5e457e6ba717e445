#!/usr/bin/env bash
# round 2, run U: new walk tests, k_direct at 6 blocks/SM, C1 probe (shell width), full suite
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_u
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
timeout 900 python -m pytest tests/test_gpu_walk.py -m gpu -x -q -s -p no:cacheprovider > gpurun_out/pytest_gpu_r2u_walk.log 2>&1; echo "walk tests rc $?"
tail -25 gpurun_out/pytest_gpu_r2u_walk.log | cut -c1-220
{
echo "default      $(q many_spheres 3)"
echo "p6           $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libp6.so q many_spheres 3)"
} > gpurun_out/quick_r2u.log 2>&1
cat gpurun_out/quick_r2u.log
timeout 900 python tools/c1_probe.py > gpurun_out/c1_probe_r2u.log 2>&1; echo "probe rc $?"
cat gpurun_out/c1_probe_r2u.log | cut -c1-700
rm -rf gpurun_out/spec_cache_u
