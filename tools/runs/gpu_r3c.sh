#!/usr/bin/env bash
# round 2, run 3c: traversal records loaded with one 256-bit load (LDG.E.256) instead of two 128-bit ones
mkdir -p gpurun_out
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "ld128        $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libld128.so q many_spheres 3)"
echo "ld256        $(q many_spheres 3)"
echo "ld128        $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libld128.so q hanging_lamps_in_row 1 0 640 360)"
echo "ld256        $(q hanging_lamps_in_row 1 0 640 360)"
} > gpurun_out/quick_r3c.log 2>&1
cat gpurun_out/quick_r3c.log
timeout 900 python -m pytest tests/test_gpu_walk.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
