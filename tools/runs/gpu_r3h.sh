#!/usr/bin/env bash
# round 2, run 3h: launch bounds of the full-featured generic kernels on the lamp scene, after the packed evaluation programs
mkdir -p gpurun_out
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "6,6,8        $(q hanging_lamps_in_row 1 0 640 360)"
for v in g7 g6 g5; do echo "$v           $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/lib$v.so q hanging_lamps_in_row 1 0 640 360)"; done
} > gpurun_out/quick_r3h.log 2>&1
cat gpurun_out/quick_r3h.log
timeout 600 python -m pytest tests/test_gpu_walk.py -m gpu -x -q -s -k "octant or tight" -p no:cacheprovider 2>&1 | grep -E "many_spheres|passed|failed" | cut -c1-200
