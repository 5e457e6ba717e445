#!/usr/bin/env bash
# round 2, run 3b: shadow rays walk a second copy of the records sorted by (area of the bound) / (cost of the test)
mkdir -p gpurun_out
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "scene order  $(ACN_NO_ANY_ORDER=1 q hanging_lamps_in_row 1 0 640 360)"
echo "any order    $(ACN_VERBOSE=1 q hanging_lamps_in_row 1 0 640 360)"
echo "scene order  $(ACN_NO_ANY_ORDER=1 q hanging_lamp 1 0 400 520)"
echo "any order    $(q hanging_lamp 1 0 400 520)"
echo "any order    $(q many_spheres 3)"
} > gpurun_out/quick_r3b.log 2>&1
cat gpurun_out/quick_r3b.log
ACN_VERBOSE=1 python tools/quick_bench.py hanging_lamps_in_row 1 0 64 36 2>&1 | grep "traversal records"
timeout 1700 python -m pytest tests/test_gpu_scripted.py tests/test_gpu_configs.py tests/test_gpu_walk.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r3b.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r3b.log
tail -4 gpurun_out/pytest_gpu_r3b.log
