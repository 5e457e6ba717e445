#!/usr/bin/env bash
# round 2, run 3k: GROUP records (pure bounds over runs of consecutive elements of a long list) on the lamp scenes
mkdir -p gpurun_out
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "no groups    $(ACN_NO_GROUP_RECORDS=1 q hanging_lamps_in_row 1 0 640 360)"
echo "groups       $(q hanging_lamps_in_row 1 0 640 360)"
echo "no groups    $(ACN_NO_GROUP_RECORDS=1 q hanging_lamp 1 0 400 520)"
echo "groups       $(q hanging_lamp 1 0 400 520)"
} > gpurun_out/quick_r3k.log 2>&1
cat gpurun_out/quick_r3k.log
ACN_VERBOSE=1 python tools/quick_bench.py hanging_lamps_in_row 1 0 64 36 2>&1 | grep "traversal records"
timeout 1500 python -m pytest tests/test_gpu_scripted.py tests/test_gpu_configs.py tests/test_gpu_walk.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
