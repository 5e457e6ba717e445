#!/usr/bin/env bash
# round 2, run 3n: packed evaluation programs — one-word table lookups (relative offset in the word), 32-bit funnel shifts
mkdir -p gpurun_out
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "generic $(q hanging_lamps_in_row 1 0 640 360)"
echo "generic $(q hanging_lamp 1 0 400 520)"
} > gpurun_out/quick_r3n.log 2>&1
cat gpurun_out/quick_r3n.log
timeout 1500 python -m pytest tests/test_gpu_scripted.py tests/test_gpu_configs.py tests/test_gpu_walk.py tests/test_gpu_spec.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
