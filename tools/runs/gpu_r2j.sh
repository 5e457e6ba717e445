#!/usr/bin/env bash
# round 2, run J: tight cull bounds in the traversal records (many_spheres) — A/B timing, parity tests, ncu source-level capture
mkdir -p gpurun_out
q() { timeout 900 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "envelopes    $(ACN_NO_TIGHT_BOUNDS=1 q many_spheres 3)"
echo "tight bounds $(ACN_VERBOSE=1 q many_spheres 3)"
echo "envelopes    $(ACN_NO_TIGHT_BOUNDS=1 q hanging_lamps_in_row 1 0 640 360)"
echo "tight bounds $(q hanging_lamps_in_row 1 0 640 360)"
echo "tight bounds spec $(ACN_SPECIALIZE=1 q wine_glass 3)"
} > gpurun_out/quick_r2j.log 2>&1
cat gpurun_out/quick_r2j.log
ACN_VERBOSE=1 python tools/quick_bench.py many_spheres 1 2>&1 | grep "acn:" | head -3
timeout 1500 python -m pytest tests/test_gpu_scripted.py tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2j.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2j.log
tail -4 gpurun_out/pytest_gpu_r2j.log
ncu --set full --clock-control none --import-source on -k regex:'k_direct|k_path|k_rays' -s 6 -c 12 -o gpurun_out/prof_r2j_spheres python tools/quick_bench.py many_spheres 1 > gpurun_out/ncu_r2j.log 2>&1; echo "ncu rc $?"
