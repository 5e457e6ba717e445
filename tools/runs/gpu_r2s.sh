#!/usr/bin/env bash
# round 2, run S: csg_state jumps over the subtrees of missed sub-envelopes
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_t
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "default      $(q hanging_lamps_in_row 1 0 640 360)"
echo "default      $(q hanging_lamp 1 0 400 520)"
echo "generic      $(ACN_SPECIALIZE=0 q wine_glass 3)"
echo "spec         $(ACN_SPECIALIZE=1 q wine_glass 3)"
} > gpurun_out/quick_r2t.log 2>&1
cat gpurun_out/quick_r2t.log
timeout 1700 python -m pytest tests/test_gpu_configs.py tests/test_gpu_scripted.py tests/test_gpu_spec.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2t.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2t.log
tail -5 gpurun_out/pytest_gpu_r2t.log
ncu --set full --clock-control none --import-source on -k regex:'k_direct' -s 1 -c 2 -o gpurun_out/prof_r2t_lamps python tools/quick_bench.py hanging_lamps_in_row 1 0 160 90 > gpurun_out/ncu_r2t.log 2>&1; echo "ncu rc $?"
rm -rf gpurun_out/spec_cache_t
