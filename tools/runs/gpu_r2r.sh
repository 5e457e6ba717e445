#!/usr/bin/env bash
# round 2, run R: pre-order records + lowest-record-first lockstep (lamps), k_direct register diet (many_spheres)
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_r
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "default      $(q many_spheres 3)"
for v in p6 p7; do echo "$v         $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/lib$v.so q many_spheres 3)"; done
echo "default      $(q hanging_lamps_in_row 1 0 640 360)"
echo "every        $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libevery.so q hanging_lamps_in_row 1 0 640 360)"
echo "default      $(q hanging_lamp 1 0 400 520)"
echo "every        $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libevery.so q hanging_lamp 1 0 400 520)"
echo "generic      $(ACN_SPECIALIZE=0 q wine_glass 3)"
echo "every        $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libevery.so ACN_SPECIALIZE=0 q wine_glass 3)"
echo "generic      $(ACN_SPECIALIZE=0 q diamond 3)"
echo "every        $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libevery.so ACN_SPECIALIZE=0 q diamond 3)"
} > gpurun_out/quick_r2r.log 2>&1
cat gpurun_out/quick_r2r.log
timeout 1700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2r.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2r.log
tail -5 gpurun_out/pytest_gpu_r2r.log
rm -rf gpurun_out/spec_cache_r
