#!/usr/bin/env bash
# round 2, run L: refill loops in the walking kernels (k_rays / k_path / k_direct), prims-only instantiation — parity first, then timings
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_l
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2l.log 2>&1; echo "pytest parity rc $?"
tail -3 gpurun_out/pytest_gpu_r2l.log
{
echo "default      $(q many_spheres 3)"
for v in p6 p5 p8r4 p8r16 p886; do echo "$v         $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/lib$v.so q many_spheres 3)"; done
echo "default      $(q hanging_lamps_in_row 1 0 640 360)"
echo "p5           $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libp5.so q hanging_lamps_in_row 1 0 640 360)"
echo "generic      $(ACN_SPECIALIZE=0 q wine_glass 3)"
echo "spec         $(ACN_SPECIALIZE=1 q wine_glass 3)"
} > gpurun_out/quick_r2l.log 2>&1
cat gpurun_out/quick_r2l.log
timeout 1700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2l_all.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2l_all.log
tail -5 gpurun_out/pytest_gpu_r2l_all.log
rm -rf gpurun_out/spec_cache_l
