#!/usr/bin/env bash
# round 2, run N: refill loops with the per-ray epilogue deferred to the refill point
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_n
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "default      $(q many_spheres 3)"
for v in p6 p6r12 p6r16 p5r12; do echo "$v         $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/lib$v.so q many_spheres 3)"; done
echo "default      $(q hanging_lamps_in_row 1 0 640 360)"
echo "p6r16        $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libp6r16.so q hanging_lamps_in_row 1 0 640 360)"
echo "generic      $(ACN_SPECIALIZE=0 q wine_glass 3)"
echo "generic      $(ACN_SPECIALIZE=0 q diamond 3)"
} > gpurun_out/quick_r2n.log 2>&1
cat gpurun_out/quick_r2n.log
timeout 1700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2n.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2n.log
tail -5 gpurun_out/pytest_gpu_r2n.log
rm -rf gpurun_out/spec_cache_n
