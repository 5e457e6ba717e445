#!/usr/bin/env bash
# round 2, run D: block-cooperative k_shade, code-size-aware specialisation, launch-bounds sweep of the specialised kernels
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_d
export ACN_SPECIALIZE=1
q() { timeout 300 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^default *//'; }
{
echo "# generic vs spec (defaults: run_unroll 6, ool squaroid)"
for s in wine_glass diamond diamond_video_000049 ruby_heart pyramid caustic_of_caustic primitives; do
  echo "generic $(ACN_SPECIALIZE=0 q $s 3)"; echo "spec    $(q $s 3)"
done
echo "# generator options"
for s in diamond diamond_video_000049 ruby_heart; do
  echo "unroll=64 ool=0 $(ACN_SPEC_RUN_UNROLL=64 ACN_SPEC_OOL_SQUAROID=0 q $s 3)"
  echo "unroll=64 ool=1 $(ACN_SPEC_RUN_UNROLL=64 q $s 3)"
  echo "unroll=0  ool=1 $(ACN_SPEC_RUN_UNROLL=0 q $s 3)"
  echo "unroll=3  ool=0 $(ACN_SPEC_RUN_UNROLL=3 ACN_SPEC_OOL_SQUAROID=0 q $s 3)"
done
echo "# launch bounds (min blocks per SM) of the specialised kernels"
for mb in "4 4 4" "6 6 6" "4 5 5" "5 4 5" "5 5 4" "6 5 5" "5 6 5" "5 5 6" "4 4 6"; do set -- $mb
  for s in wine_glass diamond; do echo "minb rays/path/direct $mb: $(ACN_SPEC_OPTS="-DACN_MINB_RAYS=$1 -DACN_MINB_PATH=$2 -DACN_MINB_DIRECT=$3" q $s 3)"; done
done
} > gpurun_out/sweep_r2d.log 2>&1
unset ACN_SPECIALIZE
timeout 900 python -m pytest tests/test_gpu_spec.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2d.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2d.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench rc $?"
cat gpurun_out/sweep_r2d.log; tail -4 gpurun_out/pytest_gpu_r2d.log; tail -c 2500 gpurun_out/bench_r2d.json; tail -3 gpurun_out/bench_r2d.err
