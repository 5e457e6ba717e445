#!/usr/bin/env bash
# round 2, run W: queues sized by the call (+ overflow retry): whole suite, memory footprint, measured outlier fractions of the f32 parity tests
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_w
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "f32_product_mode" -p no:cacheprovider 2>&1 | grep -E "f32 median|passed|failed" | cut -c1-200
ACN_VERBOSE=1 python tools/quick_bench.py primitives 1 0 64 48 2>&1 | grep -E "queues|ms/step" | cut -c1-200
{
echo "spec         $(ACN_SPECIALIZE=1 q wine_glass 3)"
echo "generic      $(q many_spheres 3)"
} > gpurun_out/quick_r2w.log 2>&1
cat gpurun_out/quick_r2w.log
nvidia-smi --query-gpu=memory.used --format=csv,noheader
timeout 1700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2w.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2w.log
tail -5 gpurun_out/pytest_gpu_r2w.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
rm -rf gpurun_out/spec_cache_w
