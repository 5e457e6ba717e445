#!/usr/bin/env bash
# round 2, run B: clean timings (generic vs scene-specialised kernels), new GPU tests
mkdir -p gpurun_out
export ACN_VERBOSE=1
for s in wine_glass diamond primitives; do
  ACN_SPECIALIZE=0 timeout 300 python tools/quick_bench.py $s 5 2>&1 | tail -1
  ACN_SPECIALIZE=1 timeout 300 python tools/quick_bench.py $s 5 2>&1 | grep -v "^acn: object" | tail -2
done > gpurun_out/quick_r2b.log 2>&1
timeout 300 python tools/quick_bench.py many_spheres 5 2>&1 | tail -1 >> gpurun_out/quick_r2b.log
unset ACN_VERBOSE
timeout 1500 python -m pytest tests/test_gpu_spec.py tests/test_gpu_dimage.py tests/test_gpu_scripted.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2b.log
python tools/parity_report.py --merge >> gpurun_out/pytest_gpu_r2b.log 2>&1
cat gpurun_out/quick_r2b.log; grep -v "^\.$" gpurun_out/pytest_gpu_r2b.log | tail -60 | cut -c1-300
