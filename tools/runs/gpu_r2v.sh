#!/usr/bin/env bash
# round 2, run V: envelopes with protruding contents as plain gates; whole suite; bench line; launch list for profiles/
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_v
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
timeout 900 python -m pytest tests/test_gpu_walk.py -m gpu -x -q -s -p no:cacheprovider > gpurun_out/pytest_gpu_r2v_walk.log 2>&1; echo "walk tests rc $?"
grep -E "^(auto|no_env|cutting|small|offset|many_spheres)|passed|failed" gpurun_out/pytest_gpu_r2v_walk.log | cut -c1-200
{
echo "generic      $(q many_spheres 3)"
echo "generic      $(q hanging_lamps_in_row 1 0 640 360)"
echo "spec         $(ACN_SPECIALIZE=1 q wine_glass 3)"
echo "spec         $(ACN_SPECIALIZE=1 q diamond 3)"
} > gpurun_out/quick_r2v.log 2>&1
cat gpurun_out/quick_r2v.log
timeout 1700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2v.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2v.log
tail -5 gpurun_out/pytest_gpu_r2v.log
python bench.py > gpurun_out/bench_r2v.json 2> gpurun_out/bench_r2v.err; echo "bench rc $?"
ACN_SPECIALIZE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2v.csv python tools/quick_bench.py wine_glass 1 > gpurun_out/ncu_list_r2v.log 2>&1
rm -rf gpurun_out/spec_cache_v
