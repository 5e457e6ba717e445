#!/usr/bin/env bash
# round 2, run P: refill for prims-only scenes, group lockstep where composite objects are met; full suite; bench; ncu summary of many_spheres
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_p
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "generic      $(q many_spheres 3)"
echo "generic      $(q hanging_lamps_in_row 1 0 640 360)"
echo "generic      $(ACN_SPECIALIZE=0 q wine_glass 3)"
echo "generic      $(ACN_SPECIALIZE=0 q diamond 3)"
echo "generic      $(ACN_SPECIALIZE=0 q primitives 3)"
echo "spec         $(ACN_SPECIALIZE=1 q primitives 3)"
} > gpurun_out/quick_r2p.log 2>&1
cat gpurun_out/quick_r2p.log
timeout 1700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2p.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2p.log
tail -5 gpurun_out/pytest_gpu_r2p.log
python bench.py > gpurun_out/bench_r2p.json 2> gpurun_out/bench_r2p.err; echo "bench rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2p_spheres.csv python tools/quick_bench.py many_spheres 1 > gpurun_out/ncu_list_r2p.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_direct|k_path|k_rays' -s 2 -c 6 -o gpurun_out/prof_r2p_spheres python tools/quick_bench.py many_spheres 1 > gpurun_out/ncu_r2p.log 2>&1; echo "ncu rc $?"
rm -rf gpurun_out/spec_cache_p
