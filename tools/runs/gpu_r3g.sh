#!/usr/bin/env bash
# round 2, run 3g: final state — whole GPU suite, smoke, bench line (+ reference arm), per-scene timings, ncu of the final walk kernels
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_g
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
timeout 1700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r3g.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r3g.log
tail -4 gpurun_out/pytest_gpu_r3g.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_r3g.json 2> gpurun_out/bench_r3g.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r3g.json 2> gpurun_out/bench_ref_r3g.err; echo "bench reference rc $?"
{
for s in wine_glass diamond primitives; do echo "spec    $(ACN_SPECIALIZE=1 q $s 3)"; done
echo "generic $(q many_spheres 3)"
echo "generic $(q hanging_lamps_in_row 1 0 640 360)"
echo "generic $(q hanging_lamp 1 0 400 520)"
} > gpurun_out/quick_r3g.log 2>&1
cat gpurun_out/quick_r3g.log
ncu --set full --clock-control none --import-source on -k regex:'k_direct|k_path|k_rays' -s 2 -c 6 -o gpurun_out/prof_r3g_spheres python tools/quick_bench.py many_spheres 1 > gpurun_out/ncu_r3g.log 2>&1; echo "ncu rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r3g_spheres.csv python tools/quick_bench.py many_spheres 1 > gpurun_out/ncu_list_r3g.log 2>&1
rm -rf gpurun_out/spec_cache_g
