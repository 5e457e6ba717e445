#!/usr/bin/env bash
# round 2, run E: eight-children records (wide walk) for the global-memory kernels: whole test suite, timings, bench
mkdir -p gpurun_out
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "# many_spheres: list walk vs wide walk, launch bounds (min blocks/SM rays,path,direct)"
echo "list walk  6,6,8  $(ACN_NO_WIDE=1 q many_spheres 3)"
echo "wide walk  6,6,8  $(q many_spheres 3)"
for v in g666 g555 g444; do echo "wide walk  $v   $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/lib$v.so q many_spheres 3)"; done
echo "# lamps (full-featured kernels, 640x360)"
echo "list walk  $(ACN_NO_WIDE=1 q hanging_lamps_in_row 1 0 640 360)"
echo "wide walk  $(q hanging_lamps_in_row 1 0 640 360)"
echo "wide g555  $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libg555.so q hanging_lamps_in_row 1 0 640 360)"
echo "list walk  $(ACN_NO_WIDE=1 q hanging_lamp 1 0 400 520)"
echo "wide walk  $(q hanging_lamp 1 0 400 520)"
} > gpurun_out/wide_r2e.log 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2e.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2e.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench rc $?"
ncu --set full --clock-control none --import-source on -k regex:'k_direct|k_path|k_rays' -s 30 -c 3 -o gpurun_out/prof_r2e_spheres python tools/quick_bench.py many_spheres 1 > gpurun_out/ncu_r2e.log 2>&1; echo "ncu rc $?"
cat gpurun_out/wide_r2e.log; tail -6 gpurun_out/pytest_gpu_r2e.log; tail -c 3000 gpurun_out/bench_r2e.json; tail -3 gpurun_out/bench_r2e.err
