#!/usr/bin/env bash
# round 2, run F: scale nodes inside the event sweep (no recursive march in FP32), wide walk variants, full suite, bench
mkdir -p gpurun_out
q() { timeout 900 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "# lamps: full-featured FP32 kernels, every object swept (scale nodes inside the sweep)"
echo "hanging_lamps_in_row 640x360  $(q hanging_lamps_in_row 1 0 640 360)"
echo "hanging_lamp 400x520          $(q hanging_lamp 1 0 400 520)"
echo "paraffin_lamp                 $(q paraffin_lamp 1)"
echo "# many_spheres: list walk vs wide walk (ACN_WIDE=1), launch bounds"
echo "list 6,6,8  $(q many_spheres 3)"
echo "wide 6,6,8  $(ACN_WIDE=1 q many_spheres 3)"
for v in g555 g444; do echo "list $v   $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/lib$v.so q many_spheres 3)"; echo "wide $v   $(ACN_WIDE=1 ACN_B200_LIBRARY=$PWD/actinon_b200/variants/lib$v.so q many_spheres 3)"; done
echo "wide g555 lamps $(ACN_WIDE=1 ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libg555.so q hanging_lamps_in_row 1 0 640 360)"
echo "list g555 lamps $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libg555.so q hanging_lamps_in_row 1 0 640 360)"
echo "list g444 lamps $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libg444.so q hanging_lamps_in_row 1 0 640 360)"
} > gpurun_out/lamps_r2f.log 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2f.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2f.log
ACN_WIDE=1 timeout 1200 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2f_wide.log 2>&1; echo "pytest(wide) rc $?" >> gpurun_out/pytest_gpu_r2f_wide.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "bench rc $?"
cat gpurun_out/lamps_r2f.log; tail -6 gpurun_out/pytest_gpu_r2f.log; tail -3 gpurun_out/pytest_gpu_r2f_wide.log; tail -c 1500 gpurun_out/bench_r2f.json; tail -3 gpurun_out/bench_r2f.err
