#!/usr/bin/env bash
# round 2, run H: leader-computed sampling bases + window prefetch; per-pass breakdown of a full render; ncu of the final kernels
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache_h
q() { timeout 900 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
for s in wine_glass diamond primitives; do echo "generic $(ACN_SPECIALIZE=0 q $s 3)"; echo "spec    $(ACN_SPECIALIZE=1 q $s 3)"; done
echo "generic $(q many_spheres 3)"
echo "generic $(q hanging_lamps_in_row 1 0 640 360)"
} > gpurun_out/quick_r2h.log 2>&1
timeout 600 python tools/render.py --scene wine_glass --verbose > gpurun_out/render_verbose_r2h.log 2>&1
timeout 600 python tools/render.py --scene wine_glass --host-controller >> gpurun_out/render_verbose_r2h.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_spec.py tests/test_gpu_dimage.py tests/test_gpu_configs.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2h.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2h.log
export ACN_SPECIALIZE=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2h.csv python tools/quick_bench.py wine_glass 1 > gpurun_out/ncu_list_r2h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_direct|k_path|k_rays|k_shade' -s 60 -c 4 -o gpurun_out/prof_r2h python tools/quick_bench.py wine_glass 1 > gpurun_out/ncu_full_r2h.log 2>&1; echo "ncu rc $?"
cat gpurun_out/quick_r2h.log; head -14 gpurun_out/render_verbose_r2h.log; tail -4 gpurun_out/render_verbose_r2h.log; tail -4 gpurun_out/pytest_gpu_r2h.log
