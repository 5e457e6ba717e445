#!/usr/bin/env bash
# round 2, run C: spec / device-image / group tests, bench line, ncu of the specialised kernels, wave-budget sweep
mkdir -p gpurun_out
export ACN_CACHE_DIR=$PWD/gpurun_out/spec_cache
timeout 900 python -m pytest tests/test_gpu_spec.py tests/test_gpu_dimage.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r2c.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2c.log 2>&1; echo "smoke rc $?" >> gpurun_out/smoke_r2c.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc $?"
export ACN_SPECIALIZE=1
for b in 1048576 2097152 4194304 8388608 16777216; do timeout 300 python tools/quick_bench.py wine_glass 5 $b 2>&1 | tail -1; done > gpurun_out/budget_r2c.log
for s in caustic_of_caustic ruby_heart pyramid diamond_video_000049; do
  ACN_SPECIALIZE=0 timeout 300 python tools/quick_bench.py $s 3 2>&1 | tail -1
  ACN_SPECIALIZE=1 timeout 300 python tools/quick_bench.py $s 3 2>&1 | tail -1
done > gpurun_out/quick_r2c.log 2>&1
# ncu: launch list (time only) of one step, then full capture of three mid-step launches of the tracing kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2c.csv python tools/quick_bench.py wine_glass 1 > gpurun_out/ncu_list_r2c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_direct|k_path|k_rays|k_shade' -s 60 -c 4 -o gpurun_out/prof_r2c python tools/quick_bench.py wine_glass 1 > gpurun_out/ncu_full_r2c.log 2>&1; echo "ncu rc $?"
tail -15 gpurun_out/pytest_gpu_r2c.log | cut -c1-300; cat gpurun_out/smoke_r2c.log | tail -5; cat gpurun_out/budget_r2c.log gpurun_out/quick_r2c.log; tail -c 1500 gpurun_out/bench_r2c.json; tail -3 gpurun_out/bench_r2c.err
