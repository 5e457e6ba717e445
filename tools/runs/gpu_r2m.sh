#!/usr/bin/env bash
# round 2, run M: where does the refill k_direct lose?  no-accumulation variant + ncu of the first k_direct launches
mkdir -p gpurun_out
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "default      $(q many_spheres 3)"
echo "noacc        $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libnoacc.so q many_spheres 3)"
} > gpurun_out/quick_r2m.log 2>&1
cat gpurun_out/quick_r2m.log
ncu --set full --clock-control none --import-source on -k regex:'k_direct|k_path' -s 0 -c 4 -o gpurun_out/prof_r2m_spheres python tools/quick_bench.py many_spheres 1 > gpurun_out/ncu_r2m.log 2>&1; echo "ncu rc $?"
