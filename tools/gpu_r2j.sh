#!/usr/bin/env bash
# round 2, run J: ncu source-level capture of the generic kernels on many_spheres (list walk)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_direct|k_path|k_rays' -s 30 -c 3 -o gpurun_out/prof_r2j_spheres python tools/quick_bench.py many_spheres 1 > gpurun_out/ncu_r2j.log 2>&1; echo "ncu rc $?"
tail -3 gpurun_out/ncu_r2j.log
