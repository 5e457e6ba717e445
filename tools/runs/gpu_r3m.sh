#!/usr/bin/env bash
# round 2, run 3m: ncu source-level capture of k_direct on the lamp scene (160x90), final kernels
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_direct' -s 1 -c 2 -o gpurun_out/prof_r3m_lamps python tools/quick_bench.py hanging_lamps_in_row 1 0 160 90 > gpurun_out/ncu_r3m.log 2>&1; echo "ncu rc $?"
