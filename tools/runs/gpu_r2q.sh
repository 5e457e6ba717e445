#!/usr/bin/env bash
# round 2, run Q: ncu source-level capture of k_direct on the lamp scene (160x90 so that the replays stay short)
mkdir -p gpurun_out
timeout 900 python tools/quick_bench.py hanging_lamps_in_row 1 0 160 90 2>&1 | tail -1 | cut -c50-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2q_lamps.csv python tools/quick_bench.py hanging_lamps_in_row 1 0 160 90 > gpurun_out/ncu_list_r2q.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_direct' -s 1 -c 3 -o gpurun_out/prof_r2q_lamps python tools/quick_bench.py hanging_lamps_in_row 1 0 160 90 > gpurun_out/ncu_r2q.log 2>&1; echo "ncu rc $?"
