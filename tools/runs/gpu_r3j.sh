#!/usr/bin/env bash
# round 2, run 3j: K shadow rays per lane (shadow_query_multi) on the lamp scenes
mkdir -p gpurun_out
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "groups       $(q hanging_lamps_in_row 1 0 640 360)"
for v in m2 m3 m4; do echo "$v           $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/lib$v.so q hanging_lamps_in_row 1 0 640 360)"; done
echo "groups       $(q hanging_lamp 1 0 400 520)"
for v in m2 m4; do echo "$v           $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/lib$v.so q hanging_lamp 1 0 400 520)"; done
} > gpurun_out/quick_r3j.log 2>&1
cat gpurun_out/quick_r3j.log
