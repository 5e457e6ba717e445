#!/usr/bin/env bash
# round 2, run 3d: many_spheres k_direct — group body (lowest-record-first lockstep, coherent shadow rays share record loads) vs refill body
mkdir -p gpurun_out
q() { timeout 600 python tools/quick_bench.py "$@" 2>&1 | tail -1 | sed 's/^[^ ]* *//'; }
{
echo "refill       $(q many_spheres 3)"
echo "groups 8blk  $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libdg8.so q many_spheres 3)"
echo "groups 6blk  $(ACN_B200_LIBRARY=$PWD/actinon_b200/variants/libdg6.so q many_spheres 3)"
} > gpurun_out/quick_r3d.log 2>&1
cat gpurun_out/quick_r3d.log
