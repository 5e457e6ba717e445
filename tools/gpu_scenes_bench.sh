#!/usr/bin/env bash
for s in ${@:-primitives wine_glass many_spheres diamond hanging_lamps_in_row}; do timeout 600 python tools/quick_bench.py $s 2 2>&1 | tail -1; done
