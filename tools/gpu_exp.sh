#!/usr/bin/env bash
# timing of every library variant in actinon_b200/variants
for s in ${@:-wine_glass}; do
for f in actinon_b200/variants/lib*.so; do
  ACN_B200_LIBRARY=$PWD/$f python tools/quick_bench.py $s 3 2>&1 | tail -1
done; done
