#!/usr/bin/env bash
# timing of every library variant in actinon_b200/variants (+ the same with the compaction switched off)
for f in actinon_b200/variants/lib*.so; do
  ACN_B200_LIBRARY=$PWD/$f python tools/quick_bench.py ${1:-wine_glass} 3 ${2:-0} 2>&1 | tail -1
done
f=actinon_b200/variants/libm5.so
[ -f $f ] && ACN_NO_COMPACTION=1 ACN_B200_LIBRARY=$PWD/$f python tools/quick_bench.py ${1:-wine_glass} 3 ${2:-0} 2>&1 | tail -1
