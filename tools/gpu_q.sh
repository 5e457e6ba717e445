#!/usr/bin/env bash
# quick timing of the default library on the scenes given (env VARIANTS="A=1 B=2" runs each setting), then the GPU parity tests
for s in ${@:-many_spheres wine_glass diamond primitives}; do
  for v in ${VARIANTS:-_=_}; do
    echo -n "$v "; env $v timeout 120 python tools/quick_bench.py $s 3 2>&1 | tail -1 | cut -c52- | sed 's/mean_rgb/rgb/'
  done
done 2>&1 | tee gpurun_out/q.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
