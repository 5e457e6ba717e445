#!/usr/bin/env bash
# A/B timing of the library variants on several scenes, then the GPU parity tests with the default library
bash tools/gpu_exp.sh ${@:-wine_glass many_spheres diamond} 2>&1 | tee gpurun_out/ab.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
