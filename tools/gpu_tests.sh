#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | tail -60
