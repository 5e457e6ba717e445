#!/usr/bin/env bash
# round 2, run G (2 GPUs): group API over both devices, torchrun bench N=2, render tool both ways
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt
timeout 900 python -m pytest tests/test_gpu_dimage.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_n2.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc $?"
{
for s in primitives diamond wine_glass; do
  python tools/render.py --scene $s --gpus 1 2>&1 | tail -1
  python tools/render.py --scene $s --gpus 2 2>&1 | tail -1
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/render.py --scene $s 2>&1 | grep "passes" | tail -1
done
python tools/render.py --frames diamond_video_000000 diamond_video_000010 diamond_video_000020 diamond_video_000030 --gpus 2 2>&1 | tail -2
} > gpurun_out/render_n2.log 2>&1
tail -8 gpurun_out/pytest_n2.log | cut -c1-200; cat gpurun_out/render_n2.log; tail -c 1200 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
