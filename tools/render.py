#!/usr/bin/env python
"""Renders a scene to a .pnm like `actinon <script.acn>` does — on 1..8 B200s.

    python tools/render.py --scene wine_glass --out wine_glass.pnm [--passes K] [--width W --height H]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/render.py --scene diamond --out d.pnm
    ... --frames diamond_video_0000{00,10,20}     # frame sequence: frame i -> rank i mod N, no communication

Within an image every rank runs the SAME host pass controller (scene.c:1103-1159: pass 0 = pixel centres, then
gradient_cycles passes of jittered re-sampling where the image gradient exceeds the threshold) on the SAME
accumulated image, traces only the samples whose 8x8 pixel tile it owns, and the per-pixel sums of the pass
(position, colour, weight: scene.c:804-813) are all-reduced over NCCL — the one exchange step of the path.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TILE = 8


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="wine_glass")
    ap.add_argument("--frames", nargs="*", default=None, help="scene names of a frame sequence (sharded by frame)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--passes", type=int, default=None, help="stop after this many passes (default: gradient_cycles + 1)")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--direct-samples", type=int, default=None)
    ap.add_argument("--path-samples", type=int, default=None)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))

    import torch
    import torch.distributed as dist
    import actinon_b200 as acn
    acn.device_count()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    ov = {}
    for k, v in (("image_width", args.width), ("image_height", args.height), ("direct_samples", args.direct_samples),
                 ("path_samples", args.path_samples)):
        if v is not None:
            ov[k] = v

    def render_one(name, out, ranks, my):
        """ranks/my: how many ranks share this image and which of them I am."""
        flat = acn.scenes.load(name, **ov)
        prm = flat.params
        W, H = prm.image_width, prm.image_height
        tracer = acn.Tracer(flat, acn.Options(device=local))
        img = acn.Image(W, H)
        t0 = time.perf_counter()
        n_pass, n_samples, rays = 0, 0, 0
        while args.passes is None or n_pass < args.passes:
            xy = img.next_pass(prm)
            if xy.shape[0] == 0:
                break
            if ranks > 1:
                tx, ty = xy[:, 0].astype(np.int64) // TILE, xy[:, 1].astype(np.int64) // TILE
                mine = xy[(tx + ty) % ranks == my]
            else:
                mine = xy
            rgb = tracer.render_samples(mine) if len(mine) else np.empty((0, 3), np.float32)
            rays += tracer.last_stats.rays if len(mine) else 0
            if ranks > 1:
                delta = acn.Image(W, H)
                delta.push(mine, rgb)
                d = torch.from_numpy(delta.sums()).to(dev)
                dist.all_reduce(d)                           # per-pixel sums of the pass, disjoint tiles
                img.push(np.empty((0, 2)), np.empty((0, 3), np.float32))    # advances cycle + jitter stream like the reference
                img.add_sums(d.cpu().numpy())
            else:
                img.push(mine, rgb)
            n_samples += len(xy)
            n_pass += 1
        dt = time.perf_counter() - t0
        h = img.write_pnm(out if (out and my == 0) else None)
        if my == 0:
            print(f"{name}: {W}x{H}, {n_pass} passes, {n_samples} samples in {dt:.2f} s on {ranks} GPU(s) "
                  f"({n_samples / dt:.0f} samples/s, {rays * ranks / dt / 1e9:.2f} G rays/s); image hash {h:016x}"
                  + (f" -> {out}" if out else ""), flush=True)
        tracer.close()
        return img

    if args.frames:
        for i, name in enumerate(args.frames):
            if i % world == rank:
                render_one(name, (args.out or "frame") + f".{name}.pnm", 1, 0)
    else:
        render_one(args.scene, args.out, world, rank)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
