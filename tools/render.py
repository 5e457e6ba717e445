#!/usr/bin/env python
"""Renders a scene to a .pnm like `actinon <script.acn>` does — on 1..8 B200s.

    python tools/render.py --scene wine_glass --out wine_glass.pnm [--gpus N] [--passes K] [--width W --height H]
    python tools/render.py --frames diamond_video_0000{00,10,20} --gpus 8       # frame sequence: frame i -> GPU i mod N
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/render.py --scene diamond --out d.pnm

The pass controller of the reference (scene.c:1103-1159: pass 0 = pixel centres, then gradient_cycles passes of jittered
re-sampling where the image gradient exceeds the threshold) runs ON THE DEVICE (acn_dimage_*).  Several GPUs:
 * --gpus N in one process: acn_group_* — a worker thread, a tracer and a device image per GPU, 4x4 pixel tiles dealt to the
   GPUs along a Morton curve, the per-pass sums exchanged through peer memory (NVLink);
 * under torchrun (one process per GPU): the same tiles, the per-pass sums all-reduced over NCCL (uint64, exact).
Either way the image is bit-identical to the single-GPU one.  --host-controller runs the host restatement of the pass
controller instead (acn_image_*; one GPU), the path the device controller is tested against.
"""
import argparse
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="wine_glass")
    ap.add_argument("--frames", nargs="*", default=None, help="scene names of a frame sequence (sharded by frame)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--gpus", type=int, default=1, help="GPUs of this process (acn_group); under torchrun: one per process")
    ap.add_argument("--passes", type=int, default=None, help="stop after this many passes (default: gradient_cycles + 1)")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--direct-samples", type=int, default=None)
    ap.add_argument("--path-samples", type=int, default=None)
    ap.add_argument("--generic", action="store_true", help="generic kernels instead of the scene-specialised ones")
    ap.add_argument("--host-controller", action="store_true")
    ap.add_argument("--verbose", action="store_true", help="per-pass samples, rays, waves, device and wall time (one GPU, device controller)")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))

    import actinon_b200 as acn
    n_dev = acn.device_count()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ov = {}
    for k, v in (("image_width", args.width), ("image_height", args.height), ("direct_samples", args.direct_samples),
                 ("path_samples", args.path_samples)):
        if v is not None:
            ov[k] = v

    def options(device):
        return dict(device=device, specialize=acn.SPECIALIZE_OFF if args.generic else acn.SPECIALIZE_ON)

    def tracer_for(flat, device):
        try:
            return acn.Tracer(flat, acn.Options(**options(device)))
        except acn.AcnError as e:
            if e.code != -5:
                raise
            return acn.Tracer(flat, acn.Options(device=device, specialize=acn.SPECIALIZE_OFF))       # the scene does not qualify

    def report(name, flat, img, n_samples, n_pass, rays, dt, gpus, out):
        h = img.write_pnm(out)
        prm = flat.params
        print(f"{name}: {prm.image_width}x{prm.image_height}, {n_pass} passes, {n_samples} samples in {dt:.2f} s on {gpus} GPU(s) "
              f"({n_samples / dt:.0f} samples/s, {rays / dt / 1e9:.2f} G rays/s); image hash {h:016x}" + (f" -> {out}" if out else ""), flush=True)

    def render_one(name, out, device, group_gpus=1, use_dist=None):
        flat = acn.scenes.load(name, **ov)
        if group_gpus > 1:
            try:
                g = acn.Group(flat, list(range(group_gpus)), acn.Options(**options(-1)))
            except acn.AcnError as e:
                if e.code != -5:
                    raise
                g = acn.Group(flat, list(range(group_gpus)), acn.Options(specialize=acn.SPECIALIZE_OFF))
            t0 = time.perf_counter()
            n_samples, n_pass, rays = g.render(args.passes)
            dt = time.perf_counter() - t0
            img = g.download(); g.close()
            report(name, flat, img, n_samples, n_pass, rays, dt, group_gpus, out)
            return
        tracer = tracer_for(flat, device)
        t0 = time.perf_counter()
        if args.host_controller:
            img = acn.Image(flat.params.image_width, flat.params.image_height)
            n_samples = n_pass = rays = 0
            while args.passes is None or n_pass < args.passes:
                xy = img.next_pass(flat.params)
                if xy.shape[0] == 0:
                    break
                img.push(xy, tracer.render_samples(xy, n_samples))
                n_samples += xy.shape[0]; n_pass += 1; rays += tracer.last_stats.rays
        else:
            log = None
            if args.verbose and rank == 0:
                def log(k, nt, nl, st, wall):
                    print(f"  pass {k:3d}: {nt:7d} samples ({nl} here), {st.rays / 1e6:7.1f} M rays, {st.waves:3d} waves, {st.kernel_launches:4d} launches, "
                          f"device {st.device_ms:7.2f} ms, wall {wall * 1e3:7.2f} ms", flush=True)
            img, n_samples, n_pass, rays = acn.render_image_device(flat, tracer, args.passes, use_dist, log)
        dt = time.perf_counter() - t0
        tracer.close()
        if use_dist is not None:
            import torch
            t = torch.tensor([float(rays)], dtype=torch.float64, device="cuda"); use_dist.all_reduce(t); rays = float(t[0])
        if use_dist is None or rank == 0:
            report(name, flat, img, n_samples, n_pass, rays, dt, world if use_dist is not None else 1, out)

    if args.frames:
        if world > 1:                                           # one process per GPU: frame i -> rank i mod N
            for i, name in enumerate(args.frames):
                if i % world == rank:
                    render_one(name, (args.out or "frame") + f".{name}.pnm", local)
        else:                                                   # one process: a thread per GPU, frame i -> GPU i mod N
            gpus = max(1, min(args.gpus, n_dev))
            def work(d):
                for i, name in enumerate(args.frames):
                    if i % gpus == d:
                        render_one(name, (args.out or "frame") + f".{name}.pnm", d)
            th = [threading.Thread(target=work, args=(d,)) for d in range(gpus)]
            t0 = time.perf_counter()
            for t in th: t.start()
            for t in th: t.join()
            print(f"{len(args.frames)} frames on {gpus} GPU(s) in {time.perf_counter() - t0:.2f} s", flush=True)
    elif world > 1:
        render_one(args.scene, args.out, local, 1, dist)
    else:
        render_one(args.scene, args.out, 0, max(1, min(args.gpus, n_dev)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
