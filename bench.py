#!/usr/bin/env python
"""bench.py — path samples/s and rays/s of the hot path (lum_machine_s_run, reference src/scene.c:1017) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scene NAME] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

One STEP = pass 0 of the reference's pass controller (every pixel centre, scene.c:1110-1120) of BASELINE.json configs[1],
wine_glass.acn exactly as scripted (direct_samples 200, path_samples 500, trace_depth 25), through the product's own pass
API (acn_dimage_render_pass: sample list built on the device, traced, accumulated per pixel in fixed point).

 N = 1   the scripted 400x400 image: 160 000 samples per step.
 N > 1   WEAK scaling: the same scene rendered at sqrt(N) x the resolution (566^2, 800^2, 1131^2 for N = 2, 4, 8), i.e.
         160 000 samples per GPU and step; pixels are dealt to the ranks in 4x4 tiles along a Morton curve, every rank traces
         its own tiles, and the per-pixel sums of the pass are exchanged with ONE all-reduce (NCCL, uint64 sums, disjoint
         support) inside the timed region.  `extras.strong` holds the STRONG-scaling figure (the scripted 400x400 pass split
         over the N ranks), `extras.configs` pass 0 of many_spheres (C3), diamond (C4) and hanging_lamps_in_row 640x360 (C5)
         split over the N ranks, `extras.video` frames of diamond_video.acn dealt to the ranks (C5).

 value    whole-job samples/s, everything resident in HBM (CUDA events on the image's stream, max over ranks)
 e2e      N = 1: the reference-facing call with HOST buffers (acn_render_samples = lum_machine_s_run(scene, lum_arr)), H2D of the
          positions and D2H of the colours inside the timed region.  N > 1: H2D of the rank's positions, trace, accumulate,
          all-reduce, D2H of the reduced image.
 roofline FP32-ALU roofline of the dominant kernel: algorithmic FLOPs (oracle counters x SURVEY.md §8d constants, reference
          traversal order) / CUDA-event time of that kernel's launches in the step
 cpu_baseline   the CPU oracle (FP64, position-hash seeding, all host threads) on the same scene
 --impl reference   the CPU oracle alone (the reference itself cannot be built: its dependency beth is absent)

The kernels are the scene-specialised ones (acn_options.specialize = ON: compiled by NVRTC at acn_tracer_create for the
scene's structure, same leaf arithmetic as the generic kernels; many_spheres and the lamp scenes do not qualify and run the
generic kernels).  Inputs are "synthetic" only in the sense of the contract: the geometry is the reference's own scripted
scene, flattened by our .acn front-end (scenes/*.npz); no dataset is involved.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "path_samples_per_sec"
UNIT = "samples/s"
CPU_FLAGS = "g++ -O3 -march=native -ffp-contract=off (FP64; the reference's makefile adds -march=native to beth's unknown base flags)"


def scaled_size(w: int, h: int, n: int):
    """Image size with n times the pixels of w x h (same aspect)."""
    if n == 1:
        return w, h
    f = np.sqrt(n)
    return int(round(w * f)), int(round(h * f))


def pass0_positions(width: int, height: int) -> np.ndarray:
    ys, xs = np.mgrid[0:height, 0:width]
    return np.ascontiguousarray(np.stack([(xs + 0.5).ravel(), (ys + 0.5).ravel()], axis=1).astype(np.float64))


# --------------------------------------------------------------------------------------------------
# clocks during the timed region (profiling recipe's clocks line)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="acn_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.device_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for k, nm in enumerate(names):
                    if f[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# CPU arm (oracle): cpu_baseline and --impl reference
# --------------------------------------------------------------------------------------------------
def cpu_sample(xy: np.ndarray, stride: int) -> np.ndarray:
    return np.ascontiguousarray(xy[::stride])


def cpu_probe_stride(orc, flat, xy, target_s: float, threads: int) -> int:
    """Sub-sampling stride so that one CPU pass takes about target_s seconds (coprime to the row length: the subset stays
    spread over the image)."""
    probe = cpu_sample(xy, 97)
    t0 = time.perf_counter()
    orc.render(flat, probe, seed_mode=0, threads=threads)
    dt = max(time.perf_counter() - t0, 1e-4)
    full = dt * len(xy) / len(probe)
    if full <= target_s:
        return 1
    stride = int(np.ceil(full / target_s))
    while np.gcd(stride, flat.params.image_width) != 1:
        stride += 1
    return stride


def workload_config(scene_name, prm, n, scripted_wh):
    W, H = prm.image_width, prm.image_height
    return {
        "workload": f"{scene_name}.acn as scripted (direct_samples {prm.direct_samples}, path_samples {prm.path_samples}, trace_depth "
                    f"{prm.trace_depth}); one step = pass 0 of the pass controller = every pixel centre of the {W}x{H} image "
                    f"({W * H} samples)" + ("" if n == 1 else f": the scripted {scripted_wh[0]}x{scripted_wh[1]} image at sqrt({n}) x the "
                    f"resolution, {W * H // n} samples per GPU (weak scaling)") + ", BASELINE.json configs[1]",
        "samples_per_step": W * H,
        "sharding": "single GPU" if n == 1 else "4x4 pixel tiles dealt to the ranks along a Morton curve; one NCCL all-reduce (uint64 sums) "
                    "of the pass's per-pixel sums per step",
        "kernels": "scene-specialised (NVRTC at acn_tracer_create; acn_options.specialize = ON)",
        "l2": "ray/task/hit queues (>1 GB per step) exceed the 126 MB L2; a 256 MB buffer is overwritten between timed steps",
    }


def run_reference_arm(args, flat, scene_name, scripted_wh):
    """The reference's CPU implementation of the path (restated: oracle/, FP64, all host threads)."""
    from tests.oracle_lib import Oracle
    orc = Oracle()
    prm = flat.params
    W, H = prm.image_width, prm.image_height
    threads = os.cpu_count() or 1
    xy_full = pass0_positions(W, H)
    stride = cpu_probe_stride(orc, flat, xy_full, 6.0, threads)
    xy = cpu_sample(xy_full, stride)
    rays = 0
    for _ in range(args.warmup):
        orc.render(flat, xy, seed_mode=0, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, info = orc.render(flat, xy, seed_mode=0, threads=threads)
        rays += info["rays"]
    dt = time.perf_counter() - t0
    v = len(xy) * args.steps / dt
    sample = f"{len(xy)} of the {len(xy_full)} pass-0 samples per step (every {stride}th pixel centre), {scene_name} as scripted"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (reference scene script, flattened)",
        "config": workload_config(scene_name, prm, args.gpus, scripted_wh),
        "rays_per_sec": rays / dt,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "flags": CPU_FLAGS},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference binary unbuildable here (depends on johsteffens/beth, absent); this is the FP64 CPU oracle "
                "that restates scene.c/compound.c/objects.c, dynamic per-sample scheduling on all host threads",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
class Rig:
    """One scene on this rank: tracer (scene-specialised kernels when the scene qualifies), sharded device image, and the
    step = pass 0 through the product's pass API + the exchange of the pass's sums."""

    def __init__(self, acn, torch, dist, flat, local_rank, rank, world):
        self.acn, self.torch, self.dist = acn, torch, dist
        self.flat, self.prm = flat, flat.params
        self.rank, self.world = rank, world
        self.dev = torch.device("cuda", local_rank)
        o = dict(seed_mode=acn.SEED_POSITION_HASH, precision=acn.PRECISION_F32, device=local_rank)
        try:
            self.tracer = acn.Tracer(flat, acn.Options(specialize=acn.SPECIALIZE_ON, **o))
            self.specialised = True
        except acn.AcnError as e:
            if e.code != -5:
                raise
            self.tracer = acn.Tracer(flat, acn.Options(specialize=acn.SPECIALIZE_OFF, **o))      # the scene does not qualify
            self.specialised = False
        W, H = self.prm.image_width, self.prm.image_height
        self.img = acn.DeviceImage(W, H, local_rank)
        self.img.set_shard(world, rank, 4)
        # The pass runs on the image's own stream and returns synchronised; the exchange of the pass sums (NCCL) and the
        # timing events use torch's current stream.  Every step ends with a host synchronisation of both, so an event pair
        # on torch's stream around a step brackets all of its device work.
        self.stream = torch.cuda.current_stream(self.dev)
        self.buf = torch.zeros(self.img.words, dtype=torch.int64, device=self.dev) if world > 1 else None
        self.launches = 0
        self.rays = 0
        self.n_local = 0

    def step(self):
        self.img.reset()
        nl, nt = self.img.render_pass(self.tracer, self.prm)
        st = self.tracer.last_stats
        self.launches += (st.kernel_launches if nl else 0) + 4 + (2 if self.world > 1 else 1)      # select, scan, emit, accumulate, commit
        self.rays += st.rays if nl else 0
        self.n_local = nl
        if self.world > 1:
            self.img.copy_delta(self.buf, self.stream)
            self.dist.all_reduce(self.buf)                 # disjoint pixel tiles, integer sums: exact, order-independent
            self.img.set_delta(self.buf, self.stream)
            self.stream.synchronize()
            self.img.end_pass()
        return st

    def timed(self, steps, warmup, flush=None, barrier=None):
        torch = self.torch
        for _ in range(warmup):
            self.step()
        if barrier:
            barrier()
        ms = []
        self.launches = 0; self.rays = 0
        by_class = {}
        for _ in range(steps):
            if flush is not None:
                flush.fill_(1)
                torch.cuda.synchronize(self.dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            st = self.step()
            e1.record(self.stream)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
            for k in ("rays_primary", "rays_reflection", "rays_chromatic", "rays_refraction", "rays_path", "rays_shadow", "diffuse_hits"):
                by_class[k] = by_class.get(k, 0) + getattr(st, k)
        if barrier:
            barrier()
        return ms, by_class

    def close(self):
        self.img.close(); self.tracer.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="wine_glass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong-scaling / other-config / video records")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import actinon_b200 as acn
    scripted = acn.scenes.load(args.scene).params
    scripted_wh = (scripted.image_width, scripted.image_height)
    n_for_size = world if args.impl == "b200" else args.gpus
    W, H = scaled_size(scripted_wh[0], scripted_wh[1], n_for_size)
    flat = acn.scenes.load(args.scene, image_width=W, image_height=H)
    prm = flat.params

    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args, flat, args.scene, scripted_wh)
        return

    import torch
    import torch.distributed as dist

    acn.device_count()                       # raises without a CUDA device: no CPU fallback
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    os.environ.pop("ACN_PROFILE_KERNELS", None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def over_ranks(vals, op):
        if world == 1:
            return list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return [float(v) for v in t]

    rig = Rig(acn, torch, dist, flat, local_rank, rank, world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # ---- timed region: exactly K steps; per-step CUDA events on the image's stream, L2 flushed in between
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    t_wall0 = time.perf_counter()
    step_ms, stats_by_class = rig.timed(args.steps, args.warmup, flush, barrier)
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop() if rank == 0 else None
    total_ms = float(sum(step_ms))
    launches, rays, n_local = rig.launches, rig.rays, rig.n_local

    # ---- one extra, untimed step with an event pair around every tracing-kernel launch (kernel shares for the roofline)
    os.environ["ACN_PROFILE_KERNELS"] = "1"
    st = rig.step()
    torch.cuda.synchronize(dev)
    os.environ.pop("ACN_PROFILE_KERNELS", None)
    kernel_ms = np.array(list(st.kernel_ms)) * args.steps            # scaled so that the per-step figures below hold
    kernel_cnt = np.array(list(st.kernel_launches_by_class)) * args.steps
    prof_step_ms = st.device_ms
    xy_np = rig.img.pass_xy(n_local)                                 # this rank's pass-0 samples (for the host legs)

    # ---- end to end with host buffers (pinned), same K steps
    h_xy = torch.from_numpy(xy_np).pin_memory()
    h_rgb = torch.empty((n_local, 3), dtype=torch.float32).pin_memory()
    xy_host, rgb_host = h_xy.numpy(), h_rgb.numpy()
    import ctypes as C
    lib = acn.load_library()
    if world == 1:
        def host_step():
            s = acn.Stats()
            rc = lib.acn_render_samples(rig.tracer._p, xy_host.ctypes.data, n_local, 0, rgb_host.ctypes.data, None, C.byref(s))
            if rc:
                raise acn.AcnError(rc, lib.acn_last_error().decode())
            return float(rgb_host[0, 0])          # the result is read on the host
        h2d, d2h = n_local * 16, n_local * 12
    else:
        d_xy = torch.empty((n_local, 2), dtype=torch.float64, device=dev)
        d_rgb = torch.empty((n_local, 3), dtype=torch.float32, device=dev)
        h_img = torch.empty(rig.img.words, dtype=torch.int64).pin_memory()

        def host_step():
            rig.img.reset()
            d_xy.copy_(h_xy, non_blocking=True)
            rig.tracer.render_samples_device(d_xy, d_rgb, stream=rig.stream)      # returns when the samples are traced
            rig.img.accumulate(d_xy, d_rgb, rig.stream)
            rig.img.copy_delta(rig.buf, rig.stream)
            dist.all_reduce(rig.buf)
            h_img.copy_(rig.buf, non_blocking=True)
            rig.stream.synchronize()
            rig.img.reset()                       # the delta of this pass is dropped again
            return int(h_img[5])                  # the reduced image is read on the host
        h2d, d2h = n_local * 16, rig.img.words * 8
    host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- max / sum over ranks
    total_ms, e2e_s = over_ranks([total_ms, e2e_s], dist.ReduceOp.MAX if world > 1 else None)
    n_total, rays_total, launches_total, h2d_total, d2h_total = over_ranks([n_local, rays, launches, h2d, d2h], dist.ReduceOp.SUM if world > 1 else None)

    # ---- other configurations and strong scaling (bounded: a few steps each)
    extras = None
    if not args.no_extras:
        extras = run_extras(acn, torch, dist, args, local_rank, rank, world, barrier, over_ranks)

    if rank != 0:
        rig.close()
        if world > 1:
            dist.destroy_process_group()
        return

    value = n_total * args.steps / (total_ms * 1e-3)
    e2e_value = n_total * args.steps / e2e_s

    # ---- roofline of the dominant kernel + CPU baseline (rank 0, N = 1 only for the CPU leg)
    fp32_peak = acn.measure_fp32_peak_tflops(local_rank)
    class_names = ["k_primary", "k_rays", "k_path", "k_direct", "k_shade", "k_index", "k_sched"]
    dom = int(np.argmax(kernel_ms[:4]))
    roof = {"bound": "fp32", "kernel": class_names[dom], "achieved": None, "peak": fp32_peak, "unit": "TFLOP/s", "frac": None,
            "traffic": None, "traffic_note": "not measured live; per-kernel DRAM bytes of an ncu --set full capture are in profiles/r02_*",
            "peak_source": "measured live: dependent-free FFMA chains on every SM (acn_measure_fp32_peak_tflops)",
            "kernel_ms_per_step": {class_names[i]: kernel_ms[i] / args.steps for i in range(7)},
            "kernel_launches_per_step": {class_names[i]: kernel_cnt[i] / args.steps for i in range(7)},
            "kernel_share_of_step": {class_names[i]: kernel_ms[i] / args.steps / prof_step_ms for i in range(7)},
            "kernel_timing": "CUDA-event pairs around every launch of the tracing kernels in one extra untimed step (kernels serialised)"}
    cpu = None
    if not args.no_cpu_baseline:
        # N > 1: the oracle runs a smaller sample on rank 0 only to count the algorithmic FLOPs of rank 0's share (the
        # roofline is per GPU); the cpu_baseline object itself is an N = 1 deliverable
        from tests.oracle_lib import Oracle
        orc = Oracle()
        threads = os.cpu_count() or 1
        stride = cpu_probe_stride(orc, flat, xy_np, args.cpu_seconds if world == 1 else min(args.cpu_seconds, 4.0), threads)
        xs = cpu_sample(xy_np, stride)
        t0 = time.perf_counter()
        _, info = orc.render(flat, xs, seed_mode=0, threads=threads)
        dt = time.perf_counter() - t0
        scale = n_local / len(xs)
        phase = info["phase_flops"]
        phase_key = ["primary", "rays", "path", "direct"][dom]
        alg_flops_step = phase[phase_key] * scale
        # the phase's surface response runs in k_shade: charge ALL of k_shade's time to the dominant kernel (conservative)
        dom_ms = (kernel_ms[dom] + kernel_ms[4]) / args.steps
        roof["kernel"] = class_names[dom] + " (+ k_shade)"
        roof["achieved"] = alg_flops_step / (dom_ms * 1e-3) / 1e12
        roof["frac"] = roof["achieved"] / fp32_peak
        roof["algorithmic_flops_per_step"] = {k: v * scale for k, v in phase.items()}
        roof["algorithmic_flops_per_launch"] = alg_flops_step / max(kernel_cnt[dom] / args.steps, 1)
        roof["whole_step_tflops"] = info["flops"] * scale / (total_ms / args.steps * 1e-3) / 1e12
        roof["whole_step_frac"] = roof["whole_step_tflops"] / fp32_peak
        roof["sf_ops_per_step"] = info["sf_ops"] * scale
        roof["queue_bytes_per_step"] = 56.0 * rays_total / args.steps / world
        roof["queue_gbs"] = roof["queue_bytes_per_step"] / (total_ms / args.steps * 1e-3) / 1e9
        roof["scope"] = "rank 0 (per GPU)"
        if world == 1:
            cpu = {"value": len(xs) / dt, "unit": UNIT, "cores": threads, "kind": "port", "flags": CPU_FLAGS,
                   "sample": f"{len(xs)} of the {n_local} samples of one step (every {stride}th), FP64 oracle, position-hash seeding",
                   "rays_per_sec": info["rays"] / dt, "oracle_rays_per_sample": info["rays"] / len(xs)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (reference scene script, flattened)",
        "config": workload_config(args.scene, prm, world, scripted_wh),
        "rays_per_sec": rays_total / (total_ms * 1e-3), "rays_per_step": rays_total / args.steps,
        "rays_by_class_per_step": {k: v / args.steps for k, v in stats_by_class.items()} if world == 1 else None,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_total), "d2h_bytes_per_step": int(d2h_total),
                "what": "acn_render_samples with pinned host buffers" if world == 1 else
                        "per rank: H2D of its positions, trace, accumulate, all-reduce of the pass sums, D2H of the reduced image"},
        "gpu_launches": int(launches_total),
        "clocks": clk,
        "roofline": roof,
        "cpu_baseline": cpu,
        "extras": extras,
        "wall_s_timed_region": t_wall,
    }
    print(json.dumps(line), flush=True)
    rig.close()
    if world > 1:
        dist.destroy_process_group()


def run_extras(acn, torch, dist, args, local_rank, rank, world, barrier, over_ranks):
    """Strong scaling of the headline pass and pass 0 of the other BASELINE.json configs, split over the ranks by pixel tile;
    frames of the video dealt to the ranks.  Device-resident, CUDA events, max over ranks, 3 steps (1 for the lamps)."""
    MAX = dist.ReduceOp.MAX if world > 1 else None
    SUM = dist.ReduceOp.SUM if world > 1 else None
    out = {"strong": None, "configs": {}, "video": None,
           "how": "pass 0 (every pixel centre) as scripted, pixel tiles dealt to the ranks, all-reduce of the pass sums inside the timed region; "
                  "ms = max over ranks of the CUDA-event time, mean of the steps"}

    def one(name, ov, steps, warmup):
        flat = acn.scenes.load(name, **ov)
        rig = Rig(acn, torch, dist, flat, local_rank, rank, world)
        ms, _ = rig.timed(steps, warmup, None, barrier)
        (t,) = over_ranks([sum(ms)], MAX)
        n, rays = over_ranks([rig.n_local, rig.rays], SUM)
        rec = {"width": flat.params.image_width, "height": flat.params.image_height, "direct_samples": flat.params.direct_samples,
               "path_samples": flat.params.path_samples, "samples_per_step": int(n), "ms_per_step": t / steps,
               "samples_per_sec": n * steps / (t * 1e-3), "rays_per_sec": rays / (t * 1e-3), "rays_per_step": rays / steps,
               "specialised_kernels": rig.specialised, "steps": steps}
        rig.close()
        return rec

    if world > 1:
        out["strong"] = one(args.scene, {}, 3, 1)
    out["configs"]["many_spheres"] = one("many_spheres", {}, 3, 1)
    out["configs"]["diamond"] = one("diamond", {}, 3, 1)
    out["configs"]["primitives_320x240_ds10_ps0"] = one("primitives", dict(image_width=320, image_height=240, direct_samples=10, path_samples=0), 3, 1)
    out["configs"]["hanging_lamps_in_row_640x360"] = one("hanging_lamps_in_row", dict(image_width=640, image_height=360), 1, 0)

    # C5: frames of diamond_video.acn (the ones committed under scenes/), frame i -> rank i mod N, no communication
    frames = sorted(f[:-4] for f in os.listdir(os.path.join(ROOT, "scenes")) if f.startswith("diamond_video_"))
    mine = [f for i, f in enumerate(frames) if i % world == rank]
    barrier()
    ms, n_v, rays_v = 0.0, 0, 0
    for f in mine:                                  # one tracer at a time; same structure: one compiled module serves all frames
        r = Rig(acn, torch, None, acn.scenes.load(f), local_rank, 0, 1)
        m, _ = r.timed(1, 1)
        ms += m[0]; n_v += r.n_local; rays_v += r.rays
        r.close()
    torch.cuda.synchronize()
    (t,) = over_ranks([ms], MAX)
    n, rays = over_ranks([n_v, rays_v], SUM)
    out["video"] = {"frames": len(frames), "what": "pass 0 of every committed frame of diamond_video.acn (400x300, ds 50, ps 50), frame i on rank i mod N",
                    "ms_total": t, "frames_per_sec": len(frames) / (t * 1e-3), "samples_per_sec": n / (t * 1e-3), "rays_per_sec": rays / (t * 1e-3)}
    return out


if __name__ == "__main__":
    main()
