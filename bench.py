#!/usr/bin/env python
"""bench.py — path samples/s and rays/s of the hot path (lum_machine_s_run, reference src/scene.c:1017) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scene NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

One STEP = one sample pass of the hot path over the image of BASELINE.json configs[1]
(wine_glass.acn exactly as scripted: 400x400, direct_samples 200, path_samples 500, trace_depth 25).
At N = 1 the pass is pass 0 of the reference's controller (every pixel centre, scene.c:1110-1120:
160 000 samples).  At N > 1 the step is N such sample passes (pass 0 + N-1 jittered passes: "image
tiles and sample passes shard across the GPUs") whose samples are dealt to the ranks by 8x8 pixel tiles,
so every rank traces W*H samples per step (weak scaling) and all samples of a pixel stay on one GPU;
the per-pixel float4 accumulators are then summed over the ranks with one NCCL all-reduce, inside the
timed region.

 value    whole-job samples/s with the sample positions already resident in HBM (device-resident C-ABI
          entry point acn_render_samples_device + acn_accumulate_device [+ all-reduce])
 e2e      the same metric through the reference-facing call with HOST buffers
          (acn_render_samples = lum_machine_s_run(scene, lum_arr)); H2D of the positions and D2H of the
          colours are inside the timed region
 roofline FP32-ALU roofline of the dominant kernel: algorithmic FLOPs (oracle counters x SURVEY.md §8d
          constants, reference traversal order) / CUDA-event time of that kernel's launches in the step
 cpu_baseline   the CPU oracle (FP64, position-hash seeding, all host threads) on the same scene
 --impl reference   the CPU oracle alone (the reference itself cannot be built: beth is absent)

Inputs are "synthetic" only in the sense of the contract: the geometry is the reference's own scripted
scene, flattened by our .acn front-end (scenes/*.npz); no dataset is involved.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TILE = 8
METRIC = "path_samples_per_sec"
UNIT = "samples/s"


# --------------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------------
def pass_positions(width: int, height: int, p: int) -> np.ndarray:
    """Sample positions of sample pass p: p = 0 pixel centres (scene.c:1110-1120), p > 0 one jittered
    position per pixel (the shape of a gradient pass over the full image, scene.c:1124-1138)."""
    ys, xs = np.mgrid[0:height, 0:width]
    if p == 0:
        jx = jy = 0.5
    else:
        rng = np.random.default_rng(21943294 + p)
        jx = rng.random((height, width))
        jy = rng.random((height, width))
    return np.stack([(xs + jx).ravel(), (ys + jy).ravel()], axis=1).astype(np.float64)


def rank_samples(width: int, height: int, n_ranks: int, rank: int) -> np.ndarray:
    """Samples of one step owned by `rank`: pixels whose 8x8 tile satisfies (tx + ty) % n_ranks == rank,
    for each of the n_ranks sample passes."""
    out = []
    for p in range(n_ranks):
        xy = pass_positions(width, height, p)
        tx = (xy[:, 0].astype(np.int64)) // TILE
        ty = (xy[:, 1].astype(np.int64)) // TILE
        out.append(xy[(tx + ty) % n_ranks == rank])
    return np.ascontiguousarray(np.concatenate(out, axis=0))


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
    """Every stride-th sample (stride coprime to the image width keeps the subset spread over the image)."""
    return np.ascontiguousarray(xy[::stride])


def cpu_probe_stride(orc, flat, xy, target_s: float, threads: int) -> int:
    """Chooses the sub-sampling stride so that one CPU pass takes about target_s seconds."""
    probe = cpu_sample(xy, 97)
    t0 = time.perf_counter()
    orc.render(flat, probe, seed_mode=0, threads=threads)
    dt = max(time.perf_counter() - t0, 1e-4)
    full = dt * len(xy) / len(probe)
    if full <= target_s:
        return 1
    stride = int(np.ceil(full / target_s))
    while np.gcd(stride, 400) != 1:          # keep it coprime to the row length
        stride += 1
    return stride


def run_reference_arm(args, flat, scene_name):
    """The reference's CPU implementation of the path (restated: oracle/, FP64, all host threads)."""
    from tests.oracle_lib import Oracle
    orc = Oracle()
    prm = flat.params
    W, H = prm.image_width, prm.image_height
    threads = os.cpu_count() or 1
    xy_full = pass_positions(W, H, 0)
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
        "config": workload_config(scene_name, prm, args.gpus),
        "rays_per_sec": rays / dt,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference binary unbuildable here (depends on johsteffens/beth, absent); this is the FP64 CPU oracle "
                "that restates scene.c/compound.c/objects.c, dynamic per-sample scheduling on all host threads",
    }
    print(json.dumps(line), flush=True)


def workload_config(scene_name, prm, n):
    return {
        "workload": f"{scene_name}.acn as scripted: {prm.image_width}x{prm.image_height}, direct_samples {prm.direct_samples}, "
                    f"path_samples {prm.path_samples}, trace_depth {prm.trace_depth}; one step = {n} sample pass(es) over the image "
                    f"({prm.image_width * prm.image_height * n} samples), BASELINE.json configs[1]",
        "samples_per_step": prm.image_width * prm.image_height * n,
        "sharding": "8x8 pixel tiles, (tx+ty) % N; one NCCL all-reduce of float4[W*H] per step" if n > 1 else "single GPU",
        "l2": "ray/task queues (>1 GB per step) exceed the 126 MB L2; a 256 MB buffer is overwritten between timed steps",
    }


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="wine_glass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import actinon_b200 as acn
    flat = acn.scenes.load(args.scene)
    prm = flat.params
    W, H = prm.image_width, prm.image_height

    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args, flat, args.scene)
        return

    import torch
    import torch.distributed as dist

    acn.device_count()                       # raises without a CUDA device: no CPU fallback
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_ranks = world

    xy_np = rank_samples(W, H, n_ranks, rank)
    n_local = xy_np.shape[0]
    tracer = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_POSITION_HASH, precision=acn.PRECISION_F32, device=local_rank))
    d_xy = torch.from_numpy(xy_np).to(dev)
    d_rgb = torch.empty((n_local, 3), dtype=torch.float32, device=dev)
    d_acc = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    os.environ.pop("ACN_PROFILE_KERNELS", None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_step():
        d_acc.zero_()
        tracer.render_samples_device(d_xy, d_rgb, stream=stream)
        tracer.accumulate_device(d_xy, d_rgb, d_acc, stream=stream)
        if world > 1:
            dist.all_reduce(d_acc)
        return tracer.last_stats

    for _ in range(args.warmup):
        device_step()
    barrier()

    # ---- timed region: exactly K steps; per-step CUDA events on the launching stream, L2 flushed in between
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    step_ms = []
    kernel_cnt = np.zeros(4)
    launches = 0
    rays = 0
    stats_by_class = {}
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = device_step()
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        launches += st.kernel_launches + 1 + (1 if world > 1 else 0)
        rays += st.rays
        for k in ("rays_primary", "rays_reflection", "rays_chromatic", "rays_refraction", "rays_path", "rays_shadow", "diffuse_hits"):
            stats_by_class[k] = stats_by_class.get(k, 0) + getattr(st, k)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop() if rank == 0 else None
    total_ms = float(sum(step_ms))

    # ---- one extra, untimed step with an event pair around every tracing-kernel launch (kernel shares for the roofline)
    os.environ["ACN_PROFILE_KERNELS"] = "1"
    st = device_step()
    torch.cuda.synchronize(dev)
    os.environ.pop("ACN_PROFILE_KERNELS", None)
    kernel_ms = np.array(list(st.kernel_ms)) * args.steps            # scaled so that the per-step figures below hold
    kernel_cnt = np.array(list(st.kernel_launches_by_class)) * args.steps
    prof_step_ms = st.device_ms

    # ---- end to end through the reference-facing call with host buffers (pinned), same K steps
    h_xy = torch.from_numpy(xy_np).pin_memory()
    h_rgb = torch.empty((n_local, 3), dtype=torch.float32).pin_memory()
    xy_host = h_xy.numpy()
    rgb_host = h_rgb.numpy()
    import ctypes as C
    lib = acn.load_library()

    def host_step():
        st = acn.Stats()
        rc = lib.acn_render_samples(tracer._p, xy_host.ctypes.data, n_local, 0, rgb_host.ctypes.data, None, C.byref(st))
        if rc:
            raise acn.AcnError(rc, lib.acn_last_error().decode())
        return float(rgb_host[0, 0])          # the result is read on the host

    host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([total_ms, e2e_s, float(n_local), float(rays), float(launches)], dtype=torch.float64, device=dev)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_s = float(tmax[0]), float(tmax[1])
        n_total, rays_total, launches_total = int(tsum[2]), float(tsum[3]), int(tsum[4])
    else:
        n_total, rays_total, launches_total = n_local, float(rays), int(launches)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = n_total * args.steps / (total_ms * 1e-3)
    e2e_value = n_total * args.steps / e2e_s

    # ---- roofline of the dominant kernel + CPU baseline (rank 0, N = 1 only for the CPU leg)
    fp32_peak = acn.measure_fp32_peak_tflops(local_rank)
    class_names = ["k_primary", "k_rays", "k_path", "k_direct", "k_shade", "k_index", "k_sched+k_pop"]
    dom = int(np.argmax(kernel_ms[:4]))
    roof = {"bound": "fp32", "kernel": class_names[dom], "achieved": None, "peak": fp32_peak, "unit": "TFLOP/s", "frac": None,
            "traffic": None, "peak_source": "measured live: dependent-free FFMA chains on every SM (acn_measure_fp32_peak_tflops)",
            "kernel_ms_per_step": {class_names[i]: kernel_ms[i] / args.steps for i in range(7)},
            "kernel_launches_per_step": {class_names[i]: kernel_cnt[i] / args.steps for i in range(7)},
            "kernel_share_of_step": {class_names[i]: kernel_ms[i] / args.steps / prof_step_ms for i in range(7)},
            "kernel_timing": "CUDA-event pairs around every launch of the four tracing kernels in one extra untimed step"}
    # DRAM traffic of the dominant kernel per launch: from the committed ncu capture (dram__bytes_read + write), not live
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        k = tr["kernels"].get(class_names[dom])
        if k:
            roof["traffic"] = k["dram_bytes_per_launch"]
            roof["traffic_source"] = tr["source"] + f"; that launch took {k['launch_us']:.0f} us under ncu"
    except Exception:
        pass
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
          cpu = {"value": len(xs) / dt, "unit": UNIT, "cores": threads, "kind": "port",
                 "sample": f"{len(xs)} of the {n_local} samples of one step (every {stride}th), FP64 oracle, position-hash seeding",
                 "rays_per_sec": info["rays"] / dt, "oracle_rays_per_sample": info["rays"] / len(xs)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (reference scene script, flattened)",
        "config": workload_config(args.scene, prm, world),
        "rays_per_sec": rays_total / (total_ms * 1e-3), "rays_per_step": rays_total / args.steps,
        "rays_by_class_per_step": {k: v / args.steps for k, v in stats_by_class.items()} if world == 1 else None,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_total * 16, "d2h_bytes_per_step": n_total * 12},
        "gpu_launches": launches_total,
        "clocks": clk,
        "roofline": roof,
        "cpu_baseline": cpu,
        "wall_s_timed_region": t_wall,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
