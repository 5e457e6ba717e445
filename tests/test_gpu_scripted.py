"""Parity at the SCRIPTED configurations of BASELINE.json (C1-C5 with the scripts' own direct_samples /
path_samples / trace_depth), and full scripted renders against the reference's shipped images.  Run on the B200 box: -m gpu.

 * f32 product mode vs the FP64 oracle, index-keyed seeding, >= 128 x 72 grids of pixel centres of the full-size image;
 * f64 validation mode vs the oracle ray for ray on a 32 x 32 subset (identical ray counts, 1e-6);
 * full renders (all gradient_cycles + 1 passes) of primitives / wine_glass / diamond / many_spheres against
   reference image/*.png (tests/golden/ref_images.npz): channel means within 0.5 %, RMSE no worse than 1.05 x the RMSE of the
   CPU oracle's own full render against the same image (tests/golden/ref_image_stats.json, produced by
   tools/parity_report.py on the GPU box's host cores).
Every test appends its measured figures to gpurun_out/parity/*.json; tools/parity_report.py merges them into
profiles/r02_parity.json.
"""
import json
import os

import numpy as np
import pytest

import actinon_b200 as acn
from tests.oracle_lib import Oracle
from tests.parity_util import (ROOT, SCRIPTED, err_stats, full_render, grid_samples, image_vs_ref, load_ref_stats, ref_image, rel_err)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def record(kind, name, stats):
    d = os.path.join(ROOT, "gpurun_out", "parity")
    os.makedirs(d, exist_ok=True)
    json.dump(stats, open(os.path.join(d, f"{kind}.{name}.json"), "w"), indent=1)


# Per-sample agreement.  BASELINE.json's north star asks for 1e-3 per pixel in the deterministic direct-only mode (C1) and,
# in full path tracing, for channel means within 0.5 % (+ the RMSE criterion, tested on full renders below).  With index-keyed
# seeding the f32 tracer and the FP64 oracle draw the same random numbers, so their samples agree to ~1e-6 in the median; the
# tail is made of samples in which ONE ray of the tree decided differently (a silhouette, a terminator, a refractive rim, one of
# 32 768 tiny spheres).  How large that tail is, is a property of the scene: it is measured with the oracle alone, shell
# thickness 2e-5 (what FP32 can resolve at scene scale 10) against 1e-6, and the f32 tracer must stay within twice that.
DIRECT_ONLY_LIMITS = {"primitives": (4e-4, 3e-4)}          # C1: fraction of samples beyond 1e-3 / 1e-2


@pytest.mark.parametrize("name", list(SCRIPTED))
def test_scripted_config_f32_product_mode_vs_oracle(orc, name):
    ov, nx, ny, frac = SCRIPTED[name]
    flat = acn.scenes.load(name, **ov)
    xy = grid_samples(flat, nx, ny, frac)
    ref, info = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED)
    wide, _ = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED, eps=2e-5)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED))
    rgb = t.render_samples(xy)
    st = t.last_stats
    t.close()
    assert np.isfinite(rgb).all()
    s = err_stats(rgb, ref)
    sw = err_stats(wide, ref)
    p = flat.params
    s.update(scene=name, width=p.image_width, height=p.image_height, direct_samples=p.direct_samples, path_samples=p.path_samples,
             trace_depth=p.trace_depth, rays_gpu=int(st.rays), rays_oracle=int(info["rays"]),
             ray_count_rel_delta=abs(st.rays - info["rays"]) / max(info["rays"], 1),
             **{"oracle_eps_2e-5_vs_1e-6": {k: sw[k] for k in ("median_rel_err", "frac_beyond_1e-3", "frac_beyond_1e-2", "mean_rel_dev")}})
    record("f32", name, s)
    print(f"{name} (scripted ds {p.direct_samples} ps {p.path_samples}, {nx}x{ny} of {p.image_width}x{p.image_height}): median {s['median_rel_err']:.2e}, "
          f"beyond 1e-3 {s['frac_beyond_1e-3']:.3%} (oracle eps 2e-5: {sw['frac_beyond_1e-3']:.3%}), beyond 1e-2 {s['frac_beyond_1e-2']:.3%} "
          f"({sw['frac_beyond_1e-2']:.3%}), mean dev {s['mean_rel_dev']}, rays {st.rays} vs {info['rays']}")
    assert s["median_rel_err"] < 1e-4
    if name in DIRECT_ONLY_LIMITS:
        l3, l2 = DIRECT_ONLY_LIMITS[name]
        assert s["frac_beyond_1e-3"] <= l3 and s["frac_beyond_1e-2"] <= l2
    else:
        assert s["frac_beyond_1e-3"] <= 2.0 * sw["frac_beyond_1e-3"] + 0.02
        assert s["frac_beyond_1e-2"] <= 2.0 * sw["frac_beyond_1e-2"] + 0.02
    assert max(s["mean_rel_dev"]) < 5e-3                      # north star: channel means within 0.5 %
    assert s["ray_count_rel_delta"] < 0.01


@pytest.mark.parametrize("name", list(SCRIPTED))
def test_scripted_config_f64_ray_for_ray(orc, name):
    ov, _, _, frac = SCRIPTED[name]
    flat = acn.scenes.load(name, **ov)
    big = name == "hanging_lamps_in_row"
    xy = grid_samples(flat, 32, 18 if big else 32, frac)
    ref, info = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED, precision=acn.PRECISION_F64, wave_budget=1 << 20))
    rgb = t.render_samples(xy)
    st = t.last_stats
    t.close()
    s = err_stats(rgb, ref)
    s.update(scene=name, rays_gpu=int(st.rays), rays_oracle=int(info["rays"]))
    record("f64", name, s)
    print(f"{name}: f64 max rel err {s['max_rel_err']:.2e}, beyond 1e-6: {s['frac_beyond_1e-6']:.4%}; rays {st.rays} vs {info['rays']}")
    # FMA contraction on the device against -ffp-contract=off in the oracle: one grazing ray in ~10^5 may flip (see
    # test_gpu_configs.py); everything else agrees to 1e-6
    assert s["frac_beyond_1e-6"] <= 2e-3 and s["max_rel_err"] < 1e-2
    assert abs(st.rays - info["rays"]) <= max(2, 1e-4 * info["rays"])


FULL = ["primitives", "wine_glass", "diamond", "many_spheres"]


@pytest.mark.parametrize("name", FULL)
def test_full_scripted_render_vs_the_shipped_reference_image(name):
    """All gradient_cycles + 1 passes as scripted (primitives at the script's own 400x400 / ds 100 defaults), against the
    image the reference ships for that script."""
    flat = acn.scenes.load(name)
    ref8 = ref_image(name)
    assert ref8.shape[:2] == (flat.params.image_height, flat.params.image_width)
    t = acn.Tracer(flat, acn.Options())
    img, n_samples, n_pass = full_render(flat, lambda xy, base: t.render_samples(xy, base))
    t.close()
    s = image_vs_ref(img.average(), ref8)
    s.update(scene=name, passes=n_pass, samples=n_samples)
    cpu = load_ref_stats().get(name)
    if cpu:
        s["rmse_cpu_oracle"] = cpu["rmse"]
    record("full", name, s)
    print(f"{name}: {n_pass} passes, {n_samples} samples; 8-bit means {s['mean8']} shipped {s['mean8_ref']} dev {s['mean_rel_dev']}; "
          f"RMSE vs shipped {s['rmse']:.5f}" + (f" (CPU oracle {cpu['rmse']:.5f})" if cpu else ""))
    assert n_pass == flat.params.gradient_cycles + 1
    if cpu and max(cpu["mean_rel_dev"]) > 2e-3:
        # diamond: the FP64 CPU oracle's own full render sits 0.3-0.5 % below the shipped image in every channel (the image
        # predates the script in the tree, or was rendered with other settings): the device must then agree with the ORACLE's
        # full render to 0.2 %, and stay within 0.75 % of the shipped image
        dev_cpu = np.abs(np.array(s["mean8"]) - np.array(cpu["mean8"])) / np.array(cpu["mean8"])
        s["mean_rel_dev_vs_cpu_oracle"] = [float(v) for v in dev_cpu]
        record("full", name, s)
        assert dev_cpu.max() < 2e-3 and max(s["mean_rel_dev"]) < 7.5e-3
    else:
        assert max(s["mean_rel_dev"]) < 5e-3
    if cpu:
        assert s["rmse"] <= 1.05 * cpu["rmse"]


def test_lamps_in_row_640x360_pass0_means_vs_the_shipped_jpeg():
    """C5: hanging_lamps_in_row at 640x360, scripted ds 30 / ps 30, pass 0, f32 product mode.  The shipped 640x360 image is
    a JPEG of the full-size render scaled down: only its channel means are comparable (2 %)."""
    flat = acn.scenes.load("hanging_lamps_in_row", image_width=640, image_height=360)
    ref8 = ref_image("hanging_lamp02_640_360")
    W, H = 640, 360
    ys, xs = np.mgrid[0:H, 0:W]
    xy = np.ascontiguousarray(np.stack([(xs + 0.5).ravel(), (ys + 0.5).ravel()], axis=1).astype(np.float64))
    t = acn.Tracer(flat, acn.Options())
    rgb = t.render_samples(xy)
    st = t.last_stats
    t.close()
    s = image_vs_ref(rgb.reshape(H, W, 3), ref8)
    s.update(scene="hanging_lamps_in_row", rays=int(st.rays), device_ms=st.device_ms, rays_per_sec=st.rays / (st.device_ms * 1e-3))
    record("full", "hanging_lamps_in_row_640x360_pass0", s)
    print(f"lamps 640x360 pass 0: {st.rays / 1e6:.0f} M rays in {st.device_ms:.0f} ms; means {s['mean8']} shipped {s['mean8_ref']} dev {s['mean_rel_dev']}")
    assert max(s["mean_rel_dev"]) < 0.02
