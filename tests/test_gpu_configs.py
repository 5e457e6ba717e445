"""Parity on the scenes of BASELINE.json's configs (the reference's own scripts, flattened): C2 wine_glass,
C3 many_spheres, C4 diamond, C5 hanging lamps + a diamond_video frame.  Run on the B200 box: -m gpu.

 * f64 validation mode (reference march on the device) vs the CPU oracle: every sample to 1e-6, identical ray counts;
 * the event-sweep CSG evaluator in f64 vs the oracle's march: the two algorithms must find the same first boundary;
 * f32 product mode vs the oracle: bounded outlier fractions, means within 1 %;
 * full-size pass 0 vs the channel means of the reference's shipped images, and size-independent
   properties at full size: batch invariance, run-to-run determinism up to atomic order, linearity in the radiance.
"""
import numpy as np
import pytest

import actinon_b200 as acn
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def rel_err(a, ref):
    return (np.abs(a - ref) / np.maximum(np.abs(ref), 1e-2)).max(axis=1)


def grid_samples(flat, nx, ny, frac=1.0):
    """nx*ny sample positions spread over the central `frac` of the full-size image (pixel centres)."""
    W, H = flat.params.image_width, flat.params.image_height
    xs = (W * (0.5 - frac / 2) + (np.arange(nx) + 0.5) * W * frac / nx).astype(int) + 0.5
    ys = (H * (0.5 - frac / 2) + (np.arange(ny) + 0.5) * H * frac / ny).astype(int) + 0.5
    gx, gy = np.meshgrid(xs, ys)
    return np.ascontiguousarray(np.stack([gx.ravel(), gy.ravel()], axis=1).astype(np.float64))


#            name                      overrides                                   nx  ny  frac
CASES = {
    "wine_glass":           (dict(direct_samples=20, path_samples=10),             64, 64, 0.9),
    "many_spheres":         (dict(direct_samples=6, path_samples=4),               64, 64, 0.9),
    "diamond":              (dict(direct_samples=10, path_samples=6),              64, 64, 0.5),
    "diamond_video_000049": (dict(direct_samples=6, path_samples=4),               48, 36, 0.6),
    "hanging_lamp":         (dict(direct_samples=4, path_samples=3),               40, 52, 0.9),
    "paraffin_lamp":        (dict(direct_samples=5, path_samples=3),               40, 60, 0.9),
}


def load_case(name):
    ov, nx, ny, frac = CASES[name]
    flat = acn.scenes.load(name, **ov)
    return flat, grid_samples(flat, nx, ny, frac)


@pytest.mark.parametrize("name", list(CASES))
def test_config_scene_f64_validation_matches_oracle(orc, name):
    flat, xy = load_case(name)
    ref, info = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED, precision=acn.PRECISION_F64, wave_budget=1 << 18))
    rgb = t.render_samples(xy)
    st = t.last_stats
    t.close()
    e = rel_err(rgb, ref)
    bad = float((e > 1e-6).mean())
    print(f"{name}: f64 max rel err {e.max():.2e}, samples beyond 1e-6: {bad:.4%}; rays gpu {st.rays} oracle {info['rays']}")
    # The device contracts a*b+c into FMAs, the oracle is built with -ffp-contract=off: among ~10^5 rays through
    # 32 768 tiny spheres (or 200-step distance-field marches) a grazing ray may flip.  One in a thousand samples
    # may therefore differ (by one shadow sample: far below the 1e-3 bar); everything else is exact.
    assert bad <= 1e-3 and e.max() < 1e-3
    assert abs(st.rays - info["rays"]) <= max(2, 1e-4 * info["rays"])


@pytest.mark.parametrize("name", ["wine_glass", "diamond", "diamond_video_000049"])
def test_event_sweep_equals_the_reference_march_in_f64(orc, name):
    """csg_eval (variables + crossings + truth table) against the oracle's alternating march, both in double:
    same boundary, same normal, so the same image up to the rare ray that grazes an edge of two facets."""
    flat, xy = load_case(name)
    ref, info = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED, precision=acn.PRECISION_F64, csg_mode=acn.CSG_INTERVALS,
                                     wave_budget=1 << 18))
    rgb = t.render_samples(xy)
    st = t.last_stats
    t.close()
    e = rel_err(rgb, ref)
    bad = float((e > 1e-6).mean())
    print(f"{name}: sweep-vs-march samples beyond 1e-6: {bad:.4%}, max {e.max():.2e}; rays {st.rays} vs {info['rays']}")
    assert bad <= 0.01
    assert abs(st.rays - info["rays"]) <= 0.01 * info["rays"]


@pytest.mark.parametrize("name", ["hanging_lamp", "paraffin_lamp"])
def test_event_sweep_with_distance_field_leaves_in_f64(orc, name):
    """Chain links and lamp parts are CSG over sphere-traced tori (objects.c:903-959).  The sweep takes a torus as a leaf
    with up to four crossings, each found by the reference's own march restarted behind the previous crossing; the
    reference's alternating march restarts from other points, and a sphere-traced crossing depends on where the march
    started by up to the shell thickness (1e-6).  So the two agree to ~1e-5 instead of 1e-6, except at grazing rays."""
    flat, xy = load_case(name)
    ref, info = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED, precision=acn.PRECISION_F64, csg_mode=acn.CSG_INTERVALS,
                                     wave_budget=1 << 18))
    rgb = t.render_samples(xy)
    st = t.last_stats
    t.close()
    e = rel_err(rgb, ref)
    b5, b3 = float((e > 1e-5).mean()), float((e > 1e-3).mean())
    dm = np.abs(rgb.mean(0) - ref.mean(0)) / ref.mean(0)
    print(f"{name}: sweep(with tori)-vs-march samples beyond 1e-5: {b5:.4%}, beyond 1e-3: {b3:.4%}, max {e.max():.2e}; "
          f"rays {st.rays} vs {info['rays']}; mean dev {dm}")
    assert b5 <= 0.05 and b3 <= 0.02
    assert (dm < 2e-3).all()
    assert abs(st.rays - info["rays"]) <= 0.01 * info["rays"]


@pytest.mark.parametrize("name", list(CASES))
def test_config_scene_f32_product_mode(orc, name):
    """FP32 cannot hold the reference's absolute 1e-6 shell at scene scale 10 (DESIGN.md "eps"), so the product path
    widens it to ~2e-5.  How far a sample may move under that change is a property of the SCENE (chaotic paths through
    tiny spheres and facets) and is measured here with the FP64 oracle alone, eps 2e-5 against eps 1e-6; the f32 tracer
    must stay within twice that, and the image means within 1 %."""
    flat, xy = load_case(name)
    ref, _ = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED)
    wide, _ = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED, eps=2e-5)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED, wave_budget=1 << 18))
    rgb = t.render_samples(xy)
    t.close()
    assert np.isfinite(rgb).all()
    e, ew = rel_err(rgb, ref), rel_err(wide, ref)
    f3, f2 = float((e > 1e-3).mean()), float((e > 1e-2).mean())
    w3, w2 = float((ew > 1e-3).mean()), float((ew > 1e-2).mean())
    dm = np.abs(rgb.mean(0) - ref.mean(0)) / ref.mean(0)
    print(f"{name}: f32 median rel err {np.median(e):.2e}; beyond 1e-3: {f3:.3%} (oracle eps 2e-5: {w3:.3%}), "
          f"beyond 1e-2: {f2:.3%} ({w2:.3%}); mean dev {dm}")
    assert np.median(e) < 1e-4
    # (paraffin_lamp used to need a slack of 0.05 here: the sweep mis-ordered the two crossings of a ray through the
    # common cut plane of two butted pieces of the liquid; fixed in csg_eval, "group of crossings")
    slack = 0.02
    assert f3 <= 2.0 * w3 + slack and f2 <= 2.0 * w2 + slack
    assert (dm < 1e-2).all()


# channel means of the reference's shipped full-quality renders (measured with PIL)
SHIPPED = {
    "primitives":   ((400, 400), (0.4701, 0.4482, 0.4257)),
    "wine_glass":   ((400, 400), (0.6944, 0.6019, 0.4722)),
    "many_spheres": ((600, 600), (0.4378, 0.4754, 0.4552)),
    "diamond":      ((400, 400), (0.5713, 0.4415, 0.3342)),
    "diamond_video_000049": ((400, 300), (0.3650, 0.2463, 0.1883)),
}


def full_pass0(flat):
    W, H = flat.params.image_width, flat.params.image_height
    ys, xs = np.mgrid[0:H, 0:W]
    return np.ascontiguousarray(np.stack([(xs + 0.5).ravel(), (ys + 0.5).ravel()], axis=1).astype(np.float64))


@pytest.mark.parametrize("name", list(SHIPPED))
def test_full_size_pass0_agrees_with_the_shipped_reference_image(name):
    """The only artefacts of the reference that pin this path are its shipped renders: pass 0 at the scripted size
    and sample counts must reproduce their channel means (the adaptive passes only refine edges)."""
    flat = acn.scenes.load(name)
    (W, H), want = SHIPPED[name]
    assert (flat.params.image_width, flat.params.image_height) == (W, H)
    t = acn.Tracer(flat, acn.Options())
    rgb = t.render_samples(full_pass0(flat))
    t.close()
    got = rgb.mean(0)
    dev = np.abs(got - np.array(want)) / np.array(want)
    print(f"{name}: pass-0 means {got} shipped {want} deviation {dev}")
    assert (dev < 0.012).all()


def test_full_size_batch_invariance_and_determinism():
    """Position-hash seeding makes a sample a pure function of its position: a sample traced alone, inside the full
    pass, or in a second run must agree up to the order of the FP32 atomic accumulation."""
    flat = acn.scenes.load("wine_glass", direct_samples=40, path_samples=30)
    xy = full_pass0(flat)
    t = acn.Tracer(flat, acn.Options())
    a = t.render_samples(xy)
    b = t.render_samples(xy)
    pick = np.random.default_rng(5).choice(len(xy), 3000, replace=False)
    c = t.render_samples(xy[pick])
    t.close()
    assert np.allclose(a, b, rtol=2e-4, atol=2e-5)
    assert np.allclose(c, a[pick], rtol=2e-4, atol=2e-5)


def test_linearity_in_the_radiance():
    """scene_s_lum is linear in the light's radiance: with gamma 1, doubling it doubles every unclamped sample."""
    sc1 = acn.scenes.primitives(160, 120, 6, 3)
    flat1 = sc1.flatten()
    xy = acn.Image(160, 120).next_pass(flat1.params)
    t = acn.Tracer(flat1, acn.Options(seed_mode=acn.SEED_INDEX_KEYED))
    r1 = t.render_samples(xy)
    t.close()
    sc2 = acn.scenes.primitives(160, 120, 6, 3, radiance_scale=0.5, background_scale=0.5)
    flat2 = sc2.flatten()
    t = acn.Tracer(flat2, acn.Options(seed_mode=acn.SEED_INDEX_KEYED))
    r2 = t.render_samples(xy)
    t.close()
    ok = (r1 < 0.999).all(axis=1) & (r1 > 1e-3).all(axis=1)
    # halving changes the integer sample counts of nothing (they depend on intensities, not on radiance)
    assert np.allclose(r2[ok] * 2.0, r1[ok], rtol=2e-3, atol=1e-5)


def _jittered(W, H, spp, seed):
    """spp jittered sample positions per pixel, raster order inside each sample layer (the shape of the reference's
    gradient passes, scene.c:1124-1138)."""
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:H, 0:W]
    out = [np.stack([(xs + rng.random((H, W))).ravel(), (ys + rng.random((H, W))).ravel()], axis=1) for _ in range(spp)]
    return np.ascontiguousarray(np.concatenate(out, axis=0).astype(np.float64))


def _image(xy, rgb, W, H):
    img = np.zeros((H, W, 3)); cnt = np.zeros((H, W, 1))
    x, y = xy[:, 0].astype(int), xy[:, 1].astype(int)
    np.add.at(img, (y, x), rgb); np.add.at(cnt, (y, x), 1.0)
    return img / cnt


@pytest.mark.parametrize("name,ov", [
    ("wine_glass", dict(image_width=40, image_height=40, direct_samples=12, path_samples=8)),
    ("diamond",    dict(image_width=40, image_height=40, direct_samples=8, path_samples=6)),
])
def test_path_tracing_rmse_against_a_converged_reference_is_no_worse_than_the_cpu_renders(orc, name, ov):
    """BASELINE.json north_star, full path tracing: "per-channel mean within 0.5 % and RMSE against a high-sample
    converged reference no worse than the CPU render's at equal samples".  Converged reference = the FP64 CPU oracle at
    48 jittered samples per pixel; contenders = the f32 CUDA tracer and the oracle itself at 4 samples per pixel on the
    same sample positions.  The position-hash seeds are taken from the bits of the hit position, which differ between f32
    and f64, so the two renders carry INDEPENDENT Monte-Carlo noise: two CPU renders with different jitter differ by
    3-6 % in RMSE at this size, hence the 10 % allowance.  The channel means are compared at 48 samples per pixel,
    where the noise of the mean (~0.1 %) is well below the 0.5 % criterion."""
    flat = acn.scenes.load(name, **ov)
    W, H = flat.params.image_width, flat.params.image_height
    xy_ref = _jittered(W, H, 48, 1)
    ref = _image(xy_ref, orc.render(flat, xy_ref, seed_mode=acn.SEED_POSITION_HASH)[0], W, H)
    xy = _jittered(W, H, 4, 2)
    cpu = _image(xy, orc.render(flat, xy, seed_mode=acn.SEED_POSITION_HASH)[0], W, H)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_POSITION_HASH, wave_budget=1 << 18))
    gpu = _image(xy, t.render_samples(xy), W, H)
    gpu48 = _image(xy_ref, t.render_samples(xy_ref), W, H)
    t.close()
    rmse_cpu = float(np.sqrt(((cpu - ref) ** 2).mean()))
    rmse_gpu = float(np.sqrt(((gpu - ref) ** 2).mean()))
    dm = np.abs(gpu48.mean((0, 1)) - ref.mean((0, 1))) / ref.mean((0, 1))
    print(f"{name}: RMSE vs 48-spp reference: cpu {rmse_cpu:.5f}, gpu {rmse_gpu:.5f}; channel-mean deviation at 48 spp {dm}")
    assert rmse_gpu <= 1.10 * rmse_cpu
    assert (dm < 5e-3).all()


ALL_SCENES = ["caustic_of_caustic", "diamond", "diamond_video_000049", "hanging_lamp", "hanging_lamps_in_row", "many_spheres",
              "paraffin_lamp", "paraffin_lamp_on_ledge", "primitives", "pyramid", "ruby_heart", "wine_glass"]


@pytest.mark.parametrize("name", ALL_SCENES)
def test_event_sweep_equals_the_reference_march_on_every_shipped_scene(orc, name):
    """Every scene script the reference ships (flattened in scenes/), reduced sample counts, a grid of samples: the f64
    event sweep (csg_mode INTERVALS: variables, crossings, truth tables, convex runs, envelope gates, distance-field
    leaves, groups of coincident crossings) must find the boundaries the oracle's alternating march finds."""
    big = name == "hanging_lamps_in_row"
    flat = acn.scenes.load(name, direct_samples=3 if big else 6, path_samples=2 if big else 4)
    xy = grid_samples(flat, 24 if big else 48, 24 if big else 48, 0.9)
    ref, info = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED, precision=acn.PRECISION_F64, csg_mode=acn.CSG_INTERVALS,
                                     wave_budget=1 << 18))
    rgb = t.render_samples(xy)
    st = t.last_stats
    t.close()
    e = rel_err(rgb, ref)
    b5, b3 = float((e > 1e-5).mean()), float((e > 1e-3).mean())
    print(f"{name}: sweep-vs-march samples beyond 1e-5: {b5:.4%}, beyond 1e-3: {b3:.4%}; rays {st.rays} vs {info['rays']}")
    assert b5 <= 0.005 and b3 <= 0.002
    assert abs(st.rays - info["rays"]) <= max(8, 1e-3 * info["rays"])
