"""Parity of the CUDA tracer (through the C ABI) with the CPU oracle.  Run on the B200 box: -m gpu.

Bars (BASELINE.json north_star):
 * deterministic direct-only mode, shared index-keyed seeding: per-pixel relative error <= 1e-3.
   - validation precision (f64 on the GPU): EVERY pixel, to 1e-6 — proves the wavefront formulation
     (queues, implicit children, LCG skip-ahead, any-hit shadow rays) computes the reference recursion.
   - product precision (f32): the reference's absolute 1e-6 shell is below FP32 resolution at scene
     scale 10, so the f32 path widens it (DESIGN.md "eps"); the pixels that move by more than 1e-3 are
     silhouettes / terminators / refractive rims.  Their FRACTION is asserted and printed, not hidden.
 * full path tracing: per-channel image mean within 0.5 %.
"""
import os

import numpy as np
import pytest

import actinon_b200 as acn
from tests import scenes_util
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def rel_err(a, ref):
    return (np.abs(a - ref) / np.maximum(np.abs(ref), 1e-2)).max(axis=1)


def full_pass(sc):
    flat = sc.flatten()
    return flat, acn.Image(flat.params.image_width, flat.params.image_height).next_pass(flat.params)


def test_loaded_library_is_the_in_tree_cuda_extension():
    assert os.path.samefile(acn.library_path(), os.path.join(os.path.dirname(acn.__file__), "libactinon_b200.so"))
    assert acn.device_count() >= 1


SCENES = {
    "primitives_c1": lambda: acn.scenes.primitives(320, 240, 10, 0),         # config C1
    "glass_ball": lambda: acn.scenes.glass_ball(160, 120, 8, 0),
    "csg_zoo": lambda: acn.scenes.csg_zoo(160, 120, 6, 0),
    "primitives_path": lambda: acn.scenes.primitives(96, 72, 10, 4),
    "glass_ball_path": lambda: acn.scenes.glass_ball(64, 48, 6, 3),
    "csg_zoo_path": lambda: acn.scenes.csg_zoo(64, 48, 4, 3),
    "textures": lambda: scenes_util.chess_floor_and_ball(8, 96, 72),        # §8 a22: chess / plain texture fields, both projections
}


@pytest.mark.parametrize("name", list(SCENES))
def test_f64_validation_mode_matches_oracle_everywhere(orc, name):
    flat, xy = full_pass(SCENES[name]())
    ref, info = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED, precision=acn.PRECISION_F64))
    rgb = t.render_samples(xy)
    st = t.last_stats
    t.close()
    e = rel_err(rgb, ref)
    print(f"{name}: f64 max rel err {e.max():.2e}; rays gpu {st.rays} oracle {info['rays']}")
    assert e.max() < 1e-6                      # float32 output quantisation only
    assert st.rays == info["rays"]             # the same ray tree, ray for ray
    assert st.rays_shadow == info["counters"]["rays_shadow"] and st.rays_path == info["counters"]["rays_path"]


@pytest.mark.parametrize("name,max_frac_1e3,max_frac_1e2", [
    # limits = twice what is measured (round 2, B200; profiles/r02_parity.json, profiles/r02_c1_probe.txt): C1 0.031 % / 0.021 %,
    # glass_ball 0.20 / 0.02, csg_zoo 0.26 / 0.13, primitives_path 0.07 / 0.06, glass_ball_path 0.42 / 0.33, csg_zoo_path
    # 0.49 / 0.26, textures 0.06 / 0.01 (round 1 allowed 1 - 3.5 % beyond 1e-3)
    ("primitives_c1", 0.0007, 0.0005), ("glass_ball", 0.0040, 0.0006), ("csg_zoo", 0.0052, 0.0030),
    ("primitives_path", 0.0016, 0.0013), ("glass_ball_path", 0.0085, 0.0070), ("csg_zoo_path", 0.0100, 0.0060),
    ("textures", 0.0015, 0.0005),              # + pixels on the edges between chess squares (llrint of an f32 coordinate)
])
def test_f32_product_mode_vs_oracle(orc, name, max_frac_1e3, max_frac_1e2):
    flat, xy = full_pass(SCENES[name]())
    ref, _ = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED)          # reference eps = 1e-6
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED))
    rgb = t.render_samples(xy)
    t.close()
    assert np.isfinite(rgb).all()
    e = rel_err(rgb, ref)
    f3, f2 = float((e > 1e-3).mean()), float((e > 1e-2).mean())
    print(f"{name}: f32 median rel err {np.median(e):.2e}; pixels beyond 1e-3: {f3:.4%}, beyond 1e-2: {f2:.4%}; "
          f"mean gpu {rgb.mean(0)} oracle {ref.mean(0)}")
    assert np.median(e) < 3e-4
    assert f3 <= max_frac_1e3 and f2 <= max_frac_1e2
    assert np.allclose(rgb.mean(0), ref.mean(0), rtol=2e-3)


@pytest.mark.parametrize("name", ["primitives_direct", "primitives_path", "glass_ball", "csg_zoo"])
def test_against_committed_golden_vectors(name):
    from tests.golden.make_golden import CASES
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    flat = CASES[name]().flatten()
    t = acn.Tracer(flat, acn.Options(seed_mode=int(g["seed_mode"]), precision=acn.PRECISION_F64))
    rgb = t.render_samples(g["xy"], index_base=int(g["index_base"]))
    t.close()
    assert rel_err(rgb, g["rgb"]).max() < 1e-6
    t = acn.Tracer(flat, acn.Options(seed_mode=int(g["seed_mode"])))
    rgb = t.render_samples(g["xy"], index_base=int(g["index_base"]))
    t.close()
    assert float((rel_err(rgb, g["rgb"]) > 1e-3).mean()) < 0.03


def test_analytic_scenes_on_gpu():
    for prec in (acn.PRECISION_F32, acn.PRECISION_F64):
        t = acn.Tracer(scenes_util.lamp_over_plane().flatten(), acn.Options(precision=prec))
        rgb = t.render_samples(scenes_util.centre_samples())
        t.close()
        # f32 widens the shell (eps ~ 4e-5 here): (r/(r+eps))^2 = 1 - 8e-4 for the lamp of radius 0.1
        assert np.allclose(rgb[0], np.array([0.8, 0.6, 0.4]) * 10 / 25, rtol=3e-4 if prec == acn.PRECISION_F64 else 1.2e-3)
        t = acn.Tracer(scenes_util.absorbing_slab((0.5, 0.8, 0.9), 2.0).flatten(), acn.Options(precision=prec))
        rgb = t.render_samples(scenes_util.centre_samples())
        t.close()
        assert np.allclose(rgb[0], np.array([0.9, 0.7, 0.5]) * np.array([0.5, 0.8, 0.9]) ** 2, rtol=2e-4)


def test_edge_cases_empty_ragged_and_out_of_frame():
    flat, xy = full_pass(acn.scenes.primitives(64, 48, 6, 2))
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED))
    assert t.render_samples(np.empty((0, 2))).shape == (0, 3)
    one = t.render_samples(xy[777:778], index_base=777)
    allp = t.render_samples(xy)
    assert np.allclose(one[0], allp[777], rtol=1e-5, atol=1e-6)             # batch-invariant
    odd = t.render_samples(xy[5:5 + 1237], index_base=5)                     # ragged size, not a multiple of anything
    assert np.allclose(odd, allp[5:5 + 1237], rtol=1e-5, atol=1e-6)
    far = t.render_samples(np.array([[-500.0, 20.0], [1e4, -1e4]]))          # outside the frame: still a valid camera ray
    assert np.isfinite(far).all()
    t.close()


def test_maximum_size_and_null_arguments_are_rejected_not_attempted():
    """More than 2^31-1 samples per call and null buffers come back as ACN_ERR_INVALID_ARG (-1) from both entry points —
    before any allocation, never as an abort (the reference's only error path is bcore_err_fa, scene.c:1006)."""
    import ctypes as C
    import torch
    flat, xy = full_pass(acn.scenes.primitives(16, 12, 2, 0))
    t = acn.Tracer(flat, acn.Options())
    lib = acn.load_library()
    st = acn.Stats()
    h_xy = np.zeros((4, 2)); h_rgb = np.zeros((4, 3), dtype=np.float32)
    d_xy = torch.zeros((4, 2), dtype=torch.float64, device="cuda"); d_rgb = torch.zeros((4, 3), dtype=torch.float32, device="cuda")
    big = 1 << 31
    assert lib.acn_render_samples(t._p, h_xy.ctypes.data, big, 0, h_rgb.ctypes.data, None, C.byref(st)) == -1
    assert lib.acn_render_samples_device(t._p, d_xy.data_ptr(), big, 0, d_rgb.data_ptr(), None, None, C.byref(st)) == -1
    assert lib.acn_render_samples(t._p, None, 4, 0, h_rgb.ctypes.data, None, C.byref(st)) == -1
    assert lib.acn_render_samples_device(t._p, d_xy.data_ptr(), 4, 0, None, None, None, C.byref(st)) == -1
    assert b"2^31" in lib.acn_last_error() or b"null" in lib.acn_last_error()
    assert np.isfinite(t.render_samples(xy[:4])).all()                     # the handle is still good
    t.close()


def test_cancel_flag_abandons_the_pass_like_sigint_in_the_reference():
    """scene.c:978,1143-1153: workers watch signal_received_g, stop early and the caller discards the pass.  Here the
    caller's `cancel` flag is polled between wavefront iterations; a set flag ends the call with ACN_ERR_CANCELLED (-6),
    and the next call on the same handle renders the pass in full."""
    import ctypes as C
    flat, xy = full_pass(acn.scenes.primitives(64, 48, 6, 5))              # needs ~17 wavefront iterations
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED, wave_budget=4096))
    lib = acn.load_library()
    st = acn.Stats()
    xs = np.ascontiguousarray(xy, dtype=np.float64)
    rgb = np.zeros((len(xs), 3), dtype=np.float32)
    flag = C.c_int(1)
    rc = lib.acn_render_samples(t._p, xs.ctypes.data, len(xs), 0, rgb.ctypes.data, C.cast(C.byref(flag), C.c_void_p), C.byref(st))
    assert rc == -6
    flag.value = 0
    rc = lib.acn_render_samples(t._p, xs.ctypes.data, len(xs), 0, rgb.ctypes.data, C.cast(C.byref(flag), C.c_void_p), C.byref(st))
    assert rc == 0 and np.isfinite(rgb).all() and rgb.max() > 0
    ref = t.render_samples(xy)
    assert np.allclose(rgb, ref, rtol=1e-4, atol=1e-5)
    t.close()


def test_tiny_wave_budget_exercises_the_scheduler(orc):
    """A wave budget far below the ray count forces many pops, stack slicing at exact child budgets and
    multiple primary chunks; the result must not change."""
    flat, xy = full_pass(acn.scenes.primitives(64, 48, 6, 5))
    ref, info = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED, precision=acn.PRECISION_F64, wave_budget=600))
    rgb = t.render_samples(xy)
    st = t.last_stats
    t.close()
    assert rel_err(rgb, ref).max() < 1e-6 and st.rays == info["rays"] and st.waves > 20


def test_position_hash_seeding_is_statistically_equivalent(orc):
    """Reference seeding (scene.c:537) depends on mantissa bits, so f32 and f64 draw different samples;
    image means must still agree (full path tracing bar: 0.5 % per channel)."""
    flat, xy = full_pass(acn.scenes.primitives(160, 120, 10, 6))
    ref, _ = orc.render(flat, xy, seed_mode=acn.SEED_POSITION_HASH)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_POSITION_HASH))
    rgb = t.render_samples(xy)
    t.close()
    d = np.abs(rgb.mean(0) - ref.mean(0)) / ref.mean(0)
    print("position-hash mode channel-mean deviation", d)
    assert (d < 5e-3).all()


def test_device_resident_api_and_accumulation():
    import torch
    flat, xy = full_pass(acn.scenes.primitives(64, 48, 6, 0))
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED))
    host = t.render_samples(xy)
    d_xy = torch.from_numpy(xy).cuda()
    d_rgb = t.render_samples_device(d_xy)
    torch.cuda.synchronize()
    assert np.allclose(d_rgb.cpu().numpy(), host, rtol=1e-5, atol=1e-6)
    acc = torch.zeros((48, 64, 4), dtype=torch.float32, device="cuda")
    t.accumulate_device(d_xy, d_rgb, acc)
    t.accumulate_device(d_xy, d_rgb, acc)
    torch.cuda.synchronize()
    a = acc.cpu().numpy()
    assert np.allclose(a[..., 3], 2.0) and np.allclose(a[..., :3].reshape(-1, 3), 2 * host, rtol=1e-5, atol=1e-6)
    t.close()


def test_full_image_matches_oracle_pass_controller(orc):
    """scene_s_create_image_file: pass 0 + gradient passes; same sample lists, same accumulated image."""
    sc = acn.scenes.primitives(80, 60, 8, 0, gradient_cycles=2)
    flat = sc.flatten()
    img, stats = acn.render_image(sc, options=acn.Options(seed_mode=acn.SEED_INDEX_KEYED, precision=acn.PRECISION_F64))
    ref_img = acn.Image(80, 60)
    base = 0
    while True:
        xy = ref_img.next_pass(flat.params)
        if xy.shape[0] == 0:
            break
        rgb, _ = orc.render(flat, xy, index_base=base, seed_mode=acn.SEED_INDEX_KEYED)
        base += xy.shape[0]
        ref_img.push(xy, rgb.astype(np.float32))
    assert img.cycle == ref_img.cycle == 3
    assert np.allclose(img.average(), ref_img.average(), atol=2e-6)
    assert stats.samples == base
