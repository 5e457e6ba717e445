"""The oracle's render loop: analytic known answers (SURVEY.md §8c ii), golden fixtures, counters."""
import os

import numpy as np
import pytest

import actinon_b200 as acn
from tests import scenes_util
from tests.oracle_lib import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def test_lamp_over_plane_closed_form(orc):
    sc = scenes_util.lamp_over_plane()
    rgb, info = orc.render(sc.flatten(), scenes_util.centre_samples(), seed_mode=0)
    # lamp radius 0.1 at height 5: cos(theta) >= 0.9998 over the cap and (r/(r+eps))^2 = 1 - 2e-5
    expect = np.array([0.8, 0.6, 0.4]) * 10.0 / 5.0 ** 2
    assert np.allclose(rgb[0], expect, rtol=2e-4)
    assert info["counters"]["rays_shadow"] == 2 * 16 and info["counters"]["rays_primary"] == 1


def test_inverse_square_and_linearity(orc):
    a, _ = orc.render(scenes_util.lamp_over_plane(height=5.0, radiance=10).flatten(), scenes_util.centre_samples())
    b, _ = orc.render(scenes_util.lamp_over_plane(height=10.0, radiance=10).flatten(), scenes_util.centre_samples())
    c, _ = orc.render(scenes_util.lamp_over_plane(height=5.0, radiance=20).flatten(), scenes_util.centre_samples())
    assert np.allclose(a / b, 4.0, rtol=1e-4) and np.allclose(c / a, 2.0, rtol=1e-12)


def test_beer_lambert(orc):
    t = (0.5, 0.8, 0.9)
    sc = scenes_util.absorbing_slab(t, thickness=2.0)
    rgb, info = orc.render(sc.flatten(), scenes_util.centre_samples())
    expect = np.array([0.9, 0.7, 0.5]) * np.array(t) ** 2.0
    assert np.allclose(rgb[0], expect, rtol=1e-5)
    assert info["counters"]["rays_refract"] == 2 and info["counters"]["absorb"] == 1


def test_seed_modes_agree_statistically_and_index_mode_is_order_independent(orc):
    sc = acn.scenes.primitives(48, 36, direct_samples=10, path_samples=0)
    flat = sc.flatten()
    xy = acn.Image(48, 36).next_pass(flat.params)
    a, _ = orc.render(flat, xy, seed_mode=0)
    b, _ = orc.render(flat, xy, seed_mode=1)
    assert abs(a.mean() - b.mean()) / a.mean() < 5e-3
    # index-keyed: a sample's colour depends on (index_base + i) only, not on batch composition or threads
    c, _ = orc.render(flat, xy[100:200], index_base=100, seed_mode=1, threads=1)
    assert np.array_equal(b[100:200], c)
    # position-hash mode is a pure function of the geometry: batch-invariant too
    d, _ = orc.render(flat, xy[100:200], index_base=12345, seed_mode=0, threads=3)
    assert np.array_equal(a[100:200], d)


@pytest.mark.parametrize("name", ["primitives_direct", "primitives_path", "glass_ball", "csg_zoo"])
def test_golden(orc, name):
    """Regression pin of the oracle itself (fixtures made by tests/golden/make_golden.py)."""
    from tests.golden.make_golden import CASES
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sc = CASES[name]()
    rgb, _ = orc.render(sc.flatten(), g["xy"], index_base=int(g["index_base"]), seed_mode=int(g["seed_mode"]))
    assert np.allclose(rgb, g["rgb"], rtol=0, atol=1e-12)


def test_counters_give_flops(orc):
    sc = acn.scenes.primitives(32, 24, direct_samples=10, path_samples=0)
    flat = sc.flatten()
    xy = acn.Image(32, 24).next_pass(flat.params)
    _, info = orc.render(flat, xy)
    c = info["counters"]
    assert c["rays_primary"] == 32 * 24 and c["camera"] == 32 * 24
    assert info["flops"] > 1000 * 32 * 24 and info["rays"] > 5 * 32 * 24
    assert c["dist_step"] > 0 and c["squaroid"] > 0 and c["plane"] > 0 and c["oren_nayar"] > 0


def test_chess_and_plain_textures(orc):
    """obj_color with a texture field (objects.c:411-422, textures.c:99-102,142-148): the colour of a diffuse hit is the
    texture's, not prp.color.  Two floor points in neighbouring squares under the same lighting geometry differ exactly
    by the ratio of the two chess colours; the plain-textured ball shows its texture colour's hue."""
    sc = scenes_util.chess_floor_and_ball()
    flat = sc.flatten()
    W, H = 48, 36
    ys, xs = np.mgrid[0:H, 0:W]
    xy = np.stack([(xs + 0.5).ravel(), (ys + 0.5).ravel()], axis=1).astype(np.float64)
    rgb, info = orc.render(flat, xy, seed_mode=1, want_linear=True)
    img = info["linear"].reshape(H, W, 3)                                   # before gamma and the clamp to [0,1]
    np.seterr(divide="ignore", invalid="ignore")                          # shadowed pixels are 0/0 in the ratios below
    # every floor pixel is a multiple of one of the two chess colours (red-ish (9,1,1) or blue-ish (1,1,9))
    floor_rows = img[H - 6:, :, :].reshape(-1, 3)
    r_over_b = floor_rows[:, 0] / floor_rows[:, 2]
    is_red = np.isclose(r_over_b, 9.0, rtol=1e-6); is_blue = np.isclose(r_over_b, 1.0 / 9.0, rtol=1e-6)
    assert (is_red | is_blue).all() and is_red.any() and is_blue.any()
    # plain texture: hue of (0.3, 0.8, 0.8) wherever the ball is hit directly, never the grey prp.color
    g_over_r = img[:, :, 1] / img[:, :, 0]
    assert np.isclose(g_over_r, 0.8 / 0.3, rtol=1e-6).sum() > 10
    # chess ball: both of its colours occur
    assert np.isclose(img[:, :, 0] / img[:, :, 2], 9.0, rtol=1e-6).sum() > 5          # (0.9, 0.9, 0.1)
    assert np.isclose(img[:, :, 1] / img[:, :, 0], 6.0, rtol=1e-6).sum() > 5          # (0.1, 0.6, 0.1)
