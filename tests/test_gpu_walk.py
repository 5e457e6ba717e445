"""The walk through the traversal records (acn_isect.cuh: scene_query / walk_*_step, acn_tracer.cuh: CullBounds, threaded
records) and the refill kernels (acn_kernels.cuh: k_*_refill) against the CPU oracle.  Run on the B200 box: -m gpu.

The records replace the reference's envelope tests (objects.c:261-266, compound.c:215-244) by tests against tighter balls where
the host can prove that the same rays are accepted; these scenes are built to break that proof if it is wrong:
envelopes that cut through their contents, envelopes smaller than the sphere they belong to, compounds without any envelope,
empty compounds, nested compounds among the root's elements (one element each, compound.c:246-299).
"""
import os

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


def grid(flat, step=1):
    W, H = flat.params.image_width, flat.params.image_height
    ys, xs = np.mgrid[0:H:step, 0:W:step]
    return np.ascontiguousarray(np.stack([(xs + 0.5).ravel(), (ys + 0.5).ravel()], axis=1).astype(np.float64))


def cluster_scene(kind, width=96, height=72, ds=6, ps=3):
    """Spheres in nested compounds over a floor: planes, spheres only -> the prims-only (refill) kernels."""
    sc = acn.Scene()
    sc.set(image_width=width, image_height=height, gamma=0.8, gradient_cycles=0, gradient_samples=2, gradient_threshold=0.03,
           trace_depth=25, trace_min_intensity=0.03, direct_samples=ds, path_samples=ps, max_path_length=2.0,
           background_color=(0.4, 0.5, 0.6))
    sc.set(camera_position=(0.0, -7.0, 3.0), camera_view_direction=(0.0, 7.0, -3.0), camera_top_direction=(0, 0, 1), camera_focal_length=3.5)
    light = sc.create_sphere(0.6).set_radiance(30)
    light.move((-3, -3, 5))
    floor = sc.create_plane().set_material("diffuse_polished").set_color((0.6, 0.6, 0.5))
    floor.move((0, 0, -1))

    def ball(r, at, env=None):
        s = sc.create_sphere(r).set_material("diffuse_polished").set_color((0.9, 0.8, 0.6))
        s.move(at)
        if env == "auto":
            s.set_auto_envelope()
        elif env is not None:
            s.set_envelope(env[0], env[1])
        return s

    inner = []
    for k, (dx, dz) in enumerate([(-1.2, 0.0), (0.0, 0.0), (1.2, 0.0), (-0.6, 0.9), (0.6, 0.9)]):
        lst = sc.create_list()
        for j in range(4):
            at = (dx + 0.28 * (j % 2) - 0.14, 0.28 * (j // 2) - 0.14, dz - 0.5)
            if kind == "small_sphere_envelopes" and j == 1:
                lst.push(ball(0.2, at, env=(at, 0.12)))               # the envelope lies INSIDE its sphere: only rays through it count
            elif kind == "offset_sphere_envelopes" and j == 2:
                lst.push(ball(0.2, at, env=((at[0] + 0.15, at[1], at[2]), 0.22)))   # cuts the sphere
            else:
                lst.push(ball(0.2, at, env="auto" if kind != "no_envelopes" else None))
        c = lst.create_compound()
        if kind == "cutting_envelopes":
            c.set_envelope((dx + 0.1, 0.0, dz - 0.5), 0.33)            # cuts through the cluster: rays that miss it see nothing of the cluster
        elif kind == "auto":
            c.set_auto_envelope()
        inner.append(c)
    outer = sc.create_list()
    for c in inner[:3]:
        outer.push(c)
    oc = outer.create_compound()
    if kind in ("auto", "cutting_envelopes"):
        oc.set_auto_envelope()
    sc.clear()
    sc.push(light)
    sc.push(floor)
    sc.push(oc)                      # a nested compound among the root's elements: ONE element
    sc.push(inner[3])
    sc.push(inner[4])
    if kind == "auto":
        sc.push(sc.create_list().create_compound())      # an empty compound
    return sc


KINDS = ["auto", "no_envelopes", "cutting_envelopes", "small_sphere_envelopes", "offset_sphere_envelopes"]


@pytest.mark.parametrize("kind", KINDS)
def test_f64_walk_equals_the_oracle_ray_for_ray(orc, kind):
    flat = cluster_scene(kind).flatten()
    xy = grid(flat)
    ref, info = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED, precision=acn.PRECISION_F64))
    rgb = t.render_samples(xy)
    st = t.last_stats
    t.close()
    e = rel_err(rgb, ref)
    print(f"{kind}: f64 max rel err {e.max():.2e}, rays {st.rays} / {info['rays']}")
    assert e.max() < 1e-6
    assert st.rays == info["rays"]


@pytest.mark.parametrize("kind", KINDS)
def test_f32_walk_close_to_the_oracle(orc, kind):
    flat = cluster_scene(kind).flatten()
    xy = grid(flat)
    ref, info = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED))
    rgb = t.render_samples(xy)
    st = t.last_stats
    t.close()
    e = rel_err(rgb, ref)
    dm = np.abs(rgb.mean(0) - ref.mean(0)) / ref.mean(0)
    print(f"{kind}: f32 median {np.median(e):.2e}, beyond 1e-3 {(e > 1e-3).mean():.3%}, mean dev {dm}, rays {st.rays} / {info['rays']}")
    assert np.median(e) < 1e-4 and (e > 1e-2).mean() < 0.02
    assert (dm < 5e-3).all()
    assert abs(st.rays - info["rays"]) <= 0.01 * info["rays"]


def test_tight_bounds_change_no_sample(monkeypatch):
    """many_spheres at the scripted sample counts: the records with enclosing balls (and spheres held in their record) give the
    image the reference's envelopes give — the same rays are accepted, the arithmetic of an accepted hit is the same."""
    flat = acn.scenes.load("many_spheres")
    xy = grid(flat, 10)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED))
    a = t.render_samples(xy); ra = t.last_stats.rays
    t.close()
    monkeypatch.setenv("ACN_NO_TIGHT_BOUNDS", "1")
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED))
    b = t.render_samples(xy); rb = t.last_stats.rays
    t.close()
    differ = float((np.abs(a - b).max(axis=1) > 0).mean())
    print(f"many_spheres {len(xy)} samples: {differ:.4%} differ between tight bounds and envelopes, rays {ra} / {rb}")
    assert differ <= 1e-3 and abs(ra - rb) <= 1e-4 * rb
    assert np.allclose(a.mean(0), b.mean(0), rtol=1e-5)


def test_packed_evaluation_programs_change_no_sample(monkeypatch):
    """The lamp's big composite objects (17 - 35 variables): truth tables over stretches of chains against the interpreter
    of the full program."""
    flat = acn.scenes.load("hanging_lamp", image_width=100, image_height=130, direct_samples=6, path_samples=4)
    xy = grid(flat, 2)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED))
    a = t.render_samples(xy)
    t.close()
    monkeypatch.setenv("ACN_NO_PACKED_EVAL", "1")
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED))
    b = t.render_samples(xy)
    t.close()
    assert np.array_equal(a, b)


def test_refill_kernels_are_deterministic_and_budget_independent():
    """Fixed-point accumulation: the per-ray atomics of the refill kernels give bit-identical samples whatever the wave budget."""
    flat = cluster_scene("auto", 64, 48, 8, 6).flatten()
    xy = grid(flat)
    out = []
    for budget in (0, 1 << 12, 1 << 15):
        t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_POSITION_HASH, wave_budget=budget))
        out.append(t.render_samples(xy))
        t.close()
    assert np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[2])


def test_octant_order_changes_no_sample(monkeypatch):
    """many_spheres: the closest-hit walks use a copy of the matter records laid out front to back for the ray's octant.  Inside a
    nested compound the closest hit does not depend on the order of the tests (two hits at exactly the same distance aside)."""
    flat = acn.scenes.load("many_spheres")
    xy = grid(flat, 10)
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED))
    a = t.render_samples(xy); ra = t.last_stats.rays
    t.close()
    monkeypatch.setenv("ACN_NO_OCTANT_ORDER", "1")
    t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED))
    b = t.render_samples(xy); rb = t.last_stats.rays
    t.close()
    differ = float((np.abs(a - b).max(axis=1) > 0).mean())
    print(f"many_spheres {len(xy)} samples: {differ:.4%} differ between octant order and scene order, rays {ra} / {rb}")
    assert differ <= 1e-3 and abs(ra - rb) <= 1e-4 * rb
    assert np.allclose(a.mean(0), b.mean(0), rtol=1e-5)
