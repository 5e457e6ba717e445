"""Scenes built through the host scene-description API.

``primitives`` is the hand transcription of the reference's ``src_acn/primitives.acn:20-108``
(SURVEY.md Appendix C); the others are small synthetic scenes that exercise the remaining shape,
CSG, texture and media-transition code paths for the parity tests.  Scripted scenes are loaded
with ``Scene.load_acn``.
"""
from __future__ import annotations

from .api import Scene, rotx, roty, rotz


def _camera(sc: Scene, pos, look_at=(0.0, 0.0, 0.0), focal=4.0):
    sc.set(camera_position=pos,
           camera_view_direction=[look_at[i] - pos[i] for i in range(3)],
           camera_top_direction=(0, 0, 1), camera_focal_length=focal)


def primitives(width=320, height=240, direct_samples=10, path_samples=0, gradient_cycles=0,
               radiance_scale=1.0, background_scale=1.0) -> Scene:
    """config C1: primitives.acn at 320x240, direct_samples=10, path_samples=0."""
    sc = Scene()
    sc.set(threads=30, image_width=width, image_height=height, gamma=1.0,
           gradient_cycles=gradient_cycles, gradient_samples=2, gradient_threshold=0.03,
           trace_depth=25, trace_min_intensity=0.03, direct_samples=direct_samples, path_samples=path_samples,
           max_path_length=1.0, background_color=tuple(0.4 * background_scale for _ in range(3)))
    _camera(sc, (0.0, -10.0, 0.0))

    # create_light( 0.5, 30 ) + vec( 0, -4, 4 )   (primitives.acn:45-51,100)
    light = sc.create_sphere(1.0) * 0.5
    light.set_radiance(30 * radiance_scale)
    light.move((0, -4, 4))

    # create_floor( -1 )   (primitives.acn:53-61)
    floor = sc.create_plane()
    floor.set_material("diffuse_polished").set_color((0.6, 0.4, 0.2)).set_refractive_index(1.2).move((0, 0, -1))

    def mat(o):     # set_material closure (primitives.acn:73-77); commutes with the moves below
        return o.set_material("diffuse_polished").set_color((0.6, 0.7, 0.8))

    sph = mat(sc.create_sphere(0.5))
    el1 = mat(sc.create_ellipsoid(0.3, 0.3, 0.5))
    el2 = mat(sc.create_ellipsoid(0.5, 0.5, 0.3))
    tor = mat(sc.create_torus(0.35, 0.15) * rotx(90))
    cyl = mat(sc.create_cylinder(0.4, 0.4))
    cne = mat(sc.create_cone(0.2, 0.2, 2))
    hyp = mat(sc.create_hyperboloid1(0.2, 0.2, 1))

    s = sc.create_list()                      # primitives.acn:79-94
    s.push(sph)
    s.push(el1 + (0, 0, 1.1))
    s.move((-1.1, 0, 0))
    s.push(tor)
    s.push(el2 + (0, 0, 1.0))
    s.move((-1.0, 0, 0))
    s.push(hyp)
    s.move((-0.9, 0, 0))
    s.push(cne)
    s.move((-1.0, 0, 0))
    s.push(cyl)
    s.move((2.0, 0, 0))

    sc.clear()
    sc.push(light)
    sc.push(floor)
    sc.push(s)
    return sc


def glass_ball(width=160, height=120, direct_samples=8, path_samples=0) -> Scene:
    """Glass sphere with a coincident water core over a chess floor: Fresnel, refraction, media
    transition (compound.c:284-295), exit absorption, chess + plain textures, a mirror."""
    sc = Scene()
    sc.set(image_width=width, image_height=height, gamma=0.9, gradient_cycles=0, gradient_samples=2,
           gradient_threshold=0.03, trace_depth=25, trace_min_intensity=0.03, direct_samples=direct_samples,
           path_samples=path_samples, max_path_length=1e30, background_color=(0.4, 0.5, 0.6))
    _camera(sc, (0.0, -8.0, 3.0))
    light = sc.create_sphere(0.7).set_radiance(35).set_color((1.0, 0.95, 0.9))
    light.move((-3, -2, 5))
    floor = sc.create_plane().set_material("diffuse").move((0, 0, -1))
    floor.set_texture_chess((0.8, 0.8, 0.7), (0.3, 0.2, 0.2), 1.0)
    outer = sc.create_sphere(1.0)
    cover = sc.create_plane()
    inner = outer * 0.9
    shell = (outer & ~inner)
    shell.set_material("glass")
    shell.set_envelope((0, 0, 0), 1.01)
    core = inner & (cover - (0, 0, 0.2))
    core.set_material("water").set_transparency((0.177, 0.2, 0.05))
    core.set_envelope((0, 0, 0), 1.01)
    mirror = sc.create_sphere(0.6).set_material("mirror")
    mirror.move((2.2, 0.5, -0.4))
    blob = sc.create_ellipsoid(0.5, 0.7, 0.4).set_material("diffuse_polished").set_texture_plain((0.9, 0.3, 0.2))
    blob.move((-2.1, 0.3, -0.6))
    sc.clear()
    sc.push(light)
    sc.push(floor)
    sc.push(shell)
    sc.push(core)
    sc.push(mirror)
    sc.push(blob)
    return sc


def csg_zoo(width=160, height=120, direct_samples=6, path_samples=0) -> Scene:
    """CSG composites, an anisotropically scaled object, rough surfaces, a nested enveloped compound and
    two lights."""
    sc = Scene()
    sc.set(image_width=width, image_height=height, gamma=1.0, gradient_cycles=0, gradient_samples=2,
           gradient_threshold=0.03, trace_depth=25, trace_min_intensity=0.03, direct_samples=direct_samples,
           path_samples=path_samples, max_path_length=2.0, background_color=(0.3, 0.35, 0.4))
    _camera(sc, (1.0, -9.0, 4.0))
    l1 = sc.create_sphere(0.5).set_radiance(25)
    l1.move((-3, -3, 5))
    l2 = sc.create_sphere(0.3).set_radiance(12).set_color((1.0, 0.8, 0.6))
    l2.move((4, -2, 3))
    floor = sc.create_plane().set_material("diffuse_polished").set_color((0.6, 0.6, 0.5)).move((0, 0, -1))

    cover = sc.create_plane()
    # cube = intersection of six half-spaces (balanced tree, container.c:376-392)
    faces = sc.create_list([
        cover + (0, 0, 0.5), (cover * rotx(180)) - (0, 0, 0.5),
        (cover * rotx(90)) - (0, 0.5, 0), (cover * rotx(-90)) + (0, 0.5, 0),
        (cover * roty(90)) + (0.5, 0, 0), (cover * roty(-90)) - (0.5, 0, 0)])
    cube = faces.create_inside_composite()
    cube.set_material("diffuse_polished").set_color((0.8, 0.3, 0.2)).set_surface_roughness(0.02)
    cube.rotate(rotz(30)).move((-2.0, 0.0, -0.5))
    cube.set_envelope((-2.0, 0.0, -0.5), 0.9)

    # lens = sphere & sphere, dumbbell = sphere | sphere | cylinder-segment
    lens = (sc.create_sphere(1.0) + (0, 0, 0.6)) & (sc.create_sphere(1.0) - (0, 0, 0.6))
    lens.set_material("glass")
    lens.move((0.2, -0.5, 0.2))
    bar = (sc.create_cylinder(0.15, 0.15) & (cover + (0, 0, 0.8))) & ~(cover - (0, 0, 0.8))
    bell = (bar | (sc.create_sphere(0.4) + (0, 0, 0.8))) | (sc.create_sphere(0.4) - (0, 0, 0.8))
    bell.set_material("gold")
    bell.rotate(roty(70)).move((2.2, 0.8, -0.2))
    bell.set_auto_envelope()

    egg = sc.create_sphere(0.5).set_material("diffuse").set_color((0.3, 0.7, 0.4)).set_surface_roughness(0.05)
    egg = egg.scaled_by_vec((1.0, 0.6, 1.6))
    egg.move((-0.6, 1.8, -0.2))

    ring = sc.create_torus(0.5, 0.12).set_material("silver")
    ring.rotate(rotx(60)).move((1.0, -2.0, -0.4))

    beads = sc.create_list()
    for i in range(4):
        b = sc.create_sphere(0.18).set_material("diffuse_polished").set_color((0.2 + 0.2 * i, 0.4, 0.9 - 0.2 * i))
        b.move((-3.2 + 0.45 * i, -2.0, -0.82))
        beads.push(b)
    cmp = beads.create_compound()
    cmp.set_auto_envelope()

    sc.clear()
    for o in (l1, l2, floor, cube, lens, bell, egg, ring, cmp):
        sc.push(o)
    return sc


def load(name: str, **overrides):
    """Loads scenes/<name>.npz (a scripted reference scene flattened by tools/make_scenes.py) and applies
    render-parameter overrides (image_width=..., direct_samples=..., ...).  Returns a FlatScene."""
    import os
    from . import api
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    flat = api.load_flat(os.path.join(root, "scenes", name + ".npz"))
    prm = flat.params
    for k, v in overrides.items():
        cur = getattr(prm, k)
        if hasattr(cur, "__len__"):
            for i in range(len(cur)):
                cur[i] = float(v[i])
        else:
            setattr(prm, k, v)
    return flat
