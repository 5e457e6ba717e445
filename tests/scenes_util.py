"""Small analytic scenes shared by the oracle tests (CPU) and the CUDA parity tests (GPU)."""
import numpy as np

import actinon_b200 as acn


def lamp_over_plane(albedo=(0.8, 0.6, 0.4), radiance=10.0, height=5.0, lamp_radius=0.1, direct_samples=16):
    """Diffuse (Lambertian: sigma = 0) plane under a small spherical lamp, seen from straight above.
    Closed form at the foot point: albedo * radiance / height^2 (the 1/r^2 of the lamp surface cancels
    the cap solid angle, scene.c:572-579)."""
    sc = acn.Scene()
    sc.set(image_width=4, image_height=4, gamma=1.0, trace_depth=11, trace_min_intensity=0.01,
           direct_samples=direct_samples, path_samples=0, background_color=(0, 0, 0),
           camera_position=(0, 0, 2.0), camera_view_direction=(0, 0, -1), camera_top_direction=(0, 1, 0),
           camera_focal_length=50.0)
    lamp = sc.create_sphere(lamp_radius).set_radiance(radiance).set_color((1, 1, 1))
    lamp.move((0, 0, height))
    floor = sc.create_plane().set_material("diffuse").set_sigma(0.0).set_color(albedo)
    sc.clear(); sc.push(lamp); sc.push(floor)
    return sc


def absorbing_slab(transparency=(0.5, 0.8, 0.9), thickness=2.0):
    """A slab of index-1 absorbing medium in front of a uniform background: Beer-Lambert t^thickness
    (scene.c:656-664); n = 1 switches surface reflection off (scene.c:451)."""
    sc = acn.Scene()
    sc.set(image_width=4, image_height=4, gamma=1.0, trace_depth=11, trace_min_intensity=0.01,
           direct_samples=1, path_samples=0, background_color=(0.9, 0.7, 0.5),
           camera_position=(0, -10, 0), camera_view_direction=(0, 1, 0), camera_top_direction=(0, 0, 1),
           camera_focal_length=50.0)
    cover = sc.create_plane()
    front = (cover * acn.api.rotx(90)) - (0, 0, 0)            # normal (0,-1,0) at y = 0
    back = (cover * acn.api.rotx(-90)) + (0, thickness, 0)    # normal (0,+1,0) at y = thickness
    slab = front & back
    slab.set_material("transparent").set_transparency(transparency)
    sc.clear(); sc.push(slab)
    return sc


def centre_samples(n=4):
    return np.array([[n / 2, n / 2]], dtype=np.float64)


def chess_floor_and_ball(direct_samples=8, width=48, height=36):
    """no shipped scene sets a texture field, so obj_color / txm_chess_s_clr (textures.c:142-148) and
    the two projections (plane u,v: objects.c:514-518; sphere azimuth, elevation: objects.c:602-617) are exercised by a
    synthetic scene: a chess floor (squares of 1/scale) under a lamp, a chess ball and a plain-textured ball."""
    sc = acn.Scene()
    sc.set(image_width=width, image_height=height, gamma=1.0, trace_depth=11, trace_min_intensity=0.01,
           direct_samples=direct_samples, path_samples=0, background_color=(0.1, 0.1, 0.1),
           camera_position=(0, -6.0, 2.5), camera_view_direction=(0, 1, -0.35), camera_top_direction=(0, 0, 1),
           camera_focal_length=3.0)
    lamp = sc.create_sphere(0.3).set_radiance(8.0).set_color((1, 1, 1))
    lamp.move((1.0, -2.0, 5.0))
    floor = sc.create_plane().set_material("diffuse").set_sigma(0.0).set_color((1, 1, 1))
    floor.set_texture_chess((0.9, 0.1, 0.1), (0.1, 0.1, 0.9), 1.0)
    ball = sc.create_sphere(0.8).set_material("diffuse").set_sigma(0.0).set_color((1, 1, 1))
    ball.set_texture_chess((0.9, 0.9, 0.1), (0.1, 0.6, 0.1), 4.0)
    ball.move((-1.0, 0.5, 0.8))
    plain = sc.create_sphere(0.6).set_material("diffuse").set_sigma(0.0).set_color((0.2, 0.2, 0.2))
    plain.set_texture_plain((0.3, 0.8, 0.8))
    plain.move((1.3, 0.2, 0.6))
    sc.clear(); sc.push(lamp); sc.push(floor); sc.push(ball); sc.push(plain)
    return sc
