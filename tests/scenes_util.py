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
