"""Shared by tests/test_gpu_scripted.py and tools/parity_report.py: the scripted-config cases of SURVEY.md §8 (C1-C5),
error statistics, full scripted renders (all gradient passes) and their comparison with the reference's shipped images."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

#  name: (scene overrides, grid nx, ny, central fraction of the image the grid covers)
# direct_samples / path_samples / trace depth are the SCRIPTED ones (BASELINE.json configs)
SCRIPTED = {
    "primitives":           (dict(image_width=320, image_height=240, direct_samples=10, path_samples=0), 160, 120, 1.0),   # C1
    "wine_glass":           (dict(),                                   128, 128, 0.9),    # C2: ds 200, ps 500
    "many_spheres":         (dict(),                                   128, 128, 0.9),    # C3: ds 20, ps 20
    "diamond":              (dict(),                                   128, 128, 0.6),    # C4: ds 50, ps 50
    "diamond_video_000049": (dict(),                                   128, 96, 0.6),     # C5: frame 49 of the video
    "hanging_lamps_in_row": (dict(image_width=640, image_height=360),  128, 72, 0.95),    # C5: ds 30, ps 30 at 640x360
}


def grid_samples(flat, nx, ny, frac=1.0):
    """nx*ny pixel centres spread over the central `frac` of the full-size image."""
    W, H = flat.params.image_width, flat.params.image_height
    xs = (W * (0.5 - frac / 2) + (np.arange(nx) + 0.5) * W * frac / nx).astype(int) + 0.5
    ys = (H * (0.5 - frac / 2) + (np.arange(ny) + 0.5) * H * frac / ny).astype(int) + 0.5
    gx, gy = np.meshgrid(xs, ys)
    return np.ascontiguousarray(np.stack([gx.ravel(), gy.ravel()], axis=1).astype(np.float64))


def rel_err(a, ref):
    return (np.abs(a - ref) / np.maximum(np.abs(ref), 1e-2)).max(axis=1)


def err_stats(rgb, ref):
    e = rel_err(rgb, ref)
    m_ref = ref.mean(0)
    return {
        "samples": int(len(e)),
        "median_rel_err": float(np.median(e)),
        "p99_rel_err": float(np.quantile(e, 0.99)),
        "max_rel_err": float(e.max()),
        "frac_beyond_1e-3": float((e > 1e-3).mean()),
        "frac_beyond_1e-2": float((e > 1e-2).mean()),
        "frac_beyond_1e-6": float((e > 1e-6).mean()),
        "mean_rgb": [float(v) for v in rgb.mean(0)],
        "mean_rgb_oracle": [float(v) for v in m_ref],
        "mean_rel_dev": [float(v) for v in np.abs(rgb.mean(0) - m_ref) / np.maximum(m_ref, 1e-6)],
    }


def pack8(c):
    """cps_from_cl (reference scene.c:76-82): floor(256 c), 255 at >= 1."""
    return np.where(c > 0, np.where(c < 1, np.floor(np.clip(c, 0, 1) * 256), 255), 0).astype(np.uint8)


def ref_image(name):
    """The reference's shipped render of `name` as uint8 [h,w,3] (tests/golden/ref_images.npz, made by make_ref_images.py)."""
    z = np.load(os.path.join(GOLDEN, "ref_images.npz"))
    return z[name]


def image_vs_ref(avg, ref8):
    """avg: float [h,w,3] per-pixel averages of a full render; ref8: the shipped image.  Both compared as 8-bit images,
    the way the reference writes them: channel means (relative deviation) and RMSE in units of the full range."""
    a8 = pack8(avg).astype(np.float64)
    r8 = ref8.astype(np.float64)
    m, mr = a8.reshape(-1, 3).mean(0), r8.reshape(-1, 3).mean(0)
    return {
        "mean8": [float(v) for v in m / 256.0], "mean8_ref": [float(v) for v in mr / 256.0],
        "mean_rel_dev": [float(v) for v in np.abs(m - mr) / mr],
        "rmse": float(np.sqrt(((a8 - r8) ** 2).mean()) / 256.0),
    }


def full_render(flat, render_fn, passes=None):
    """scene_s_create_image_file (reference scene.c:1103-1159) with `render_fn(xy, index_base) -> rgb float32 [n,3]` as
    lum_machine_s_run: all gradient_cycles + 1 passes through the host pass controller.  Returns (Image, n_samples, passes)."""
    import actinon_b200 as acn
    prm = flat.params
    img = acn.Image(prm.image_width, prm.image_height)
    n_samples = n_pass = 0
    while passes is None or n_pass < passes:
        xy = img.next_pass(prm)
        if xy.shape[0] == 0:
            break
        rgb = np.ascontiguousarray(render_fn(xy, n_samples), dtype=np.float32)
        img.push(xy, rgb)
        n_samples += xy.shape[0]
        n_pass += 1
    return img, n_samples, n_pass


def load_ref_stats():
    p = os.path.join(GOLDEN, "ref_image_stats.json")
    return json.load(open(p)) if os.path.exists(p) else {}
