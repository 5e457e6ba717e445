#!/usr/bin/env python
"""Packs the reference's shipped full-quality renders (reference image/*.png) into tests/golden/ref_images.npz.

They are the only artefacts of the reference that pin the sample loop above leaf level (the program itself cannot be built:
its dependency `beth` is not in the tree), so the GPU tests compare full scripted renders with them.  The GPU box has no
/root/reference; run this here whenever the reference tree changes:

    python tests/golden/make_ref_images.py
"""
import os
import numpy as np
from PIL import Image

SRC = "/root/reference/image"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_images.npz")
FILES = {
    "primitives": "primitives.acn.png",
    "wine_glass": "wine_glass.acn.png",
    "many_spheres": "many_spheres.acn.png",
    "diamond": "diamond.acn.png",
    "diamond_video_000049": "diamond_video.acn.image_000049.png",
    "hanging_lamp02_640_360": "hanging_lamp02.acn.640_360.jpg",
}

if __name__ == "__main__":
    out = {}
    for k, f in FILES.items():
        a = np.asarray(Image.open(os.path.join(SRC, f)).convert("RGB"), dtype=np.uint8)
        out[k] = a
        print(k, a.shape, (a.reshape(-1, 3).mean(0) / 256.0 + 0.5 / 256.0).round(4))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
