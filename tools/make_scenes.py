"""Flattens the reference's scene scripts with OUR .acn front-end and stores the result under scenes/.

/root/reference does not exist on the GPU box, and its scripts are not copied into this repository;
what is committed is the flattened node/material tables our interpreter produced from them (own
format, see actinon_b200.api.save_flat) plus this generating script.

    python tools/make_scenes.py            # needs /root/reference/src_acn
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import actinon_b200 as acn  # noqa: E402
from actinon_b200 import api  # noqa: E402

SRC = "/root/reference/src_acn"
SCENES = {
    "primitives": "primitives.acn",
    "wine_glass": "wine_glass.acn",
    "many_spheres": "many_spheres.acn",
    "diamond": "diamond.acn",
    "hanging_lamps_in_row": "hanging_lamps_in_row/hanging_lamps_in_row.acn",
    "hanging_lamp": "hanging_lamp/hanging_lamp.acn",
    "pyramid": "pyramid.acn",
    "ruby_heart": "ruby_heart.acn",
    "caustic_of_caustic": "caustic_of_caustic.acn",
    "paraffin_lamp": "paraffin_lamp/paraffin_lamp.acn",
    "paraffin_lamp_on_ledge": "paraffin_lamp_on_ledge/paraffin_lamp_on_ledge.acn",
}


def main():
    out = os.path.join(ROOT, "scenes")
    os.makedirs(out, exist_ok=True)
    for name, rel in SCENES.items():
        sc = acn.Scene()
        n = sc.load_acn(os.path.join(SRC, rel))
        assert n == 1, (name, n)
        sc.select_image(0)
        flat = sc.flatten()
        api.save_flat(flat, os.path.join(out, name + ".npz"), name)
        print(name, flat.struct.n_nodes, "nodes")
    # diamond_video: 90 frames (angle = 25 + index, diamond_video.acn:218); keep every 10th + frame 49
    sc = acn.Scene()
    n = sc.load_acn(os.path.join(SRC, "diamond_video.acn"))
    assert n == 90
    for i in list(range(0, 90, 10)) + [49, 89]:
        sc.select_image(i)
        api.save_flat(sc.flatten(), os.path.join(out, f"diamond_video_{i:06d}.npz"), f"diamond_video frame {i}")
    print("diamond_video frames written")


if __name__ == "__main__":
    main()
