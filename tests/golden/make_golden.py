"""Generates the golden fixtures of the oracle (run from the repo root: python -m tests.golden.make_golden).

The reference itself cannot be executed here (its dependency beth is absent), so these vectors pin the
ORACLE against regressions; the oracle in turn is pinned against the reference's own leaf math
(tests/test_oracle_leaf.py) and against closed forms (tests/test_oracle_render.py)."""
import os

import numpy as np

import actinon_b200 as acn

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "primitives_direct": lambda: acn.scenes.primitives(320, 240, 10, 0),
    "primitives_path": lambda: acn.scenes.primitives(320, 240, 10, 3),
    "glass_ball": lambda: acn.scenes.glass_ball(160, 120, 8, 2),
    "csg_zoo": lambda: acn.scenes.csg_zoo(160, 120, 6, 2),
}


def sample_positions(w, h, n, seed):
    rng = np.random.default_rng(seed)
    xy = np.stack([rng.uniform(0, w, n), rng.uniform(0, h, n)], axis=1)
    xy[: n // 2] = np.floor(xy[: n // 2]) + 0.5          # half pixel centres, half jittered
    return xy


def main():
    from tests.oracle_lib import Oracle
    orc = Oracle()
    for i, (name, mk) in enumerate(CASES.items()):
        sc = mk()
        flat = sc.flatten()
        xy = sample_positions(flat.params.image_width, flat.params.image_height, 400, 100 + i)
        seed_mode = 1
        index_base = 1000 * i
        rgb, info = orc.render(flat, xy, index_base=index_base, seed_mode=seed_mode)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), xy=xy, rgb=rgb, seed_mode=seed_mode, index_base=index_base)
        print(name, rgb.mean(0), info["rays"])


if __name__ == "__main__":
    main()
