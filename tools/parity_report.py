#!/usr/bin/env python
"""Parity report of the scripted configurations (run on the GPU box):

    python tools/parity_report.py --cpu-full primitives wine_glass diamond many_spheres     # CPU oracle full renders (minutes)
    python tools/parity_report.py --merge                                                     # after `pytest -m gpu tests/test_gpu_scripted.py`

--cpu-full renders each scene as scripted (all gradient passes, host pass controller) with the FP64 CPU oracle on all host
threads and compares the result with the reference's shipped image: gpurun_out/ref_image_stats.json (commit it as
tests/golden/ref_image_stats.json — the RMSE bar of test_full_scripted_render_vs_the_shipped_reference_image).
--merge collects gpurun_out/parity/*.json (written by the GPU tests) into gpurun_out/r02_parity.json (commit under profiles/).
"""
import argparse
import glob
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-full", nargs="*", default=None)
    ap.add_argument("--merge", action="store_true")
    args = ap.parse_args()
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    import actinon_b200 as acn
    from tests.parity_util import full_render, image_vs_ref, ref_image

    if args.cpu_full is not None:
        from tests.oracle_lib import Oracle
        orc = Oracle()
        path = os.path.join(out_dir, "ref_image_stats.json")
        stats = json.load(open(path)) if os.path.exists(path) else {}
        for name in args.cpu_full:
            flat = acn.scenes.load(name)
            t0 = time.perf_counter()
            img, n_samples, n_pass = full_render(flat, lambda xy, base: orc.render(flat, xy, index_base=base, seed_mode=0)[0])
            dt = time.perf_counter() - t0
            s = image_vs_ref(img.average(), ref_image(name))
            s.update(passes=n_pass, samples=n_samples, seconds=dt, threads=os.cpu_count(),
                     what="FP64 CPU oracle, position-hash seeding, full scripted render vs reference image/*.png (8-bit)")
            stats[name] = s
            print(name, json.dumps(s), flush=True)
            json.dump(stats, open(path, "w"), indent=1)

    if args.merge:
        rep = {"what": "parity of the CUDA path at the scripted configurations (SURVEY.md §8 C1-C5); written by tests/test_gpu_scripted.py "
                       "on a B200 and merged by tools/parity_report.py", "f32_vs_oracle": {}, "f64_vs_oracle": {}, "full_render_vs_shipped_image": {}}
        key = {"f32": "f32_vs_oracle", "f64": "f64_vs_oracle", "full": "full_render_vs_shipped_image"}
        for f in sorted(glob.glob(os.path.join(out_dir, "parity", "*.json"))):
            kind, name = os.path.basename(f)[:-5].split(".", 1)
            rep[key[kind]][name] = json.load(open(f))
        json.dump(rep, open(os.path.join(out_dir, "r02_parity.json"), "w"), indent=1)
        print("wrote gpurun_out/r02_parity.json")


if __name__ == "__main__":
    main()
