#!/usr/bin/env python
"""Audit of the CSG event sweep on every shipped scene: f64 sweep (csg_mode INTERVALS) and f32 product mode against the
CPU oracle's reference march, on a grid of samples with reduced sample counts.  usage: sweep_audit.py [scene ...]"""
import sys, os, glob, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import actinon_b200 as acn
from tests.oracle_lib import Oracle
from tests.test_gpu_configs import grid_samples, rel_err
orc = Oracle()
names = sys.argv[1:] or sorted(os.path.basename(p)[:-4] for p in glob.glob("scenes/*.npz") if "video_0000" not in p or p.endswith("000049.npz"))
for name in names:
    big = name in ("hanging_lamps_in_row",)
    flat = acn.scenes.load(name, direct_samples=3 if big else 6, path_samples=2 if big else 4)
    xy = grid_samples(flat, 24 if big else 48, 24 if big else 48, 0.9)
    t0 = time.time(); ref, info = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED); tc = time.time() - t0
    out = [f"{name:28s} oracle {tc:5.1f}s rays {info['rays']:8d}"]
    for prec, tag in ((acn.PRECISION_F64, "f64-sweep"), (acn.PRECISION_F32, "f32")):
        t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED, precision=prec, csg_mode=acn.CSG_INTERVALS, wave_budget=1 << 18))
        rgb = t.render_samples(xy); st = t.last_stats; t.close()
        e = rel_err(rgb, ref)
        out.append(f"{tag}: >1e-5 {(e > 1e-5).mean():.4f} >1e-3 {(e > 1e-3).mean():.4f} >1e-2 {(e > 1e-2).mean():.4f} rays {st.rays - info['rays']:+d}")
    wide, _ = orc.render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED, eps=2e-5)      # the f32 path's shell thickness, in the FP64 oracle
    ew = rel_err(wide, ref)
    out.append(f"oracle eps 2e-5: >1e-3 {(ew > 1e-3).mean():.4f} >1e-2 {(ew > 1e-2).mean():.4f}")
    print(" | ".join(out), flush=True)
