#!/usr/bin/env python
"""C1 (primitives 320x240, ds 10, ps 0, index-keyed): which samples of the FP32 product path lie beyond 1e-3 of the FP64 oracle, for
several shell widths (ACN_EPS_ULPS), generic and specialised kernels.  python tools/c1_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import actinon_b200 as acn
from tests.oracle_lib import Oracle
W, H = 320, 240
flat = acn.scenes.primitives(W, H, direct_samples=10, path_samples=0).flatten()
ys, xs = np.mgrid[0:H, 0:W]
xy = np.stack([(xs + 0.5).ravel(), (ys + 0.5).ravel()], axis=1).astype(np.float64)
ref, _ = Oracle().render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED)
def report(tag, rgb):
    e = (np.abs(rgb - ref) / np.maximum(np.abs(ref), 1e-2)).max(axis=1)
    bad = np.nonzero(e > 1e-3)[0]
    print(f"{tag:28s} beyond 1e-3: {len(bad):3d} = {len(bad)/len(e):.4%}  beyond 1e-2: {(e>1e-2).sum():3d}  median {np.median(e):.2e}",
          [(int(i % W), int(i // W), round(float(e[i]), 4)) for i in bad[:40]], flush=True)
for ulps in (16, 8, 4, 2, 1):
    os.environ["ACN_EPS_ULPS"] = str(ulps)
    for spec, nm in ((acn.SPECIALIZE_OFF, "generic"), (acn.SPECIALIZE_ON, "spec")):
        t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED, specialize=spec))
        rgb = t.render_samples(xy); t.close()
        report(f"f32 {ulps:2d} ulps {nm}", rgb)
del os.environ["ACN_EPS_ULPS"]
t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_INDEX_KEYED, precision=acn.PRECISION_F64))
report("f64", t.render_samples(xy)); t.close()
for e in (5e-6, 1e-5, 2e-5):
    w, _ = Oracle().render(flat, xy, seed_mode=acn.SEED_INDEX_KEYED, eps=e)
    report(f"oracle eps {e:g}", w)
