import numpy as np, time, sys
import actinon_b200 as acn
from tests.oracle_lib import Oracle
o = Oracle()
cases = [
    ("primitives", acn.scenes.primitives(320, 240, 10, 0)),
    ("glass_ball", acn.scenes.glass_ball(160, 120, 8, 0)),
    ("csg_zoo", acn.scenes.csg_zoo(160, 120, 6, 0)),
]
for name, sc in cases:
    flat = sc.flatten()
    W, H = flat.params.image_width, flat.params.image_height
    xy = acn.Image(W, H).next_pass(flat.params)
    ref, info = o.render(flat, xy, seed_mode=1)
    for eps in (0.0, 4e-5, 2e-5, 1e-5, 5e-6):
        t = acn.Tracer(flat, acn.Options(seed_mode=1, precision=0, eps=eps))
        rgb = t.render_samples(xy)
        st = t.last_stats
        err = np.abs(rgb - ref) / np.maximum(np.abs(ref), 1e-2)
        msg = "  eps %g ms %.3f waves %d | vs oracle(1e-6): frac>1e-3 %.5f frac>1e-2 %.5f max %.3g" % (eps, st.device_ms, st.waves, (err.max(1) > 1e-3).mean(), (err.max(1) > 1e-2).mean(), err.max())
        if eps > 0:
            ref2, _ = o.render(flat, xy, seed_mode=1, eps=eps)
            err2 = np.abs(rgb - ref2) / np.maximum(np.abs(ref2), 1e-2)
            msg += " | vs oracle(same eps): frac>1e-3 %.5f frac>1e-2 %.5f" % ((err2.max(1) > 1e-3).mean(), (err2.max(1) > 1e-2).mean())
        print(name, msg)
        t.close()
    ref3, _ = o.render(flat, xy, seed_mode=1, eps=7.6e-5)
    err3 = np.abs(ref3 - ref) / np.maximum(np.abs(ref), 1e-2)
    print(name, "  oracle(7.6e-5) vs oracle(1e-6): frac>1e-3 %.5f frac>1e-2 %.5f" % ((err3.max(1) > 1e-3).mean(), (err3.max(1) > 1e-2).mean()))
