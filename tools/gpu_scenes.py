import numpy as np, time, sys
import actinon_b200 as acn
from tests.oracle_lib import Oracle
o = Oracle()
names = sys.argv[1:] or ["wine_glass", "many_spheres", "diamond", "hanging_lamps_in_row"]
for name in names:
    over = dict(image_width=640, image_height=360) if name == "hanging_lamps_in_row" else {}
    flat = acn.scenes.load(name, **over)
    p = flat.params
    W, H = p.image_width, p.image_height
    rng = np.random.default_rng(1)
    n = 2048
    xy = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1)
    t0 = time.time()
    ref, info = o.render(flat, xy, seed_mode=1)
    dt = time.time() - t0
    print(f"{name}: oracle {n} samples {dt:.2f}s ({info['threads']} thr) rays/sample {info['rays']/n:.0f} flops/sample {info['flops']/n:.3g}", flush=True)
    for label, opt in (("f64 march", acn.Options(seed_mode=1, precision=1, csg_mode=2)), ("f64 intervals", acn.Options(seed_mode=1, precision=1, csg_mode=1)),
                       ("f32 march", acn.Options(seed_mode=1, csg_mode=2)), ("f32 intervals", acn.Options(seed_mode=1, csg_mode=1))):
        t = acn.Tracer(flat, opt)
        rgb = t.render_samples(xy)
        st = t.last_stats
        err = (np.abs(rgb - ref) / np.maximum(np.abs(ref), 1e-2)).max(1)
        print(f"   {label:14s}: rays {st.rays} vs {info['rays']}  frac>1e-6 {(err>1e-6).mean():.4f} frac>1e-3 {(err>1e-3).mean():.4f} frac>1e-2 {(err>1e-2).mean():.4f} ms {st.device_ms:.1f} waves {st.waves}", flush=True)
        t.close()
    t = acn.Tracer(flat, acn.Options(seed_mode=1))
    xy0 = acn.Image(W, H).next_pass(p)
    for rep in range(2):
        out = t.render_samples(xy0)
        st = t.last_stats
        print(f"   pass0 f32: {len(xy0)} samples device {st.device_ms:.1f} ms  rays {st.rays:.4g} ({st.rays/st.device_ms/1e6:.2f} Grays/s) samples/s {len(xy0)/st.device_ms*1e3:.4g} waves {st.waves} launches {st.kernel_launches}", flush=True)
    np.save(f"gpurun_out/{name}_pass0.npy", out.reshape(H, W, 3))
    t.close()
