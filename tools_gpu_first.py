import numpy as np, time, sys
import actinon_b200 as acn
from tests.oracle_lib import Oracle
o = Oracle()
print("devices", acn.device_count())
for name, W, H, ds, ps in (("primitives", 320, 240, 10, 0), ("primitives_path", 160, 120, 10, 4)):
    sc = acn.scenes.primitives(W, H, ds, ps)
    flat = sc.flatten()
    img = acn.Image(W, H)
    xy = img.next_pass(flat.params)
    ref, info = o.render(flat, xy, seed_mode=1)
    for prec in (acn.PRECISION_F64, acn.PRECISION_F32):
        t = acn.Tracer(flat, acn.Options(seed_mode=1, precision=prec))
        rgb = t.render_samples(xy)
        rgb = t.render_samples(xy)
        st = t.last_stats
        err = np.abs(rgb - ref) / np.maximum(np.abs(ref), 1e-2)
        print(name, "prec", prec, "ms", st.device_ms, "rays", st.rays, "oracle rays", info["rays"], "launches", st.kernel_launches, "waves", st.waves)
        print("   max err", err.max(), "median", np.median(err), "frac>1e-3", (err.max(1) > 1e-3).mean(), "mean gpu", rgb.mean(0), "mean ref", ref.mean(0))
        np.save(f"gpurun_out/{name}_{prec}.npy", rgb)
        t.close()
    np.save(f"gpurun_out/{name}_ref.npy", ref)
print("fp32 peak TF", acn.measure_fp32_peak_tflops())
