import numpy as np, time, sys
import actinon_b200 as acn
from tests.oracle_lib import Oracle
o = Oracle()
print("devices", acn.device_count())
cases = [
    ("primitives", acn.scenes.primitives(320, 240, 10, 0)),
    ("primitives_path", acn.scenes.primitives(160, 120, 10, 4)),
    ("glass_ball", acn.scenes.glass_ball(160, 120, 8, 0)),
    ("glass_ball_path", acn.scenes.glass_ball(80, 60, 6, 3)),
    ("csg_zoo", acn.scenes.csg_zoo(160, 120, 6, 0)),
    ("csg_zoo_path", acn.scenes.csg_zoo(80, 60, 4, 3)),
]
for name, sc in cases:
    flat = sc.flatten()
    W, H = flat.params.image_width, flat.params.image_height
    img = acn.Image(W, H)
    xy = img.next_pass(flat.params)
    t0 = time.time()
    ref, info = o.render(flat, xy, seed_mode=1)
    print(name, "oracle s", time.time() - t0, "rays", info["rays"], "flops/sample", info["flops"] / len(xy))
    for prec in (acn.PRECISION_F64, acn.PRECISION_F32):
        t = acn.Tracer(flat, acn.Options(seed_mode=1, precision=prec))
        rgb = t.render_samples(xy)
        rgb = t.render_samples(xy)
        st = t.last_stats
        err = np.abs(rgb - ref) / np.maximum(np.abs(ref), 1e-2)
        print("  prec", prec, "ms %.3f" % st.device_ms, "rays", st.rays, "launches", st.kernel_launches, "waves", st.waves,
              "| max err %.3g median %.3g frac>1e-3 %.5f" % (err.max(), np.median(err), (err.max(1) > 1e-3).mean()),
              "mean gpu", rgb.mean(0), "mean ref", ref.mean(0))
        np.save(f"gpurun_out/{name}_{prec}.npy", rgb)
        t.close()
    np.save(f"gpurun_out/{name}_ref.npy", ref)
print("fp32 peak TF", acn.measure_fp32_peak_tflops())
