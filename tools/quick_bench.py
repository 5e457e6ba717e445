#!/usr/bin/env python
"""Device-resident timing of one scene: python tools/quick_bench.py [scene] [steps] [budget] [width height]  (env ACN_B200_LIBRARY selects a variant)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import actinon_b200 as acn
scene = sys.argv[1] if len(sys.argv) > 1 else "wine_glass"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
budget = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ov = {}
if len(sys.argv) > 5: ov = dict(image_width=int(sys.argv[4]), image_height=int(sys.argv[5]))
flat = acn.scenes.load(scene, **ov)
p = flat.params
W, H = p.image_width, p.image_height
ys, xs = np.mgrid[0:H, 0:W]
xy = np.stack([(xs + 0.5).ravel(), (ys + 0.5).ravel()], axis=1).astype(np.float64)
t = acn.Tracer(flat, acn.Options(seed_mode=acn.SEED_POSITION_HASH, wave_budget=budget))
d_xy = torch.from_numpy(xy).cuda(); d_rgb = torch.empty((len(xy), 3), dtype=torch.float32, device="cuda")
for _ in range(1): t.render_samples_device(d_xy, d_rgb)
torch.cuda.synchronize()
ms = []
for _ in range(steps):
    t.render_samples_device(d_xy, d_rgb); ms.append(t.last_stats.device_ms)
st = t.last_stats
os.environ["ACN_PROFILE_KERNELS"] = "1"
t.render_samples_device(d_xy, d_rgb); sp = t.last_stats
print(f"{os.environ.get('ACN_B200_LIBRARY','default'):50s} {scene} ms/step {np.mean(ms):8.2f} (min {min(ms):.2f}) rays {st.rays/1e6:.1f}M waves {st.waves} "
      f"kernel_ms prim {sp.kernel_ms[0]:.2f} rays {sp.kernel_ms[1]:.2f} path {sp.kernel_ms[2]:.2f} direct {sp.kernel_ms[3]:.2f} shade {sp.kernel_ms[4]:.2f} index {sp.kernel_ms[5]:.2f} sched+pop {sp.kernel_ms[6]:.2f} mean_rgb {d_rgb.mean(0).cpu().numpy()}")
