import sys; sys.path.insert(0,'/root/repo')
import numpy as np, actinon_b200 as acn
from tests.oracle_lib import Oracle
from tests.test_gpu_configs import load_case, rel_err
o=Oracle()
for name in ['many_spheres','paraffin_lamp','diamond']:
    flat,xy=load_case(name)
    ref,info=o.render(flat,xy,seed_mode=1)
    for prec,csg in ((acn.PRECISION_F64,0),(acn.PRECISION_F32,0)):
        t=acn.Tracer(flat,acn.Options(seed_mode=1,precision=prec,csg_mode=csg,wave_budget=1<<18))
        rgb=t.render_samples(xy); st=t.last_stats; t.close()
        e=rel_err(rgb,ref); bad=np.where(e>1e-6 if prec else e>1e-3)[0]
        print(name,'prec',prec,'max',e.max(),'nbad',len(bad),'of',len(e),'rays',st.rays,info['rays'],'shadow',st.rays_shadow,info['counters']['rays_shadow'],'path',st.rays_path,info['counters']['rays_path'],'refl',st.rays_reflection,info['counters']['rays_reflect'],'refr',st.rays_refraction,info['counters']['rays_refract'])
        for b in bad[:5]: print('   ',b,xy[b],rgb[b],ref[b])
print('---- f32 stats')
for name in ['wine_glass','many_spheres','diamond','paraffin_lamp','hanging_lamp']:
    flat,xy=load_case(name)
    ref,info=o.render(flat,xy,seed_mode=1)
    t=acn.Tracer(flat,acn.Options(seed_mode=1,wave_budget=1<<18)); rgb=t.render_samples(xy); t.close()
    e=rel_err(rgb,ref)
    print(name,'median',np.median(e),'p90',np.percentile(e,90),'>1e-3',(e>1e-3).mean(),'>1e-2',(e>1e-2).mean(),'>1e-1',(e>1e-1).mean(),'mean dev',np.abs(rgb.mean(0)-ref.mean(0))/ref.mean(0))
    # oracle with the f32 path's shell thickness: how much of the difference is the eps policy?
    ref2,_=o.render(flat,xy,seed_mode=1,eps=2e-5)
    e2=rel_err(ref2,ref)
    print('    oracle(eps=2e-5) vs oracle(1e-6): >1e-3',(e2>1e-3).mean(),'>1e-2',(e2>1e-2).mean())
