import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'oracle')); sys.path.insert(0,os.path.join(ROOT,'tests'))
import numpy as np, orc_py
from prt_b200 import scenes
from prt_b200.engine import DeviceScene
from test_gpu_parity import _random_rays
for name,order,seed in [('Sphere_Box','mitsuba',7),('Cone_Box','mitsuba',7),('Cone_Box','intended',7),('Plate_Box','mitsuba',7),('ring','',11)]:
    if name=='ring':
        desc=scenes.test_ring_scene(); o,d=_random_rays(50000,11,lo=(-0.07,-0.03,-0.02),hi=(0.07,0.03,0.15)); d[::2,2]*=-1
    else:
        desc=scenes.ultrasound_scene(name,order); o,d=_random_rays(20000,seed)
    g=DeviceScene(desc).trace_closest(o,d); oc=orc_py.OracleScene(desc)
    c=oc.trace_closest(o,d,prec=32); c64=oc.trace_closest(o,d,prec=64)
    ok=(g['prim']==c['prim'])&(g['prim']>=0)&(c64['prim']==c['prim'])
    err=np.abs(g['t'].astype(np.float64)-c64['t']); cos=np.abs((d*c64['ng']).sum(1))
    bad=ok&(err>1e-5*np.abs(c64['t'])+3e-8)&(cos>=5e-3)
    print(name,order,'bad',bad.sum())
    for j in np.nonzero(bad)[0][:8]:
        print('  prim %d cos %.4f  t_gpu %.9e t_f32 %.9e t_f64 %.9e  o %s d %s'%(g['prim'][j],cos[j],g['t'][j],c['t'][j],c64['t'][j],o[j],d[j]))
