import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vofod_b200 import abi, capi, synth
import ctypes as C
W,H=2048,128
d=synth.sim_lut(W,H)
p=abi.default_params()
for i,(o,s) in enumerate(zip((0.,0.,-1.25),(200.,200.,80.))):
    p.oparea_offset[i]=o; p.oparea_size[i]=s
v=capi.Vofod(0); v.reset(p,0.5); v.set_sensor(W,H,d)
v.set_option(abi.OPT_GRAPH,0)
scans=[synth.generate(0,k,W,H,d) for k in (30,31,32,33)]
for mode, rb in ((0,64),(0,128),(0,256),(0,64),(0,128),(0,256),(1,128)):
    v.set_option(2,mode)
    v.set_option(5,rb)
    v.reset(p,0.5)
    ts=[]
    for (scan,pose,rp,_) in scans*3:
        res,_=v.process_scan(scan,pose,p,abi.schedule_s1(rp))
        ts.append(v.stage_times()['raycasting'])
    print("no_agg" if mode else "agg", rb, "raycasting ms:", np.round(ts[4:],4), "trav", res.n_traversals, flush=True)
