"""CPU only: random parameterisations of whole scan sequences (voxel size, operation area, clustering tolerance, classification gates, ray length and
update rule, its_diff of both background threads, sensor mask + beam offsets, intensity gate, all three scenes) through the REFERENCE's own code
(oracle/_ref: the member functions of vofod_nodelet.cpp + voxel_map.cpp + voxel_grid_*.cpp compiled from /root/reference) and through the oracle,
compared bit for bit scan by scan (tests/nodelet_cases.py::compare_exact) — a wider net than the committed fixture.  The oracle runs with PCL's own
(unstable) cluster sort (set_std_sort_ties): the order of equal-size clusters is the one known difference.
    python tools/fuzz_oracle_vs_reference.py SEED N_CASES"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from nodelet_cases import run_case, compare_exact
from harness import params_for
from oracle import oracle, ref
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 1)
n_cases=int(sys.argv[2]) if len(sys.argv)>2 else 10
bad=0
for ci in range(n_cases):
    size=[(80.,80.,30.),(60.,60.,24.),(100.,70.,20.)][rng.integers(3)]
    p=params_for(size); p.background_sufficient_points_ratio=float(rng.choice([0.01,0.02,0.05]))
    p.ground_points_max_distance=float(rng.choice([1.0,1.5,2.0,1.2]))
    p.cls_max_size=float(rng.choice([1.5,3.0,5.0])); p.cls_max_explore_distance=float(rng.choice([1.0,3.0,4.5])); p.cls_max_distance=float(rng.choice([20.,50.]))
    p.cls_min_points=int(rng.choice([1,2,4]))
    p.raycast_max_distance=float(rng.choice([8.,20.,30.])); p.raycast_weight_coefficient=float(rng.choice([0.05,0.1,0.3]))
    p.raycast_new_update_rule=int(rng.integers(2))
    p.sep_max_bg_distance=float(rng.choice([1.0,1.5])); p.sep_min_sure_points=int(rng.choice([5,24,60]))
    vs=float(rng.choice([0.5,0.5,0.25,1.0]))
    if vs==0.25: size=(40.,40.,16.)
    scene=int(rng.integers(3)); W,H=[(512,32),(1024,64)][int(scene==2)]
    nsc=int(rng.integers(12,34))
    sched=dict(raycast_its_diff=int(rng.integers(1,4)), sep_its_diff=int(rng.integers(1,4)))
    if vs==0.25:
        for i,(o,s) in enumerate(zip((0.,0.,-1.25),size)): p.oparea_offset[i]=o; p.oparea_size[i]=s
    c=dict(W=W,H=H,p=p,vs=vs,scene=scene,scans=range(0,nsc),sched=sched,map_scale=(0.5 if vs==0.25 else 1.0))
    if rng.random()<0.3: c["mask_offsets"]=int(rng.integers(100)); 
    if rng.random()<0.3: c["dim_every"]=int(rng.integers(3,9)); p.raycast_min_intensity=50.0
    t0=time.time()
    try:
        r=ref.RefNodelet(); want=run_case(r,c); r.close(); want.pop("map_last",None)
        o=oracle.Oracle(track_counts=False, apply_from_fixed=False); o.set_std_sort_ties(True); got=run_case(o,c); o.close()
        compare_exact(got,want,"oracle",f"fuzz{ci}"); ok="EXACT"
    except AssertionError as e:
        ok="MISMATCH "+str(e)[:200]; bad+=1
    except Exception as e:
        ok="ERROR "+type(e).__name__+" "+str(e)[:200]; bad+=1
    print(ci, dict(vs=vs,scene=scene,W=W,n=nsc,new=p.raycast_new_update_rule,tol=p.ground_points_max_distance,ms=p.cls_max_size,ex=p.cls_max_explore_distance,mask=c.get("mask_offsets"),dim=c.get("dim_every"),**sched), "dets", int(want["n_det"].sum()) if 'want' in dir() else -1, ok, round(time.time()-t0,1), flush=True)
print("bad", bad)
