import sys, time, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from distance3d_b200 import gjk, random as R
names=sys.argv[1].split(",") if len(sys.argv)>1 else ["sphere","ellipsoid","capsule","cylinder","box"]
n=int(os.environ.get("D3D_N", 1<<20))
DTYPE=os.environ.get("D3D_DTYPE","f64")
def run(cs,pairs,label):
    dc=cs.device(); pd=torch.from_numpy(pairs).cuda()
    out=gjk.gjk_distance_batch(dc,pd,dtype=DTYPE)
    for _ in range(2): gjk.gjk_distance_batch(dc,pd,out=out,dtype=DTYPE)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): gjk.gjk_distance_batch(dc,pd,out=out,dtype=DTYPE)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    it=out.iters.double().mean().item()
    print(f"{label:22s} {n/ms/1e3:8.1f} Mpairs/s  iters {it:5.1f}  -> {n*it/ms/1e6:7.2f} G pair-iters/s  hit {float((out.dist==0).double().mean()):.2f}")
for a in names:
    for b in names:
        if names.index(b)<names.index(a): continue
        rs=np.random.RandomState(1)
        csa=R.random_collider_set(rs,n,names=(a,)); csb=R.random_collider_set(rs,n,names=(b,))
        # merge: interleave
        from distance3d_b200.pack import ColliderSet
        import numpy as np
        type_=np.concatenate([csa.type,csb.type]); pose=np.concatenate([csa.pose,csb.pose]); param=np.concatenate([csa.param,csb.param])
        vl=np.concatenate([csa.vert_len,csb.vert_len]); vo=np.zeros(2*n,dtype=np.int64); vo[1:]=np.cumsum(vl[:-1])
        cs=ColliderSet(type_,pose,param,vo,vl,np.concatenate([csa.verts[:int(csa.vert_len.sum())] if csa.vert_len.sum() else np.zeros((0,3)), csb.verts[:int(csb.vert_len.sum())] if csb.vert_len.sum() else np.zeros((0,3))]))
        pairs=np.stack([np.arange(n),np.arange(n)+n],axis=1).astype(np.int32)
        run(cs,pairs,a+"-"+b)
if os.environ.get("D3D_SKIP_MIX"): sys.exit(0)
rs=np.random.RandomState(1)
cs=R.random_collider_set(rs,2*n,names=R.PRIMITIVES); pairs=np.arange(2*n,dtype=np.int32).reshape(n,2)
run(cs,pairs,"mix")
cs=R.random_collider_set(rs,2*n,names=R.PRIMITIVES+("mesh",)); run(cs,pairs,"mix6 (10-vertex hulls)")
