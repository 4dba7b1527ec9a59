import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, bench, synth
rng, scan, R0, t0, Rt, tt = bench.make_inputs()
q = scan[:, :3].astype(np.float64) @ R0.T + t0
c = np.floor(q).astype(np.int64)
key = ((c[:,0]+(1<<20))<<42)|((c[:,1]+(1<<20))<<21)|(c[:,2]+(1<<20))
u, cnt = np.unique(key, return_counts=True)
print("points", len(q), "distinct voxels", len(u))
for t in (1,2,4,8,16,32,64,128,256,1024):
    print(f"voxels with >= {t} pts: {np.sum(cnt>=t)}  points in them: {cnt[cnt>=t].sum()} ({cnt[cnt>=t].sum()/len(q):.3f})")
print("max per voxel", cnt.max())
rng_ = np.linalg.norm(scan[:,:3],axis=1)
print("range pctiles", np.percentile(rng_,[1,10,25,50,75,90,99]))
# warps of 32 in sorted order: distinct voxels per warp
order = np.argsort(key, kind="stable")
ks = key[order]
nw = len(q)//32
d = np.array([len(np.unique(ks[w*32:(w+1)*32])) for w in range(nw)])
print("distinct voxels per sorted warp: mean", d.mean(), "pctiles", np.percentile(d,[10,25,50,75,90,99]))
# beam-order
d2 = np.array([len(np.unique(key[w*32:(w+1)*32])) for w in range(nw)])
print("distinct voxels per beam-order warp: mean", d2.mean(), "pctiles", np.percentile(d2,[10,25,50,75,90,99]))
# 2x2x2 supervoxel / 4x4x4 block grouping
b = c>>2
bkey = ((b[:,0]+(1<<20))<<42)|((b[:,1]+(1<<20))<<21)|(b[:,2]+(1<<20))
ub, bc = np.unique(bkey, return_counts=True)
print("distinct 4^3 blocks", len(ub), "mean pts/block", bc.mean())
