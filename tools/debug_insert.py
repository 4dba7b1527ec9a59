import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")]
import numpy as np, synth, oracle_py as o
from mimosa_b200 import Context, IncrementalVoxelMap
rng = np.random.default_rng(101)
pts = synth.sample_ground(5000, 10.0, rng)
ctx = Context(0)
mg = IncrementalVoxelMap(ctx, 1.0, 0.2, 20, 19, 1000)
mo = o.IVoxRef(1.0, 0.2, 20, 19, 1000)
mg.insert(pts); mo.insert(pts)
cg, ng, lg, pg, _ = mg.download(); co, no, lo, po, _ = mo.download()
print("sizes", mg.size(), mo.size(), "coords equal", np.array_equal(cg, co))
bad = np.flatnonzero(ng != no)
print("voxels with different counts:", bad.size, bad[:10])
vox_of = {tuple(c): i for i, c in enumerate(co.tolist())}
keys = np.floor(pts.astype(np.float64)).astype(int)
for v in bad[:3]:
    members = np.flatnonzero((keys == co[v]).all(1))
    print("voxel", v, co[v], "oracle count", no[v], "gpu count", ng[v], "input members", members.tolist())
    print(" oracle pts:\n", po[v, :no[v]])
    print(" gpu pts:\n", pg[v, :ng[v]])
    print(" input pts:\n", pts[members])
    # which input index each stored point is
    for name, arr, n in (("oracle", po, no[v]), ("gpu", pg, ng[v])):
        ids = [int(members[np.flatnonzero((pts[members] == arr[v, j]).all(1))[0]]) for j in range(n)]
        print(" ", name, "kept input ids", ids)
