import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import numpy as np, synth
from mimosa_b200 import Context, IncrementalVoxelMap, ICPFactor, hornbill_config, HORNBILL_MAP
rng = np.random.default_rng(1)
ctx = Context(0)
m = IncrementalVoxelMap(ctx, **HORNBILL_MAP)
for _ in range(2):
    m.insert(synth.sample_world(100000, 40.0, rng))
R, t = synth.rot_from_rpy(0, 0.01, 0.2), np.array([1.0, 0.5, 0.1])
scan = synth.make_scan(R, t, int(os.environ.get("N_SCAN", "131072")), rng, max_range=36.0)
f = ICPFactor(ctx, m, scan, hornbill_config())
f.linearize(R, t)
fn = ctx.lib.mb_debug_time_finalize
fn.argtypes = [C.c_void_p, C.c_uint, C.c_int, C.POINTER(C.c_float)]
for mask, name in [(0, "none (launch only)"), (1, "role0 eig Hrr"), (2, "role1 eig Htt"), (4, "role2 Schur rr"), (8, "role3 Schur tt"), (16, "role4 pack"), (144, "role4 pack+solve+retract"), (31, "all"), (159, "all + step")]:
    us = C.c_float()
    assert fn(f.h, mask, 200, C.byref(us)) == 0
    print(f"mask {mask:2d} {name:22s} {us.value:7.2f} us/launch", flush=True)
