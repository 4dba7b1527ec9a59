"""Development: the restricted k-NN launch under every MB_KNN_VARIANT (thread = the default one-query-per-thread
search, coop4 / coop8 = mb_search_coop.cuh) on the bench's spread query set: results compared bit for bit with the
default variant's, then CUDA-event times with the L2 flushed before every launch.  One process, one map build.
usage: python tools/knn_variants.py [n_timed]   (MB_BENCH_SMALL=1 for a dry run on a small map)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np
import torch  # noqa: F401  (maps torch's bundled NCCL before libmimosa_b200.so asks for it)

import bench
import synth
from mimosa_b200 import HORNBILL_MAP, Context, IncrementalVoxelMap

n_timed = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = Context(0)
rng, scan, R0, t0, _, _ = bench.make_inputs()
mg = IncrementalVoxelMap(ctx, **HORNBILL_MAP)
synth.build_map(mg.insert, bench.MAP_POINTS, bench.MAP_HALF_EXTENT, rng, size_fn=lambda: mg.size()[1])
coords, counts, _, pts, _ = mg.download()
cloud = pts[np.arange(pts.shape[1])[None, :] < counts[:, None]]
q_spread = synth.spread_queries(cloud, bench.N_SCAN, synth.rng_for(40))
q_local = scan[:, :3].astype(np.float64) @ R0.T + t0
b_alg, _ = bench.knn_algorithmic_bytes(q_spread, coords, counts, bench.K_NN)
peak, _ = bench.peaks()
out = {"algorithmic_bytes": int(b_alg), "peak_gbs": peak, "variants": {}}
ref = {}
for variant in ("thread", "coop4", "coop8"):
    os.environ["MB_KNN_VARIANT"] = variant
    rec = {}
    for name, q in (("spread", q_spread), ("local", q_local)):
        res = mg.knn_search(q, bench.K_NN)
        if variant == "thread":
            ref[name] = res
        else:
            ok = np.array_equal(res[2], ref[name][2])
            sel = ref[name][2]
            rec[name + "_bit_exact"] = bool(ok and np.array_equal(res[0][sel], ref[name][0][sel]) and np.array_equal(res[1][sel], ref[name][1][sel]))
        mg.knn_stage(q, bench.K_NN)
        ms = []
        for it in range(3 + n_timed):
            ctx.flush_l2()
            ctx.sync()
            ctx.timer_begin()
            mg.knn_staged_run()
            ms.append(ctx.timer_end())
        rec[name + "_us"] = float(np.mean(ms[3:])) * 1e3
        rec[name + "_us_min"] = float(np.min(ms[3:])) * 1e3
    rec["spread_frac_of_hbm_peak"] = b_alg / (rec["spread_us"] * 1e-6) / 1e9 / peak
    out["variants"][variant] = rec
    print(variant, json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "knn_variants.json"), "w"), indent=1)
