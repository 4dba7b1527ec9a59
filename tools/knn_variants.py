"""Development: the restricted k-NN launch under every MB_KNN_VARIANT (thread = the default one-query-per-thread
search, threadq = the same with the warp-wide chunk queue, coop4 / coop8 / ... = mb_search_coop.cuh) on the bench's spread query set: results compared bit for bit with the
default variant's, then CUDA-event times with the L2 flushed before every launch.  One process, one map build.
Also timed: every 16th spread query alone (8192 queries: how long the launch lasts when throughput cannot matter).
usage: python tools/knn_variants.py [n_timed [variant ...]]   (MB_BENCH_SMALL=1 for a dry run on a small map)"""
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
variants = ["thread"] + [v for v in (sys.argv[2:] or ["threadq", "coop4b6", "coop8"]) if v != "thread"]
for variant in variants:
    os.environ["MB_KNN_VARIANT"] = variant
    rec = {}
    for name, q in (("spread", q_spread), ("local", q_local), ("spread_16th", q_spread[::16])):
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

# ---- the same launches with a CLEAN cold L2: the prescribed flush WRITES 256 MiB, which leaves the L2 full of dirty
# lines that the timed kernel's misses have to write back; here the L2 is filled by READING a 512 MiB tensor instead
big = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
big.fill_(1)
torch.cuda.synchronize()
out["clean_l2"] = {}
for variant in variants:
    os.environ["MB_KNN_VARIANT"] = variant
    mg.knn_stage(q_spread, bench.K_NN)
    ms = []
    for it in range(3 + n_timed):
        big.sum()
        torch.cuda.synchronize()
        ctx.timer_begin()
        mg.knn_staged_run()
        ms.append(ctx.timer_end())
    out["clean_l2"][variant] = {"spread_us": float(np.mean(ms[3:])) * 1e3, "spread_us_min": float(np.min(ms[3:])) * 1e3}
    print("clean-L2", variant, json.dumps(out["clean_l2"][variant]), flush=True)
# and the bare gather of the same byte volume (tools/probe/gather_probe.cu) under both protocols
try:
    import ctypes as C

    lib = C.CDLL(os.path.join(ROOT, "tools", "probe", "libgather_probe.so"))
    lib.gather_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int]
    n_vox, n_items = 1_268_577, 807_154
    pts_t = torch.rand(n_vox * 20 * 4, device="cuda:0")
    o_t = torch.zeros(16, device="cuda:0")
    prng = np.random.default_rng(1)
    runs = np.sort(prng.choice(n_vox // 8, n_items // 8, replace=False)).astype(np.int64)
    sl = torch.from_numpy((runs[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)).to("cuda:0").to(torch.int32)
    out["gather_probe"] = []
    for proto in ("write-flush", "read-flush"):
        for shape in (0, 1):
            ms = []
            for it in range(8):
                if proto == "write-flush":
                    big[: 256 << 20].fill_(it)
                else:
                    big.sum()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                lib.gather_probe(pts_t.data_ptr(), sl.data_ptr(), sl.shape[0], 8, 20, o_t.data_ptr(), shape)
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            rec = {"protocol": proto, "shape": shape, "bytes": int(sl.shape[0]) * 128, "us": float(np.mean(ms[3:])) * 1e3}
            out["gather_probe"].append(rec)
            print("gather", json.dumps(rec), flush=True)
except OSError as e:
    print("gather probe not built:", e)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "knn_variants.json"), "w"), indent=1)
