"""Development microbenchmark: the ceiling of the k-NN bucket gather on this GPU.  A 10 M-point map's search mirror
is 1.27 M buckets of 320 B (406 MB); one spread launch touches ~807 k distinct buckets and ~104 MB of them.  Here the
same volume is fetched by a kernel that does nothing else (tools/probe/gather_probe.cu), for several access patterns,
timed with CUDA events after an L2 flush.  Output: gpurun_out/gather_probe.json.
Build (CPU box): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC tools/probe/gather_probe.cu -o tools/probe/libgather_probe.so"""
import ctypes as C
import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(HERE, "libgather_probe.so"))
lib.gather_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int]
dev = torch.device("cuda:0")
N_VOX, CAP = 1_268_577, 20
N_ITEMS = 807_154
pts = torch.rand(N_VOX * CAP * 4, device=dev)
out = torch.zeros(16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rng = np.random.default_rng(1)
rand = rng.choice(N_VOX, N_ITEMS, replace=False).astype(np.uint32)
# "clustered": the distinct buckets of Morton-sorted queries are neighbours in the Morton-ordered mirror: runs of
# consecutive slots (here runs of 8) at random places, visited in ascending order
runs = np.sort(rng.choice(N_VOX // 8, N_ITEMS // 8, replace=False)).astype(np.uint32)
clustered = (runs[:, None] * 8 + np.arange(8, dtype=np.uint32)[None, :]).reshape(-1)
patterns = {
    "random order, random buckets": rand,
    "ascending order, random buckets": np.sort(rand),
    "ascending runs of 8 buckets": clustered,
    "dense prefix (streaming)": np.arange(N_ITEMS, dtype=np.uint32),
}
res = {"n_items": N_ITEMS, "results": []}
for name, sl in patterns.items():
    d_sl = torch.from_numpy(sl.astype(np.int64)).to(dev).to(torch.int32)  # uint32 values fit
    for ppi, shape in ((8, 0), (4, 0), (20, 0), (8, 1)):
        ms = []
        for it in range(8):
            flush.fill_(it)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.gather_probe(pts.data_ptr(), d_sl.data_ptr(), sl.shape[0], ppi, CAP, out.data_ptr(), shape)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0, rc
            ms.append(e0.elapsed_time(e1))
        us = float(np.mean(ms[3:])) * 1e3
        nbytes = sl.shape[0] * ppi * 16
        rec = {"pattern": name, "points_per_item": ppi, "shape": "thread per float4" if shape == 0 else "thread per bucket (4 x LDG.128 per step)",
               "bytes": nbytes, "us": us, "gbs": nbytes / us / 1e3}
        res["results"].append(rec)
        print(json.dumps(rec), flush=True)
os.makedirs(os.path.join(HERE, "..", "..", "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(HERE, "..", "..", "gpurun_out", "gather_probe.json"), "w"), indent=1)
