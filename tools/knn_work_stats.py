"""Work distribution of the restricted k-NN on the bench's 'spread' query set (CPU, numpy + oracle map), at the
bench's point density but a 25x smaller area.  Simulates the scan order of the GPU kernel (own voxel, then
faces / edges by class with radius re-checks) and prints the per-query histogram of 4-point chunks, so kernel
restructurings can be judged before spending GPU time.  Development tool; not part of the product or tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import oracle_py as orc  # noqa: E402
import synth  # noqa: E402
from mimosa_b200.host import HORNBILL_MAP  # noqa: E402

HALF, M, NQ = 100.0, 400_000, 8192
rng = synth.rng_for(4)
mo = orc.IVoxRef(**HORNBILL_MAP)
synth.build_map(mo.insert, M, HALF, rng, size_fn=lambda: mo.size()[1])
coords, counts, _, pts, _ = mo.download()
print("map", mo.size(), "mean fill", counts.mean())
cloud = pts[np.arange(pts.shape[1])[None, :] < counts[:, None]]
q = synth.spread_queries(cloud, NQ, synth.rng_for(40))
vox = {tuple(c): i for i, c in enumerate(coords)}
offs = [(i, j, k) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1) if not (i and j and k)]
order = sorted(range(19), key=lambda o: (sum(abs(x) for x in offs[o]), o))
chunks, surv, own_cnt, occ = [], [], [], []
for qq in q:
    c = np.floor(qq).astype(int)
    f = qq - c
    best = []
    n_chunks = 0
    n_surv = 0
    n_occ = 0
    for o in order:
        off = offs[o]
        v = vox.get((c[0] + off[0], c[1] + off[1], c[2] + off[2]))
        if v is None:
            continue
        n_occ += 1
        worst = best[4] if len(best) >= 5 else np.inf
        g = [(f[a] if off[a] < 0 else (1 - f[a]) if off[a] > 0 else 0.0) for a in range(3)]
        if off != (0, 0, 0) and g[0] ** 2 + g[1] ** 2 + g[2] ** 2 > worst:
            continue
        if off == (0, 0, 0):
            own_cnt.append(counts[v])
        else:
            n_surv += 1
        d = ((pts[v, :counts[v], :3].astype(np.float64) - qq) ** 2).sum(1)
        best = sorted(best + list(d))[:5]
        n_chunks += (counts[v] + 3) // 4
    chunks.append(n_chunks)
    surv.append(n_surv)
    occ.append(n_occ)
chunks, surv, occ = np.array(chunks), np.array(surv), np.array(occ)
print("occupied nbrs mean", occ.mean(), "survivors mean", surv.mean(), "max", surv.max())
print("chunks/query mean", chunks.mean(), "p50", np.percentile(chunks, 50), "p90", np.percentile(chunks, 90),
      "p99", np.percentile(chunks, 99), "max", chunks.max())
w = chunks[: NQ // 32 * 32].reshape(-1, 32)
print("per-warp max chunks: mean", w.max(1).mean(), "max", w.max(1).max(), " per-warp mean", w.mean(1).mean())
print("hist survivors", np.bincount(surv))
print("own count hist", np.bincount(np.array(own_cnt)))
