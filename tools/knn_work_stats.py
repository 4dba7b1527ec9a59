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

HALF, M, NQ = 100.0, 400_000, 65536
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
per_query = []  # (own chunks, [neighbour bucket chunks in scan order])
for qq in q:
    c = np.floor(qq).astype(int)
    f = qq - c
    best = []
    n_chunks = 0
    n_surv = 0
    own_c, nb_c = 0, []
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
        if off == (0, 0, 0):
            own_c = (counts[v] + 3) // 4
        else:
            nb_c.append((counts[v] + 3) // 4)
    chunks.append(n_chunks)
    per_query.append((own_c, nb_c))
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
wm = w.max(1)
print("per-warp trip count percentiles (p50, p90, p99, max):", np.percentile(wm, [50, 90, 99, 100]))
print("per-warp mean-lane chunks percentiles:", np.percentile(w.mean(1), [50, 90, 99, 100]))
print("lane efficiency inside warps (mean/max):", (w.mean(1) / wm).mean())


def simulate(warp, handoff):
    """iterations of the neighbour loop for one warp (list of per-lane bucket chunk lists)"""
    work = [list(nb) for _, nb in warp]   # remaining buckets (first = in progress)
    cur = [0] * len(work)                 # remaining chunks of the bucket in progress
    done_handoff = 0
    it = 0
    while True:
        busy = [cur[l] > 0 or len(work[l]) > 0 for l in range(len(work))]
        if not any(busy):
            return it
        if done_handoff < handoff:
            donors = [l for l in range(len(work)) if len(work[l]) >= 2]
            idle = [l for l in range(len(work)) if not busy[l]]
            if donors and len(idle) >= len(donors):
                for d, h in zip(donors, idle):
                    give = work[d][1::2]
                    work[d] = work[d][0::2]
                    work[h] = give
                done_handoff += 1
        for l in range(len(work)):
            if cur[l] == 0 and work[l]:
                cur[l] = work[l].pop(0)
            if cur[l] > 0:
                cur[l] -= 1
        it += 1

W = [per_query[i:i + 32] for i in range(0, len(per_query) // 32 * 32, 32)]
for h in (0, 1, 2, 3):
    its = np.array([simulate(w_, h) for w_ in W])
    print("handoffs", h, "neighbour-loop iterations per warp: mean %.2f p50 %d p90 %d p99 %d max %d" % (
        its.mean(), *np.percentile(its, [50, 90, 99, 100])))
