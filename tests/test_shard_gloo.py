"""CPU, world_size 2 over gloo: the multi-GPU decomposition (contiguous scan blocks, replicated map, one
all-reduce of the packed normal equations per iteration, identical solve on every rank) reproduces the
single-rank result.  The per-rank linearisation is played by the oracle; the sharding / packing logic is the
host code the GPU path uses (mimosa_b200.host.shard_range)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from mimosa_b200.host import shard_range

    for n in (0, 1, 7, 32, 131072, 65280):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert max(e - b for b, e in blocks) <= (n + world - 1) // world


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
        sys.path.insert(0, p)
    import oracle_py as orc
    import synth
    from mimosa_b200.host import HORNBILL_MAP, hornbill_config, shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    m = orc.IVoxRef(**HORNBILL_MAP)  # replicated map: every rank applies the same deterministic insert
    m.insert(synth.sample_world(150000, 40.0, rng))
    R_true, t_true = synth.rot_from_rpy(0, 0, 0.1), np.array([0.5, 0.2, 0.0])
    scan = synth.make_scan(R_true, t_true, 6000, rng, max_range=36.0)
    R, t = synth.perturbed_start(R_true, t_true)
    b, e = shard_range(scan.shape[0], rank, world)
    f = orc.IcpFactorRef(m, scan[b:e], hornbill_config())
    for _ in range(4):
        L = f.linearize(R, t)
        packed = torch.tensor(list(L.H) + list(L.g) + [L.f] + [float(c) for c in L.counts], dtype=torch.float64)
        dist.all_reduce(packed)  # the single exchange step of the path
        H, g = packed[:36].numpy().reshape(6, 6), packed[36:42].numpy()
        ok, d = orc.solve6(H, 0.0, g)
        assert ok
        R, t = orc.se3_retract(R, t, d)  # identical on every rank: no broadcast needed
    if rank == 0:
        np.savez(out, R=R, t=t, H=H, counts=packed[43:].numpy())
    gathered = [torch.zeros(12, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.tensor(np.concatenate([R.ravel(), t])))
    assert all(torch.equal(gathered[0], x) for x in gathered)
    dist.destroy_process_group()


def test_sharded_icp_equals_single_rank(tmp_path, oracle):
    import synth
    from mimosa_b200.host import HORNBILL_MAP, hornbill_config

    out = str(tmp_path / "r0.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    rng = np.random.default_rng(7)
    m = oracle.IVoxRef(**HORNBILL_MAP)
    m.insert(synth.sample_world(150000, 40.0, rng))
    R_true, t_true = synth.rot_from_rpy(0, 0, 0.1), np.array([0.5, 0.2, 0.0])
    scan = synth.make_scan(R_true, t_true, 6000, rng, max_range=36.0)
    R, t = synth.perturbed_start(R_true, t_true)
    f = oracle.IcpFactorRef(m, scan, hornbill_config())
    R1, t1, trace, _ = f.icp_run(R, t, 4, 0.0)
    assert np.abs(got["R"] - R1).max() < 1e-10 and np.abs(got["t"] - t1).max() < 1e-10
    assert got["counts"].tolist() == [float(c) for c in trace[-1].counts]
    H1 = np.array(trace[-1].H).reshape(6, 6)
    assert np.linalg.norm(got["H"] - H1) <= 1e-9 * np.linalg.norm(H1)
