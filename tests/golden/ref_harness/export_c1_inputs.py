"""Writes the C1 inputs (tests/golden/make_golden.py: the point chunks fed to the map in order, the scan, the start pose)
as raw little-endian arrays for make_golden_ref.cpp.  usage: python tests/golden/ref_harness/export_c1_inputs.py <dir>"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import oracle_py as orc  # noqa: E402
import synth  # noqa: E402
from mimosa_b200.host import HORNBILL_MAP  # noqa: E402


def main(out):
    os.makedirs(out, exist_ok=True)
    rng = synth.rng_for(1)  # the same stream as make_golden.c1_inputs()
    m = orc.IVoxRef(**HORNBILL_MAP)
    chunks = []
    while m.size()[1] < 100_000:
        c = synth.sample_ground(20_000, 50.0, rng)
        m.insert(c)
        chunks.append(np.ascontiguousarray(c[:, :3], np.float32))
    scan = synth.plane_scan(8192, 30.0, -synth.GROUND_Z, rng)
    R0, t0 = synth.perturbed_start(np.eye(3), np.zeros(3))
    np.array([c.shape[0] for c in chunks], np.int64).tofile(os.path.join(out, "c1_chunk_sizes.bin"))
    np.concatenate(chunks).astype("<f4").tofile(os.path.join(out, "c1_chunks_xyz.bin"))
    np.ascontiguousarray(scan[:, :3], "<f4").tofile(os.path.join(out, "c1_scan_xyz.bin"))
    np.concatenate([R0.ravel(), t0, [5.0, 1.0]]).astype("<f8").tofile(os.path.join(out, "c1_pose0.bin"))
    print("wrote", out, "chunks", len(chunks), "map", m.size())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "c1_ref_io")
