"""mimosa_b200 — B200-native LiDAR geometric-factor (scan-to-map point-to-plane ICP) path.

The product is the C-ABI shared library (include/mimosa_b200.h, mimosa_b200/csrc/); `host` mirrors the
reference's C++ interface for the Python test/benchmark harness.  No CPU fallback exists.
"""
from .capi import CloudLayout, CloudOrder, InputFilter  # noqa: F401
from .host import (  # noqa: F401
    HORNBILL_MAP,
    Context,
    ICPFactor,
    IncrementalVoxelMap,
    RegistrationConfig,
    Scan,
    degeneracy_flags,
    gn_step,
    hornbill_config,
    shard_range,
)
