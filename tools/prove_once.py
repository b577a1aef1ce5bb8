#!/usr/bin/env python
"""Profiling helper: `python tools/prove_once.py WORKLOAD [COUNT]` proves COUNT device-resident synthetic traces.
Meant to be wrapped by ncu (B200_PROFILING.md); prints per-stage CUDA-event times of the last proof."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import zk_stark_project_b200 as Z  # noqa: E402
from zk_stark_project_b200 import lib as L  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "training_2p16"
count = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kind, n, w, beta, _ = bench.WORKLOADS[name]
ctx = L.Context(0)
air, data, opts = bench.build_workload(name, 0x5EED0000)
if kind == "mimc":
    raw = ctx.mimc_trace([j + 1 for j in range(w)], n, Z.get_round_constants())
    data = np.frombuffer(raw, dtype=np.uint64).reshape(w, n, 2)
    get = lambda c, r: int(data[c, r, 0]) | (int(data[c, r, 1]) << 64)
    air = bench.mimc_air(opts, w, n, [get(j, 0) for j in range(w)], [get(j, n - 1) for j in range(w)])
buf = np.ascontiguousarray(data)
d = ctx.upload_trace(buf.ctypes.data, w, n)
air = ctx.prepare(air)
l0 = 0
for i in range(count):
    l0 = ctx.launches()
    proof, ts = ctx.prove_device(air, d)
print(json.dumps({"workload": name, "launches_per_proof": ctx.launches() - l0, "stages_ms": ctx.stage_times(), "proof_bytes": len(proof)}))
