#!/usr/bin/env python3
"""Photon octree construction at water_caustics scale: GPU (mcrt_octree_build) vs the scalar CPU
restatement in oracle/ (checker only), array-for-array, with timings. Synthetic photons: a thin caustic sheet + a volume-filling cloud."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mcrt = importlib.import_module("monte-carlo-ray-tracer_b200")
from oracle import port  # noqa: E402

if __name__ == "__main__":
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 12_000_000
    rng = np.random.default_rng(11)
    ph = np.zeros((n, 8), dtype=np.float32)
    half = n // 2
    ph[:half, 3:6] = rng.uniform(-3, 3, (half, 3))
    xy = rng.normal(0, 0.8, (n - half, 2))
    ph[half:, 3] = xy[:, 0]; ph[half:, 5] = xy[:, 1]; ph[half:, 4] = -2.0 + 0.02 * np.sin(7 * xy[:, 0]) * np.cos(5 * xy[:, 1])
    ph[:, :3] = 0.1; ph[:, 6:] = 1.0
    bounds = (-3.5, -3.5, -3.5, 3.5, 3.5, 3.5)
    t0 = time.perf_counter(); host = port.build_photon_octree(ph, 200, bounds, mcrt.PhotonMapDesc, mcrt._map_arrays); t_host = time.perf_counter() - t0
    best = None
    for _ in range(3):
        t0 = time.perf_counter(); gpu, ms = mcrt.build_photon_octree(ph, 200, bounds); wall = time.perf_counter() - t0
        best = ms if best is None else min(best, ms)
    same = all(np.array_equal(host[k], gpu[k]) for k in ("octant_bounds", "octant_start", "octant_count", "octant_next", "octant_leaf", "photons"))
    res = dict(photons=n, octants=int(gpu["octant_leaf"].size), identical=bool(same), gpu_build_ms=best, oracle_cpu_build_s=t_host,
               gpu_call_wall_s_incl_h2d_d2h=wall)
    print(json.dumps(res))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "octree_build.json"), "w"), indent=1)
