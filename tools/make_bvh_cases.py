#!/usr/bin/env python3
"""Reference BVHs of the large OBJ scenes, for the GPU builder's parity tests and timing
(mcrt_bvh_build). Run in the build container (needs /root/reference).

For every scene: Surface::BB() of Scene::surfaces (the builder's input) and, for each BVH type, the
arrays of the tree BVH::BVH built from them (oracle/ref_driver.cpp: ref_bvh_build) plus the CPU
build time. Output: bench_data/bvh_<scene>.npz (git-ignored; travels to the GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

SCENES = {
    "spaceship": ("spaceship.json", [("quaternary_sah", 0), ("binary_sah", 0), ("octree", 0), ("quaternary_sah", 4),
                                     ("binary_sah", 32)]),
    "lego_bulldozer": ("lego_bulldozer.json", [("quaternary_sah", 0), ("binary_sah", 0)]),
}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(SCENES)):
        scene_file, types = SCENES[name]
        s = ref.RefScene(scene_file, dict(width=8, height=8, sqrtspp=1, bvh_type="none"))
        bounds, scene_bounds = s.prim_bounds()
        out = dict(prim_bounds=bounds, scene_bounds=scene_bounds, cases=np.array([f"{t}:{b}" for t, b in types]))
        for t, b in types:
            r = s.build_bvh(t, b)
            key = f"{t}:{b}"
            for k in ("node_bounds", "node_first_prim", "node_prim_count", "node_next_sibling", "prim_order"):
                out[f"{key}/{k}"] = r[k]
            out[f"{key}/cpu_seconds"] = np.float64(r["seconds"])
            print(f"{name} {key}: {len(r['node_first_prim'])} nodes, reference CPU build {r['seconds']:.2f} s")
        s.close()
        path = os.path.join(ROOT, "bench_data", f"bvh_{name}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) >> 20, "MiB")
