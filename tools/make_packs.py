#!/usr/bin/env python3
"""Writes the scene packs bench.py renders (BASELINE.json configs) into bench_data/ using the
reference's own loader + BVH builders through oracle/_ref (container only: needs /root/reference).
Small packs (inline geometry) are committed; large ones (OBJ scenes) are git-ignored but travel
to the GPU box with the gpurun snapshot."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "bench_data")

CONFIGS = {
    # id: (scene, overrides, photon_map)
    "c1_hexagon_room_diffuse": ("hexagon_room_diffuse.json", dict(width=256, height=256, sqrtspp=2, bvh_type="binary_sah", bins_per_axis=16), False),
    "c2_hexagon_room": ("hexagon_room.json", dict(width=1920, height=1080, sqrtspp=16, bvh_type="quaternary_sah"), False),
    "c3_spaceship": ("spaceship.json", dict(width=1920, height=1080, sqrtspp=32), False),
    "c4_water_caustics": ("water_caustics.json", dict(width=1024, height=1024, sqrtspp=23, emissions=1e6), True),
    "c5_lego_bulldozer": ("lego_bulldozer.json", dict(width=3840, height=2160, sqrtspp=64), False),
}

if __name__ == "__main__":
    which = sys.argv[1:] or ["c1_hexagon_room_diffuse", "c2_hexagon_room"]
    os.makedirs(OUT, exist_ok=True)
    for cid in which:
        scene, ov, pm = CONFIGS[cid]
        t = time.time()
        s = ref.RefScene(scene, ov, photon_map=pm)
        path = s.export_pack(os.path.join(OUT, cid + ".mcrtpack"))
        if os.path.getsize(path) > (8 << 20):   # OBJ scenes: ship compressed (gpurun snapshots are capped at 512 MiB)
            os.system(f"xz -1 -T8 -f {path}")
            path += ".xz"
        print(cid, "prims", s.n_prims, "nodes", s.n_nodes, "lights", s.n_lights, os.path.getsize(path), "bytes", round(time.time() - t, 1), "s")
        s.close()
