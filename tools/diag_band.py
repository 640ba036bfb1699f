#!/usr/bin/env python3
"""Diagnostic: which (pixel, sample) of the C2 benchmark rows differs from the reference band, and by what.
Uses the CPU oracle (test infrastructure) for per-sample values: oracle == reference bit for bit on this band."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import port
m = importlib.import_module("monte-carlo-ray-tracer_b200")
k = np.load(os.path.join(ROOT, "tests", "golden", "c2_band_kat.npz"))
scene = m.Scene.from_pack(os.path.join(ROOT, "bench_data", "c2_hexagon_room.mcrtpack"))
cam = scene.cameras()[0]
seed = int(k["seed"]); y0, y1 = int(k["y0"]), int(k["y1"])
pt = m.PathTracer(scene, precision=m.PRECISION_F64, global_seed=seed)
ps = port.PortScene(scene)
out = {}
for exact in (0, 1):
    pt.set_option("exact_traversal", exact)
    band = pt.render_rows(cam, y0, y1)
    d = np.abs(band - k["band"])
    bad = np.argwhere(d.max(axis=2) > 1e-10)
    print(f"exact={exact}: max diff {d.max():.3e}, bad pixels {len(bad)}: {bad[:10].tolist()}", flush=True)
    out[f"band_{exact}"] = band
    for (r, c) in bad[:4]:
        pixel = (y0 + r) * cam.width + c
        spp = cam.sqrtspp ** 2
        px = np.full(spp, pixel, np.uint32); sm = np.arange(spp, dtype=np.uint32)
        rgb_ref, rays = ps.sample_pixels(cam, px, sm, seed)
        rgb_gpu = pt.sampleRay(rays, px, sm)
        ds = np.abs(rgb_gpu - rgb_ref).max(axis=1)
        bs = np.nonzero(ds > 1e-12)[0]
        print(f"  pixel ({y0 + r},{c}) = {pixel}: pixel diff {d[r, c]}, samples differing {bs.tolist()}", flush=True)
        for s_ in bs[:4]:
            print(f"    sample {s_}: gpu {rgb_gpu[s_]} ref {rgb_ref[s_]} ray {rays[s_].tolist()}", flush=True)
        out[f"bad_{exact}_{pixel}_rays"] = rays; out[f"bad_{exact}_{pixel}_gpu"] = rgb_gpu; out[f"bad_{exact}_{pixel}_ref"] = rgb_ref
pt.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "diag_band.npz"), **out)
