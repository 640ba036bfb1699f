#!/usr/bin/env python3
"""Developer probe: GPU photon pass (mcrt_photon_emit) + photon-mapped render on a pack."""
import argparse, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("monte-carlo-ray-tracer_b200")
ap = argparse.ArgumentParser()
ap.add_argument("pack"); ap.add_argument("--emissions", type=float, default=1e6); ap.add_argument("--caustic-factor", type=float, default=10.0)
ap.add_argument("--width", type=int, default=1024); ap.add_argument("--height", type=int, default=1024); ap.add_argument("--sqrtspp", type=int, default=4)
ap.add_argument("--modes", default="f64"); ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
scene = m.Scene.from_pack(a.pack)
params = scene.extra["photon_emit_params"]
cam = scene.cameras()[0].resized(a.width, a.height, a.sqrtspp)
for mode in a.modes.split(","):
    prec = m.PRECISION_F64 if mode == "f64" else m.PRECISION_F32
    t = time.time()
    pm = m.PhotonMapper(scene, precision=prec, emit=dict(emissions=int(a.emissions), caustic_factor=a.caustic_factor,
                        max_photons_per_octree_leaf=int(params[2]), k_nearest_photons=50, scene_bounds=params[3:9]))
    wall = time.time() - t
    st = pm.last_stats
    nc, ng = pm.n_photons
    print(f"{mode} emit: emissions={int(a.emissions * a.caustic_factor)} photon_rays={st['extension_rays']} emission_gpu_ms={st['gpu_ms_total']:.1f} "
          f"({st['extension_rays'] / st['gpu_ms_total'] / 1e3:.1f} Mray/s) octree_build_gpu_ms={st['gpu_ms_knn']:.1f} "
          f"wall_whole_photon_pass={wall:.3f}s caustic={nc} global={ng} iters={st['wavefront_iterations']}", flush=True)
    pm.set_option("stage_timing", 1)
    for r in range(a.reps):
        img = pm.render_rows(cam); st = pm.last_stats
        rays = st["extension_rays"] + st["shadow_rays"]
        print(f"{mode} pm render rep{r}: {cam.width}x{cam.height}x{cam.sqrtspp**2}spp gpu_ms={st['gpu_ms_total']:.1f} Mray/s={rays / st['gpu_ms_total'] / 1e3:.1f} "
              f"knn_queries={st['knn_queries']} ({st['knn_queries'] / st['gpu_ms_total'] / 1e3:.1f} Mquery/s) paths={st['paths']} "
              f"stages ext={st['gpu_ms_extend']:.1f} shade+knn={st['gpu_ms_shade']:.1f} (knn {st['gpu_ms_knn']:.1f} = {st['knn_queries'] / max(1e-9, st['gpu_ms_knn']) / 1e3:.1f} Mquery/s in-kernel) shadow={st['gpu_ms_shadow']:.1f} mean={img.mean():.5f}", flush=True)
    pm.close()
