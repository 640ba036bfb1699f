#!/usr/bin/env python3
"""Developer probe: render a pack at a given size in both precisions, print throughput + stats."""
import argparse, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("monte-carlo-ray-tracer_b200")

ap = argparse.ArgumentParser()
ap.add_argument("pack"); ap.add_argument("--width", type=int); ap.add_argument("--height", type=int)
ap.add_argument("--sqrtspp", type=int); ap.add_argument("--modes", default="f64,f32")
ap.add_argument("--pool", type=float); ap.add_argument("--bps", type=float); ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--compare", action="store_true")
ap.add_argument("--opt", action="append", default=[], help="key=value option")
a = ap.parse_args()
scene = m.Scene.from_pack(a.pack)
cam = scene.cameras()[0]
cam = cam.resized(a.width or cam.width, a.height or cam.height, a.sqrtspp or cam.sqrtspp)
imgs = {}
for mode in a.modes.split(","):
    prec = m.PRECISION_F64 if mode == "f64" else m.PRECISION_F32
    pt = m.PathTracer(scene, precision=prec)
    if a.pool: pt.set_option("pool_paths", a.pool)
    if a.bps: pt.set_option("blocks_per_sm", a.bps)
    pt.set_option("stage_timing", 1)
    for kv in a.opt:
        k, v = kv.split("="); pt.set_option(k, float(v))
    for r in range(a.reps):
        t = time.time(); img = pt.render_rows(cam); wall = time.time() - t
        st = pt.last_stats
        rays = st["extension_rays"] + st["shadow_rays"]
        print(f"{mode} rep{r}: {cam.width}x{cam.height}x{cam.sqrtspp**2}spp gpu_ms={st['gpu_ms_total']:.1f} wall={wall*1e3:.1f} "
              f"Mray/s={rays/st['gpu_ms_total']/1e3:.1f} paths={st['paths']} ext={st['extension_rays']} sh={st['shadow_rays']} "
              f"box/ray={st['box_tests']/rays:.1f} prim/ray={st['prim_tests']/rays:.1f} iters={st['wavefront_iterations']} "
              f"launches={st['kernel_launches']} maxdepth={st['max_depth']} mean={img.mean():.6f} "
              f"tail_util={st['extend_work_sum']/max(1,st['extend_work_warpmax']):.3f} stages ext={st['gpu_ms_extend']:.1f} shade={st['gpu_ms_shade']:.1f} shadow={st['gpu_ms_shadow']:.1f} gen={st['gpu_ms_generate']:.1f}", flush=True)
    imgs[mode] = img
    pt.close()
if a.compare and len(imgs) == 2:
    x, y = imgs["f64"], imgs["f32"]
    print("f32 vs f64: rmse", float(np.sqrt(np.mean((x - y) ** 2))), "mean rel diff", float((y.mean() - x.mean()) / x.mean()))
