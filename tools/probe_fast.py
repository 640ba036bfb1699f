#!/usr/bin/env python3
"""Order-free search (csrc/bvh4.cuh) vs the reference-order replay, on the GPU: ray-by-ray equality of
mcrt_trace_closest on incoherent rays, image / ray-count equality of full renders, stage times.
  python tools/probe_fast.py <pack> [--width W --height H --sqrtspp S] [--rays N] [--lib suffix]"""
import argparse, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("pack"); ap.add_argument("--width", type=int); ap.add_argument("--height", type=int)
ap.add_argument("--sqrtspp", type=int); ap.add_argument("--rays", type=int, default=2_000_000)
ap.add_argument("--reps", type=int, default=2); ap.add_argument("--pool", type=float, default=float(1 << 24))
ap.add_argument("--skip-exact-render", action="store_true"); ap.add_argument("--tag", default="")
ap.add_argument("--skip-trace", action="store_true"); ap.add_argument("--opt", action="append", default=[], help="key=value option")
a = ap.parse_args()
m = importlib.import_module("monte-carlo-ray-tracer_b200")
scene = m.Scene.from_pack(a.pack)
cam = scene.cameras()[0]
cam = cam.resized(a.width or cam.width, a.height or cam.height, a.sqrtspp or cam.sqrtspp)
pt = m.PathTracer(scene, precision=m.PRECISION_F64)
pt.set_option("pool_paths", a.pool); pt.set_option("stage_timing", 1)
for kv in a.opt:
    k_, v_ = kv.split("="); pt.set_option(k_, float(v_))
out = {"pack": os.path.basename(a.pack), "tag": a.tag, "lib": os.environ.get("MCRT_LIB", "")}

# ---- incoherent rays: segments between points on the surfaces the camera sees, plus the camera rays
if a.skip_trace:
    a.rays = 1000
rng = np.random.default_rng(5)
n = a.rays
px = rng.integers(0, cam.width * cam.height, n // 4).astype(np.uint32)
b = np.asarray(scene.a["node_bounds"][:6] if scene.n_nodes else [-1, -1, -1, 1, 1, 1], dtype=np.float64)
lo, hi = b[:3], b[3:]
o = rng.uniform(lo, hi, (n, 3)); t = rng.uniform(lo, hi, (n, 3))
d = t - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = np.concatenate([o, d], axis=1)
h0 = pt.intersect(rays)
hitp = rays[:, :3] + rays[:, 3:] * np.where(h0["prim"] != m.NO_PRIM, h0["t"], 0.0)[:, None]
# second generation: rays leaving the hit points (what path extension / shadow rays look like)
ok = h0["prim"] != m.NO_PRIM
o2 = hitp[ok] - rays[ok, 3:] * 1e-9
d2 = rng.normal(size=o2.shape); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
rays2 = np.concatenate([o2, d2], axis=1)
allrays = np.concatenate([rays, rays2], axis=0)
t0 = time.time(); hf = pt.intersect(allrays); st_f = dict(pt.last_stats); tf = time.time() - t0
pt.set_option("exact_traversal", 1)
t0 = time.time(); he = pt.intersect(allrays); st_e = dict(pt.last_stats); te = time.time() - t0
pt.set_option("exact_traversal", 0)
same = (hf["prim"] == he["prim"]) & (hf["t"] == he["t"]) & (hf["u"] == he["u"]) & (hf["v"] == he["v"])
out["trace"] = dict(rays=len(allrays), identical=int(same.sum()), mismatches=int((~same).sum()), hit_fraction=float((he["prim"] != m.NO_PRIM).mean()),
                    replayed=int(st_f["replayed_rays"]), fast_gpu_ms=st_f["gpu_ms_total"], exact_gpu_ms=st_e["gpu_ms_total"],
                    fast_box_per_ray=st_f["box_tests"] / len(allrays), fast_prim_per_ray=st_f["prim_tests"] / len(allrays),
                    exact_box_per_ray=st_e["box_tests"] / len(allrays), exact_prim_per_ray=st_e["prim_tests"] / len(allrays))
print("trace", json.dumps(out["trace"]), flush=True)
if (~same).any():
    bad = np.nonzero(~same)[0][:5]
    for i in bad:
        print("  mismatch ray", i, "fast", hf["prim"][i], hf["t"][i], "exact", he["prim"][i], he["t"][i], flush=True)

# ---- renders
def render(tag):
    best = None
    for r in range(a.reps):
        img = pt.render_rows(cam); st = dict(pt.last_stats)
        if best is None or st["gpu_ms_total"] < best[1]["gpu_ms_total"]:
            best = (img, st)
    img, st = best
    rays_ = st["extension_rays"] + st["shadow_rays"]
    rec = dict(gpu_ms=st["gpu_ms_total"], mray_s=rays_ / st["gpu_ms_total"] / 1e3, ext=st["extension_rays"], sh=st["shadow_rays"],
               box_per_ray=st["box_tests"] / rays_, prim_per_ray=st["prim_tests"] / rays_, replayed=st["replayed_rays"],
               ms_extend=st["gpu_ms_extend"], ms_shade=st["gpu_ms_shade"], ms_shadow=st["gpu_ms_shadow"], ms_gen=st["gpu_ms_generate"],
               iters=st["wavefront_iterations"], mean=float(img.mean()))
    print(tag, f"{cam.width}x{cam.height}x{cam.sqrtspp**2}", json.dumps(rec), flush=True)
    return img, rec
img_f, out["render_fast"] = render("render fast ")
if not a.skip_exact_render:
    pt.set_option("exact_traversal", 1)
    img_e, out["render_exact"] = render("render exact")
    pt.set_option("exact_traversal", 0)
    out["render_max_abs_diff"] = float(np.abs(img_f - img_e).max())
    out["render_rel_rmse"] = float(np.sqrt(np.mean((img_f - img_e) ** 2)) / max(1e-300, np.abs(img_e).mean()))
    out["ray_counts_equal"] = bool(out["render_fast"]["ext"] == out["render_exact"]["ext"] and out["render_fast"]["sh"] == out["render_exact"]["sh"])
    print("fast vs exact render: max|d|", out["render_max_abs_diff"], "rel rmse", out["render_rel_rmse"], "ray counts equal", out["ray_counts_equal"], flush=True)
pt.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "probe_fast.jsonl"), "a") as f:
    f.write(json.dumps(out) + "\n")
