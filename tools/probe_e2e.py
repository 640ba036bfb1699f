#!/usr/bin/env python3
"""Where does the host-buffer (e2e) call spend its time? upload / render / copy, next to the resident step."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("monte-carlo-ray-tracer_b200")
import torch
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scene = m.Scene.from_pack(os.path.join(ROOT, "bench_data", "c2_hexagon_room.mcrtpack"))
cam = scene.cameras()[0].resized(1920, 1080, spp)
pt = m.PathTracer(scene, precision=m.PRECISION_F64)
pt.set_option("pool_paths", 1 << 24); pt.set_option("stage_timing", 1)
fb = torch.empty((1080, 1920, 3), dtype=torch.float64, device="cuda")
host = torch.empty((1080, 1920, 3), dtype=torch.float64).pin_memory().numpy()
for _ in range(3):
    st = pt.render_rows_dev(cam, fb.data_ptr())
print("resident: gpu_ms", round(st["gpu_ms_total"], 1), "shade", round(st["gpu_ms_shade"], 1), "extend", round(st["gpu_ms_extend"], 1), "launches", st["kernel_launches"])
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pt.upload_scene(); t1 = time.perf_counter()
    pt.render_rows(cam, 0, 1080, out=host); t2 = time.perf_counter()
    st = pt.last_stats
    print(f"e2e rep{rep}: upload {1e3 * (t1 - t0):.1f} ms, render call {1e3 * (t2 - t1):.1f} ms (gpu_ms {st['gpu_ms_total']:.1f}, shade {st['gpu_ms_shade']:.1f}, extend {st['gpu_ms_extend']:.1f}, "
          f"launches {st['kernel_launches']}, iters {st['wavefront_iterations']})")
for _ in range(2):
    st = pt.render_rows_dev(cam, fb.data_ptr())
print("resident again: gpu_ms", round(st["gpu_ms_total"], 1), "shade", round(st["gpu_ms_shade"], 1))
pt.close()
