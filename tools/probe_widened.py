#!/usr/bin/env python3
"""Timings of the widened-scope stages next to the reference's CPU code (run on the GPU box):
  * image pipeline: mcrt_image_tonemap (host buffers) / _dev (device buffers) vs Image::save of the
    unmodified reference (oracle/_ref) on the same HDR frame;
  * film filters: C2 frame through the Mitchell-Netravali film vs the default box film.
Writes gpurun_out/widened.json."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mcrt = importlib.import_module("monte-carlo-ray-tracer_b200")

if __name__ == "__main__":
    import torch
    res = {}
    scene = mcrt.Scene.from_pack(os.path.join(ROOT, "bench_data", "c2_hexagon_room.mcrtpack"))
    pt = mcrt.PathTracer(scene, precision=mcrt.PRECISION_F64)
    for (w, h) in ((1920, 1080), (3840, 2160)):
        cam = scene.cameras()[0].resized(w, h, 2)
        fb = torch.empty((h, w, 3), dtype=torch.float64, device="cuda")
        out = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
        pt.render_rows_dev(cam, fb.data_ptr())
        torch.cuda.synchronize()
        img = fb.cpu().numpy()
        for _ in range(2):
            pt.tonemap_dev(fb.data_ptr(), out.data_ptr(), w, h, {})
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            e, g = pt.tonemap_dev(fb.data_ptr(), out.data_ptr(), w, h, {})
        torch.cuda.synchronize()
        dev_ms = (time.perf_counter() - t0) * 100
        t0 = time.perf_counter(); b_host, e2, g2 = pt.tonemap(img, {}); host_ms = (time.perf_counter() - t0) * 1e3
        entry = dict(gpu_dev_ms=dev_ms, gpu_host_buffers_ms=host_ms, exposure=e, gain=g,
                     same_bytes_dev_vs_host=bool(np.array_equal(out.cpu().numpy(), b_host)))
        try:
            from oracle import ref
            t0 = time.perf_counter(); b_ref, er, gr = ref.image_save(img, {}); entry["reference_cpu_ms"] = (time.perf_counter() - t0) * 1e3
            entry["bytes_differing_from_reference"] = int(np.count_nonzero(b_ref != b_host))
            entry["factors_equal"] = bool((er, gr) == (e2, g2))
        except Exception as ex:   # no oracle/_ref on this machine
            entry["reference"] = f"unavailable: {ex}"
        res[f"image_{w}x{h}"] = entry
        print(f"image {w}x{h}:", json.dumps(entry), flush=True)
    cam = scene.cameras()[0].resized(1920, 1080, 4)
    for name, film in (("box", None), ("mitchell", dict(filter="mitchell-netravali")), ("gaussian_cached", dict(filter="gaussian", cache_size=256))):
        cam.film = film
        fb = torch.empty((1080, 1920, 3), dtype=torch.float64, device="cuda")
        best = None
        for _ in range(3):
            st = pt.render_rows_dev(cam, fb.data_ptr())
            ms = st["gpu_ms_total"]
            best = ms if best is None else min(best, ms)
        rays = st["extension_rays"] + st["shadow_rays"]
        res[f"film_{name}"] = dict(gpu_ms=best, mray_s=rays / best / 1e3)
        print(f"film {name}: {best:.1f} ms, {rays / best / 1e3:.0f} Mray/s", flush=True)
    pt.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "widened.json"), "w"), indent=1)
