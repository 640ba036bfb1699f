#!/usr/bin/env python3
"""Parity fixtures for the OBJ scenes of BASELINE.json (configs 3-5).
  make  (build container, needs /root/reference): reference loader/BVH builder/photon pass -> scene pack
        (bench_data/v*_*.mcrtpack.xz: 35-81 MB, git-ignored, shipped to the GPU box by gpurun; the same packs
        are bench.py's c3/c4/c5 workloads with the camera resized) + reference outputs at reduced resolution
        (tests/golden/big/*.npz: committed; read by tests/test_gpu_big_scenes.py)
  check (GPU box): render the same packs on the GPU in parity mode and print the comparison."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "bench_data")
GOLD = os.path.join(ROOT, "tests", "golden", "big")
SEED = 0x12345678
CASES = {
    "v3_spaceship": ("spaceship.json", dict(width=160, height=90, sqrtspp=2), False),
    "v4_water_caustics": ("water_caustics.json", dict(width=96, height=96, sqrtspp=2, emissions=1e5, num_render_threads=1), True),
    "v5_lego_bulldozer": ("lego_bulldozer.json", dict(width=128, height=72, sqrtspp=2), False),
}

def make(which):
    from oracle import ref
    rng = np.random.default_rng(7)
    for cid in which:
        scene, ov, pm = CASES[cid]
        t0 = time.time()
        ref.set_seed(SEED)
        s = ref.RefScene(scene, ov, photon_map=pm)
        s.export_pack(os.path.join(OUT, cid + ".mcrtpack"))
        os.system(f"xz -1 -T8 -f {os.path.join(OUT, cid + '.mcrtpack')}")
        img, sec, rays, sh = s.render(threads=8)
        n = 4096
        px = rng.integers(0, s.width * s.height, n).astype(np.uint32); sm = rng.integers(0, s.sqrtspp ** 2, n).astype(np.uint32)
        rgb, r6 = s.sample_pixels(px, sm)
        # traversal KAT: the camera rays + incoherent segments between jittered primary hit points
        tc, prim0, _, _ = s.trace(r6)
        hit = prim0 != 0xFFFFFFFF
        pts = r6[hit, :3] + r6[hit, 3:] * tc[hit, None]
        scale = float(np.abs(pts).max()) if len(pts) else 1.0
        a = pts[rng.integers(0, len(pts), 8192)] + rng.normal(0, 0.01 * scale, (8192, 3))
        b = pts[rng.integers(0, len(pts), 8192)] + rng.normal(0, 0.01 * scale, (8192, 3))
        d = b - a; nrm = np.linalg.norm(d, axis=1, keepdims=True); ok = nrm[:, 0] > 1e-9
        tr_rays = np.concatenate([r6, np.concatenate([a[ok], d[ok] / nrm[ok]], axis=1)], axis=0)
        t, prim, uv, ip = s.trace(tr_rays)
        extra = {}
        if pm:
            q = pts[rng.integers(0, len(pts), 512)] + rng.normal(0, 0.002 * scale, (512, 3))
            for which, name in ((0, "caustic"), (1, "global")):
                ph, d2, cnt = s.knn(which, q, 50)
                extra[f"knn_{name}_d2"] = d2; extra[f"knn_{name}_count"] = cnt; extra[f"knn_{name}_pos"] = ph[:, :, 3:6]
            extra["knn_points"] = q; extra["knn_k"] = np.uint32(50)
        os.makedirs(GOLD, exist_ok=True)
        np.savez_compressed(os.path.join(GOLD, cid + ".npz"), image=img, total_rays=np.uint64(rays), shadow_rays=np.uint64(sh),
                            ps_pixel=px, ps_sample=sm, ps_rays=r6, ps_rgb=rgb, tr_rays=tr_rays, tr_t=t, tr_prim=prim, tr_uv=uv,
                            tr_interp=ip, seed=np.uint32(SEED), n_prims=np.uint32(s.n_prims), n_lights=np.uint32(s.n_lights),
                            width=np.uint32(s.width), height=np.uint32(s.height), sqrtspp=np.uint32(s.sqrtspp), **extra)
        print(f"   interp fraction of traversal hits: {ip[prim != 0xFFFFFFFF].mean():.3f}")
        print(f"{cid}: prims={s.n_prims} nodes={s.n_nodes} lights={s.n_lights} rays={rays} mean={img.mean():.5f} "
              f"pack={os.path.getsize(os.path.join(OUT, cid + '.mcrtpack.xz')) >> 20} MiB  ({time.time() - t0:.0f} s)", flush=True)
        s.close()

def check(which):
    m = importlib.import_module("monte-carlo-ray-tracer_b200")
    lines = []
    for cid in which:
        pack = os.path.join(OUT, cid + ".mcrtpack.xz")
        if not os.path.exists(pack):
            continue
        g = np.load(os.path.join(GOLD, cid + ".npz"))
        scene = m.Scene.from_pack(pack)
        cls = m.PhotonMapper if scene.photon_maps() is not None else m.PathTracer
        pt = cls(scene, precision=m.PRECISION_F64, global_seed=int(g["seed"]))
        cam = scene.cameras()[0]
        img = pt.render_rows(cam); st = pt.last_stats
        rmse = float(np.sqrt(np.mean((img - g["image"]) ** 2)))
        hits = pt.intersect(g["ps_rays"])
        same_prim = float(np.mean(hits["prim"] == g["tr_prim"]))
        hit = g["tr_prim"] != m.NO_PRIM
        dt = float(np.max(np.abs(hits["t"][hit] - g["tr_t"][hit]) / np.maximum(1.0, g["tr_t"][hit]))) if hit.any() else 0.0
        rgb = pt.sampleRay(g["ps_rays"], g["ps_pixel"], g["ps_sample"])
        err = np.abs(rgb - g["ps_rgb"]) / np.maximum(1.0, np.abs(g["ps_rgb"]))
        line = (f"{cid}: prims={scene.n_prims} image {cam.width}x{cam.height}x{cam.sqrtspp**2}spp RMSE={rmse:.3e} (mean {g['image'].mean():.4f}) "
                f"ext_rays gpu={st['extension_rays']} ref={int(g['total_rays']) - int(g['shadow_rays'])} "
                f"intersect same_prim={same_prim:.6f} max_rel_dt={dt:.2e} sampleRay max_rel_err={err.max():.2e} "
                f"(#>1e-9: {int((err > 1e-9).sum())} of {err.size}) max_depth={st['max_depth']} ior_overflows={st['ior_stack_overflows']}")
        print(line, flush=True); lines.append(line)
        pt.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "validate_big.txt"), "w").write("\n".join(lines) + "\n")

if __name__ == "__main__":
    mode = sys.argv[1]; which = sys.argv[2:] or list(CASES)
    (make if mode == "make" else check)(which)
