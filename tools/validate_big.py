#!/usr/bin/env python3
"""One-off parity validation on the OBJ scenes of BASELINE.json (configs 3-5), too large to commit.
  make  (build container, needs /root/reference): reference loader/BVH builder/photon pass -> scene pack
        + reference outputs at reduced resolution, written to bench_data/ (git-ignored, shipped by gpurun)
  check (GPU box): render the same packs on the GPU in parity mode and compare."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "bench_data")
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
        t, prim, uv, ip = s.trace(r6)
        np.savez_compressed(os.path.join(OUT, "validate_" + cid + ".npz"), image=img, total_rays=np.uint64(rays), shadow_rays=np.uint64(sh),
                            ps_pixel=px, ps_sample=sm, ps_rays=r6, ps_rgb=rgb, tr_t=t, tr_prim=prim, seed=np.uint32(SEED),
                            width=np.uint32(s.width), height=np.uint32(s.height), sqrtspp=np.uint32(s.sqrtspp))
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
        g = np.load(os.path.join(OUT, "validate_" + cid + ".npz"))
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
