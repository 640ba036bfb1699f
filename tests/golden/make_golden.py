#!/usr/bin/env python3
"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref, built by
oracle/build_ref.py from /root/reference). Run in the build container only; the GPU box has no
/root/reference and just reads the files this script wrote.

For every case: a scene pack (flattened Scene/BVH/materials/lights/camera as the reference built
them) and an .npz with reference outputs at the pinned seed:
    image        Film::scan of a full Camera render (1 thread)                     [H, W, 3] f64
    ps_pixel/ps_sample/ps_rays/ps_rgb
                 random (pixel, sample) pairs: camera ray of Camera::samplePixel and the radiance
                 Integrator::sampleRay returned for it
    tr_rays/tr_t/tr_prim/tr_uv/tr_interp
                 Scene::intersect on camera rays + random rays through the scene bounds
    total_rays/shadow_rays   Scene::intersect call counts of the full render
plus sampler_kat.npz (Sampler streams), bsdf_kat.npz (Fresnel / GGX values) and film_kat.npz
(Film::deposit / Film::scan with every reconstruction filter: full renders of one scene).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402

SEED = 0x12345678

# id -> (scene json, overrides, photon_map)
CASES = {
    # BASELINE config 1 exactly: hexagon_room_diffuse 256x256, 4 spp, binary_sah/16
    "c1_hexagon_diffuse_256": ("hexagon_room_diffuse.json",
                               dict(width=256, height=256, sqrtspp=2, bvh_type="binary_sah", bins_per_axis=16), False),
    # BASELINE config 2 scene/BVH, down-scaled: glass, scene ior 1.75, triangle quad light
    "c2_hexagon_room_96": ("hexagon_room.json",
                           dict(width=96, height=54, sqrtspp=2, bvh_type="quaternary_sah"), False),
    # octree BVH (the shipped default) on the same scene
    "hexagon_room_octree_64": ("hexagon_room.json", dict(width=64, height=48, sqrtspp=1), False),
    "oren_nayar_64": ("oren_nayar_test.json", dict(width=64, height=48, sqrtspp=2), False),
    "ggx_64": ("ggx_test.json", dict(width=64, height=48, sqrtspp=2), False),
    "ior_test_nobvh_64": ("ior_test.json", dict(width=64, height=48, sqrtspp=2), False),
    "quadric_64": ("quadric.json", dict(width=64, height=48, sqrtspp=2), False),
    "veach_mis_64": ("veach_mis.json", dict(width=64, height=48, sqrtspp=2), False),
    "metals_64": ("metals.json", dict(width=64, height=48, sqrtspp=2), False),
    # photon-mapped render (PhotonMapper::sampleRay + k-NN estimates); maps built by the reference's
    # CPU photon pass with 1 thread, shipped inside the pack
    "pm_hexagon_room_64": ("hexagon_room.json", dict(width=64, height=48, sqrtspp=2, emissions=4000,
                                                      num_render_threads=1), True),
    # the OBJ scene written for this repository (tests/golden/make_mesh_scene.py): interpolated vertex
    # normals (generated and from the file), the shading-normal fall-back, GGX / complex-IOR / glass on
    # meshes, a 320-triangle mesh light (light CDF with 320 entries)
    "smooth_mesh_64": ("smooth_mesh.json", dict(), False, os.path.join(HERE, "scenes")),
}


def make_case(cid, scene_file, overrides, photon_map, rng, scenes=None):
    ref.set_seed(SEED)
    s = ref.RefScene(scene_file, overrides, photon_map=photon_map, scenes=scenes)
    pack = os.path.join(HERE, cid + ".mcrtpack")
    s.export_pack(pack)

    image, _, total_rays, shadow_rays = s.render(threads=1)

    n_ps = 4096
    spp = s.sqrtspp ** 2
    ps_pixel = rng.integers(0, s.width * s.height, n_ps).astype(np.uint32)
    ps_sample = rng.integers(0, spp, n_ps).astype(np.uint32)
    ps_rgb, ps_rays = s.sample_pixels(ps_pixel, ps_sample)

    # rays for the traversal KAT: the camera rays above + random segments through the camera-visible
    # region (origins/targets jittered around primary hit points)
    t0, prim0, _, _ = s.trace(ps_rays)
    hit = prim0 != 0xFFFFFFFF
    pts = ps_rays[hit, :3] + ps_rays[hit, 3:] * t0[hit, None]
    if len(pts) >= 2:
        a = pts[rng.integers(0, len(pts), 4096)] + rng.normal(0, 0.05, (4096, 3))
        b = pts[rng.integers(0, len(pts), 4096)] + rng.normal(0, 0.05, (4096, 3))
        d = b - a
        nrm = np.linalg.norm(d, axis=1, keepdims=True)
        ok = nrm[:, 0] > 1e-6
        extra = np.concatenate([a[ok], d[ok] / nrm[ok]], axis=1)
        tr_rays = np.concatenate([ps_rays, extra], axis=0)
    else:
        tr_rays = ps_rays
    tr_t, tr_prim, tr_uv, tr_interp = s.trace(tr_rays)

    extra = {}
    if photon_map:
        # LinearOctree::knnSearch KAT: query points = primary hit points (+ jitter)
        q = pts[rng.integers(0, len(pts), 256)] + rng.normal(0, 0.02, (256, 3))
        k = 50
        for which, name in ((0, "caustic"), (1, "global")):
            ph, d2, cnt = s.knn(which, q, k)
            extra[f"knn_{name}_d2"] = d2
            extra[f"knn_{name}_count"] = cnt
            extra[f"knn_{name}_pos"] = ph[:, :, 3:6]
        extra["knn_points"] = q
        extra["knn_k"] = np.uint32(k)

    np.savez_compressed(os.path.join(HERE, cid + ".npz"), photon_map=np.uint8(photon_map), **extra,
                        seed=np.uint32(SEED), width=np.uint32(s.width), height=np.uint32(s.height),
                        sqrtspp=np.uint32(s.sqrtspp), image=image,
                        ps_pixel=ps_pixel, ps_sample=ps_sample, ps_rays=ps_rays, ps_rgb=ps_rgb,
                        tr_rays=tr_rays, tr_t=tr_t, tr_prim=tr_prim, tr_uv=tr_uv, tr_interp=tr_interp,
                        total_rays=np.uint64(total_rays), shadow_rays=np.uint64(shadow_rays))
    print(f"{cid}: prims={s.n_prims} nodes={s.n_nodes} lights={s.n_lights} rays={total_rays} "
          f"(shadow {shadow_rays}) mean={image.mean():.6f} miss={np.mean(~hit):.3f} "
          f"interpolated hits={int((tr_interp.astype(bool) & (tr_prim != 0xFFFFFFFF)).sum())}")
    s.close()


# name -> the camera's "film" object (source/camera/film.cpp:19-59)
FILMS = {
    "mitchell": dict(filter="mitchell-netravali"),
    "catmull_rom": dict(filter="catmull-rom"),
    "b_spline": dict(filter="b-spline"),
    "hermite": dict(filter="hermite"),
    "gaussian_cached": dict(filter="gaussian", cache_size=256),
    "lanczos_r3": dict(filter="lanczos", radius=3.0),
    "lanczos_cached": dict(filter="Lanczos", radius=2.5, cache_size=64),
    "box_r1p5": dict(filter="box", radius=1.5),
    "box_default_cached": dict(filter="box", cache_size=16),
}


def make_film_kat():
    import json
    out = {}
    for name, film in FILMS.items():
        ref.set_seed(SEED)
        s = ref.RefScene("hexagon_room.json", dict(width=64, height=48, sqrtspp=2, bvh_type="quaternary_sah", film=film))
        if name == "mitchell":
            s.export_pack(os.path.join(HERE, "film_hexagon_room_64.mcrtpack"))
        image, _, _, _ = s.render(threads=1)
        out["image_" + name] = image
        print(f"film_kat {name}: mean={image.mean():.6f} min={image.min():.3e}")
        s.close()
    # photon-mapped render through a filtered film: same scene, overrides and (single-threaded, seeded)
    # photon pass as pm_hexagon_room_64, whose pack holds the photon maps
    ref.set_seed(SEED)
    _, pm_over, _ = CASES["pm_hexagon_room_64"]
    s = ref.RefScene("hexagon_room.json", dict(pm_over, film=FILMS["mitchell"]), photon_map=True)
    tmp = os.path.join(HERE, "_pm_film_tmp.mcrtpack")
    s.export_pack(tmp)
    import importlib
    mcrt = importlib.import_module("monte-carlo-ray-tracer_b200")
    a, b = mcrt.read_pack(tmp), mcrt.read_pack(os.path.join(HERE, "pm_hexagon_room_64.mcrtpack"))
    os.remove(tmp)
    for key in b:
        assert np.array_equal(a[key], b[key]), key     # same scene, same photon maps
    image, _, _, _ = s.render(threads=1)
    out["image_pm_mitchell"] = image
    print(f"film_kat pm_mitchell: mean={image.mean():.6f}")
    s.close()
    np.savez_compressed(os.path.join(HERE, "film_kat.npz"), seed=np.uint32(SEED), films=json.dumps(FILMS), **out)


# the camera's "image" object (image.cpp:10-35) minus the size
IMAGE_CONFIGS = {
    "hable": {},
    "aces": {"tonemapper": "ACES"},
    "plain": {"plain": True},
    "hable_ev": {"exposure_compensation": -0.25, "gain_compensation": 0.5},
    "aces_ev": {"tonemapper": "aces", "exposure_compensation": 1.0},
}


def make_image_kat():
    """Image::save (auto exposure, tone mapping, auto gain, gamma, bytes) of reference renders and of
    synthetic HDR images -> image_kat.npz. Inputs that are golden images are referenced by name."""
    import json
    rng = np.random.default_rng(7)
    inputs = {}
    for cid in ("c2_hexagon_room_96", "veach_mis_64", "metals_64"):
        inputs["golden:" + cid] = np.load(os.path.join(HERE, cid + ".npz"))["image"]
    inputs["hdr_lognormal_128"] = np.exp(rng.normal(-1.0, 2.0, (128, 128, 3)))
    inputs["black_16"] = np.zeros((16, 16, 3))
    sparse = np.zeros((32, 32, 3)); sparse[5, 7] = (3.0, 2.0, 1.0); sparse[20:, :10] = 0.25
    inputs["sparse_32"] = sparse
    out = {}
    for iname, img in inputs.items():
        if not iname.startswith("golden:"):
            out["input/" + iname] = img
        for cname, cfg in IMAGE_CONFIGS.items():
            b, e, g = ref.image_save(img, cfg)
            out[f"bytes/{iname}/{cname}"] = b
            out[f"factors/{iname}/{cname}"] = np.array([e, g])
            print(f"image_kat {iname} {cname}: exposure {e:.6g} gain {g:.6g} mean byte {b.mean():.2f}")
    np.savez_compressed(os.path.join(HERE, "image_kat.npz"), inputs=json.dumps(list(inputs)), configs=json.dumps(IMAGE_CONFIGS), **out)


def make_obj_kat():
    """Scene::parseOBJ / generateVertexNormals on the OBJ fixtures written for this repository
    (tests/golden/quirks.obj) -> obj_kat.npz."""
    s = ref.RefScene("ior_test.json", dict(width=8, height=8, sqrtspp=1))
    r = s.parse_obj(os.path.join(HERE, "quirks.obj"))
    try:
        s.parse_obj(os.path.join(HERE, "quirks_negative.obj"))
        threw = False
    except RuntimeError:
        threw = True
    valid = (r["tri_v"] < len(r["vertices"])).all(axis=1)
    normals, _ = s.vertex_normals(r["vertices"], r["tri_v"][valid])
    s.close()
    np.savez_compressed(os.path.join(HERE, "obj_kat.npz"), vertices=r["vertices"], normals=r["normals"], tri_v=r["tri_v"],
                        tri_vt=r["tri_vt"], tri_vn=r["tri_vn"], negative_throws=np.uint8(threw), generated_normals=normals)
    print("obj_kat:", {k: v.shape for k, v in r.items() if hasattr(v, "shape")}, "negative throws:", threw)


BVH_KAT_TYPES = [("binary_sah", 16), ("quaternary_sah", 8), ("binary_sah", 4), ("quaternary_sah", 3), ("octree", 0)]


def make_bvh_kat():
    """BVH::BVH of the reference on synthetic sphere sets that reach the builders' rare branches
    (arbitrarySplit after a degenerate extent or a failed SAH cost test, the quaternary -> binary
    fall-back) -> bvh_kat.npz. Spheres are the simplest surface whose box is exact (origin +- radius)."""
    import json
    import tempfile
    rng = np.random.default_rng(99)
    template = json.load(open("/root/reference/scenes/ior_test.json"))
    material = template["surfaces"][0]["material"]

    def blobs():
        sets = {}
        sets["random_400"] = (rng.uniform(-5, 5, (400, 3)), rng.uniform(0.05, 0.6, 400))
        sets["coincident_600"] = (np.tile([[1.0, 2.0, 3.0]], (600, 1)), rng.uniform(0.1, 0.5, 600))
        line = np.zeros((700, 3)); line[:, 0] = rng.uniform(-20, 20, 700); line[:, 1] = 0.5; line[:, 2] = -1.25
        sets["line_700"] = (line, rng.uniform(0.05, 0.3, 700))
        cl = np.concatenate([np.tile([[0.0, 0.0, 0.0]], (300, 1)), np.tile([[4.0, 0.5, 0.0]], (300, 1)),
                             np.tile([[0.0, 3.0, -2.0]], (300, 1)), rng.uniform(-6, 6, (100, 3))])
        sets["clusters_1000"] = (cl[rng.permutation(len(cl))], rng.uniform(0.05, 0.4, 1000))
        sets["sah_fail_500"] = (rng.uniform(-1e-3, 1e-3, (500, 3)), rng.uniform(8.0, 12.0, 500))
        plane = np.zeros((900, 3)); plane[:, 0] = rng.uniform(-3, 3, 900); plane[:, 2] = rng.uniform(-3, 3, 900); plane[:, 1] = 2.0
        sets["plane_900"] = (plane, rng.uniform(0.02, 0.2, 900))
        return sets

    out = {"types": json.dumps(BVH_KAT_TYPES)}
    names = []
    with tempfile.TemporaryDirectory() as tmp:
        for name, (pos, rad) in blobs().items():
            j = dict(template)
            j.pop("bvh", None)
            j["surfaces"] = [dict(type="sphere", material=material, radius=float(r), position=[float(x) for x in p])
                             for p, r in zip(pos, rad)]
            with open(os.path.join(tmp, name + ".json"), "w") as f:
                json.dump(j, f)
            s = ref.RefScene(name + ".json", dict(width=8, height=8, sqrtspp=1), scenes=tmp)
            bounds, scene_bounds = s.prim_bounds()
            assert np.array_equal(bounds[:, :3], pos - rad[:, None]) and np.array_equal(bounds[:, 3:], pos + rad[:, None])
            out[f"{name}/prim_bounds"] = bounds; out[f"{name}/scene_bounds"] = scene_bounds
            names.append(name)
            for t, b in BVH_KAT_TYPES:
                if t == "octree" and name != "random_400":
                    continue   # > 8 coincident centroids: the reference's octree recursion does not terminate
                r = s.build_bvh(t, b)
                for k in ("node_bounds", "node_first_prim", "node_prim_count", "node_next_sibling", "prim_order"):
                    out[f"{name}/{t}:{b}/{k}"] = r[k]
                inner = r["node_prim_count"] == 0
                print(f"bvh_kat {name} {t}:{b}: {len(inner)} nodes, max leaf {r['node_prim_count'].max()}")
            s.close()
    out["sets"] = json.dumps(names)
    np.savez_compressed(os.path.join(HERE, "bvh_kat.npz"), **out)


def make_sampler_kat(rng):
    ref.set_seed(SEED)
    n = 8192
    pixel = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    pixel[:64] = np.arange(64)
    sample = rng.integers(0, 2 ** 16, n).astype(np.uint32)
    sample[:64] = np.arange(64)
    sample[64:96] = 0xFFFFFFFF - np.arange(32)
    out = {}
    for ns in (0, 1, 2, 5, 17, 81):
        out[f"stream_{ns}"] = ref.sampler_stream(pixel, sample, ns)
    np.savez_compressed(os.path.join(HERE, "sampler_kat.npz"), seed=np.uint32(SEED), pixel=pixel, sample=sample, **out)
    print("sampler_kat: ", {k: v.shape for k, v in out.items()})


def make_bsdf_kat(rng):
    import ctypes as C
    L = ref.lib()
    n = 2048
    n1 = rng.uniform(1.0, 2.5, n); n2 = rng.uniform(1.0, 3.5, n); c = rng.uniform(0.0, 1.0, n)
    fd = np.array([L.ref_fresnel_dielectric(a, b, x) for a, b, x in zip(n1, n2, c)])
    real = rng.uniform(0.1, 3.0, (n, 3)); imag = rng.uniform(0.5, 8.0, (n, 3))
    fc = np.zeros((n, 3))
    for i in range(n):
        L.ref_fresnel_conductor(n1[i], real[i].ctypes.data, imag[i].ctypes.data, c[i], fc[i].ctypes.data)

    def hemi(k, sign=1.0):
        v = rng.normal(size=(k, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
        v[:, 2] = sign * np.abs(v[:, 2]) + sign * 1e-3
        return v / np.linalg.norm(v, axis=1, keepdims=True)
    wo = hemi(n); wi_r = hemi(n); wi_t = hemi(n, -1.0)
    alpha = rng.uniform(0.02, 0.8, n)
    gr = np.zeros(n); gr_pdf = np.zeros(n); gt = np.zeros(n); gt_pdf = np.zeros(n); vm = np.zeros((n, 3))
    uv = rng.uniform(0, 1, (n, 2))
    pdf = C.c_double()
    for i in range(n):
        gr[i] = L.ref_ggx_reflection(wi_r[i].ctypes.data, wo[i].ctypes.data, alpha[i], C.byref(pdf)); gr_pdf[i] = pdf.value
        gt[i] = L.ref_ggx_transmission(wi_t[i].ctypes.data, wo[i].ctypes.data, n1[i], n2[i], alpha[i], C.byref(pdf)); gt_pdf[i] = pdf.value
        L.ref_ggx_visible_microfacet(uv[i, 0], uv[i, 1], wo[i].ctypes.data, alpha[i], vm[i].ctypes.data)
    np.savez_compressed(os.path.join(HERE, "bsdf_kat.npz"), n1=n1, n2=n2, cos=c, fresnel_dielectric=fd,
                        ior_real=real, ior_imag=imag, fresnel_conductor=fc, wo=wo, wi_r=wi_r, wi_t=wi_t,
                        alpha=alpha, ggx_refl=gr, ggx_refl_pdf=gr_pdf, ggx_trans=gt, ggx_trans_pdf=gt_pdf,
                        uv=uv, ggx_vndf=vm)
    print("bsdf_kat: n =", n)


def make_c2_band():
    """BASELINE config 2 at its benchmark size (hexagon_room 1920x1080, 256 spp, quaternary SAH): rows
    [538, 542) rendered by the reference (Camera::samplePixel over those buckets only) -> c2_band_kat.npz.
    The GPU renders the same rows of the same frame through mcrt_render_rows(cam, 538, 542) from
    bench_data/c2_hexagon_room.mcrtpack, the pack bench.py times."""
    ref.set_seed(SEED)
    ov = dict(width=1920, height=1080, sqrtspp=16, bvh_type="quaternary_sah")
    s = ref.RefScene("hexagon_room.json", ov)
    y0, y1 = 538, 542
    band, sec, rays, sh = s.render(threads=8, y0=y0, y1=y1)
    np.savez_compressed(os.path.join(HERE, "c2_band_kat.npz"), seed=np.uint32(SEED), y0=np.uint32(y0), y1=np.uint32(y1),
                        width=np.uint32(1920), height=np.uint32(1080), sqrtspp=np.uint32(16), band=band,
                        total_rays=np.uint64(rays), shadow_rays=np.uint64(sh))
    print(f"c2_band_kat: rows {y0}..{y1} rays={rays} (shadow {sh}) mean={band.mean():.6f} in {sec:.1f} s")
    s.close()


def remake_packs(only):
    """Re-exports the scene packs only (pack format grew; the reference outputs in the .npz stay)."""
    for cid, case in CASES.items():
        scene_file, overrides, pm = case[:3]
        if only and cid not in only:
            continue
        ref.set_seed(SEED)
        s = ref.RefScene(scene_file, overrides, photon_map=pm, scenes=case[3] if len(case) > 3 else None)
        s.export_pack(os.path.join(HERE, cid + ".mcrtpack"))
        s.close()
        print("pack", cid)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--packs-only":
        remake_packs(sys.argv[2:])
        sys.exit(0)
    only = sys.argv[1:]
    rng = np.random.default_rng(20260923)
    for cid, case in CASES.items():
        scene_file, overrides, pm = case[:3]
        if only and cid not in only:
            continue
        make_case(cid, scene_file, overrides, pm, rng, scenes=case[3] if len(case) > 3 else None)
    if not only or "c2_band_kat" in only:
        make_c2_band()
    if not only or "sampler_kat" in only:
        make_sampler_kat(rng)
    if not only or "film_kat" in only:
        make_film_kat()
    if not only or "image_kat" in only:
        make_image_kat()
    if not only or "obj_kat" in only:
        make_obj_kat()
    if not only or "bvh_kat" in only:
        make_bvh_kat()
