"""CPU tests (no GPU): the scalar float64 restatement (oracle/mcrt_oracle.cpp) against golden outputs
of the unmodified reference (tests/golden/*.npz). This pins the restatement; the GPU parity tests
use the same fixtures directly."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases
from oracle import port

PT_CASES = [c for c in golden_cases() if not c.startswith("pm_")]


@pytest.fixture(scope="module")
def scenes(mcrt):
    cache = {}

    def get(cid):
        if cid not in cache:
            scene = mcrt.Scene.from_pack(os.path.join(GOLDEN, cid + ".mcrtpack"))
            cache[cid] = (port.PortScene(scene), scene, np.load(os.path.join(GOLDEN, cid + ".npz")))
        return cache[cid]
    yield get
    for ps, _, _ in cache.values():
        ps.close()


def test_sampler_streams_bit_exact():
    k = np.load(os.path.join(GOLDEN, "sampler_kat.npz"))
    for name in k.files:
        if name.startswith("stream_"):
            got = port.sampler_stream(k["pixel"], k["sample"], int(name.split("_")[1]), int(k["seed"]))
            assert np.array_equal(got, k[name]), name


@pytest.mark.parametrize("cid", PT_CASES)
def test_intersect(cid, mcrt, scenes):
    ps, _, g = scenes(cid)
    hits = ps.trace(g["tr_rays"])
    assert np.array_equal(hits["prim"], g["tr_prim"])
    hit = g["tr_prim"] != mcrt.NO_PRIM
    assert np.array_equal(hits["t"][hit], g["tr_t"][hit])          # same arithmetic, same compiler: bit-exact
    interp = g["tr_interp"].astype(bool) & hit
    assert np.array_equal(hits["u"][interp], g["tr_uv"][interp, 0])
    assert np.array_equal(hits["interpolate"].astype(bool), g["tr_interp"].astype(bool) & hit)


@pytest.mark.parametrize("cid", PT_CASES)
def test_sample_ray(cid, scenes):
    ps, _, g = scenes(cid)
    rgb = ps.sample_rays(g["ps_rays"], g["ps_pixel"], g["ps_sample"], int(g["seed"]))
    err = np.abs(rgb - g["ps_rgb"]) / np.maximum(1.0, np.abs(g["ps_rgb"]))
    assert err.max() <= 1e-12, f"worst {err.max():.3e}"


@pytest.mark.parametrize("cid", [c for c in PT_CASES if c != "c1_hexagon_diffuse_256"])
def test_image(cid, scenes):
    ps, scene, g = scenes(cid)
    cam = scene.cameras()[0]
    img, rays = ps.render_rows(cam, 0, cam.height, cam.sqrtspp, int(g["seed"]))
    assert rays == int(g["total_rays"])
    assert np.abs(img - g["image"]).max() <= 1e-12 * max(1.0, np.abs(g["image"]).max())


def test_image_config1_rows(scenes):
    # BASELINE config 1 (256x256, 4 spp): a band of rows keeps the CPU suite fast
    ps, scene, g = scenes("c1_hexagon_diffuse_256")
    cam = scene.cameras()[0]
    img, _ = ps.render_rows(cam, 96, 160, cam.sqrtspp, int(g["seed"]))
    assert np.abs(img - g["image"][96:160]).max() <= 1e-12 * max(1.0, np.abs(g["image"]).max())


def _film_cases():
    import json
    k = np.load(os.path.join(GOLDEN, "film_kat.npz"))
    return k, json.loads(str(k["films"]))


@pytest.mark.parametrize("name", ["mitchell", "catmull_rom", "b_spline", "hermite", "gaussian_cached", "lanczos_r3",
                                  "lanczos_cached", "box_r1p5", "box_default_cached"])
def test_film_filters(name, mcrt):
    # Film::deposit / Film::scan with the reference's reconstruction filters (film.cpp:61-113)
    k, films = _film_cases()
    scene = mcrt.Scene.from_pack(os.path.join(GOLDEN, "film_hexagon_room_64.mcrtpack"))
    cam = scene.cameras()[0]
    assert cam.film is not None and cam.film["filter"] == mcrt.FILM_FILTERS["mitchell-netravali"]   # from the pack
    cam.film = films[name]
    ps = port.PortScene(scene)
    try:
        img = ps.render_film(cam, cam.sqrtspp, int(k["seed"]))
    finally:
        ps.close()
    ref = k["image_" + name]
    assert np.abs(img - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


def _image_cases():
    import json
    k = np.load(os.path.join(GOLDEN, "image_kat.npz"))
    cases = []
    for iname in json.loads(str(k["inputs"])):
        for cname in json.loads(str(k["configs"])):
            cases.append((iname, cname))
    return cases


def image_kat_input(k, iname):
    if iname.startswith("golden:"):
        return np.load(os.path.join(GOLDEN, iname[7:] + ".npz"))["image"]
    return k["input/" + iname]


@pytest.mark.parametrize("iname,cname", _image_cases())
def test_image_pipeline(iname, cname, mcrt):
    # Image::save: auto exposure / tone mapping / auto gain / gamma / bytes (image.cpp:37-88)
    import json
    k = np.load(os.path.join(GOLDEN, "image_kat.npz"))
    params = mcrt.ImageParams.from_json(json.loads(str(k["configs"]))[cname])
    got, e, g = port.image_tonemap(image_kat_input(k, iname), params)
    assert np.array_equal(got, k[f"bytes/{iname}/{cname}"])
    assert (e, g) == tuple(k[f"factors/{iname}/{cname}"])
