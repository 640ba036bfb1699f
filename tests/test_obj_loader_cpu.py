"""CPU tests of the scene-ingest helpers (host code in the product library): the
parallel OBJ reader and the vertex-normal generator against outputs of the reference's
Scene::parseOBJ / Scene::generateVertexNormals (tests/golden/obj_kat.npz, made by make_golden.py
from tests/golden/quirks.obj)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


def test_obj_reader_matches_reference(mcrt):
    k = np.load(os.path.join(GOLDEN, "obj_kat.npz"))
    for threads in (1, 3):
        g = mcrt.load_obj(os.path.join(GOLDEN, "quirks.obj"), threads=threads)
        for key in ("vertices", "normals", "tri_v", "tri_vt", "tri_vn"):
            assert np.array_equal(g[key], k[key]), key
    assert int(k["negative_throws"]) == 1
    with pytest.raises(mcrt.McrtError, match="negative offsets"):
        mcrt.load_obj(os.path.join(GOLDEN, "quirks_negative.obj"))
    with pytest.raises(mcrt.McrtError, match="not found"):
        mcrt.load_obj(os.path.join(GOLDEN, "no_such_file.obj"))


def test_obj_reader_chunking_is_order_preserving(mcrt, tmp_path):
    # a file large enough to be split over threads parses to the same arrays as with one thread
    rng = np.random.default_rng(3)
    n = 20000
    v = rng.normal(size=(n, 3))
    f = rng.integers(1, n + 1, (3 * n, 3))
    path = tmp_path / "big.obj"
    with open(path, "w") as out:
        for i in range(n):
            out.write("v " + " ".join(repr(float(x)) for x in v[i]) + "\n")
            for j in range(3):
                a, b, c = f[3 * i + j]
                out.write(f"f {a}//{b} {b}//{c} {c}//{a}\n" if j else f"f {a} {b} {c}\n")
    one = mcrt.load_obj(str(path), threads=1)
    many = mcrt.load_obj(str(path), threads=7)
    assert np.array_equal(one["vertices"], v)            # repr() round-trips through a correctly rounded parse
    assert np.array_equal(one["tri_v"], f.astype(np.uint64) - 1)
    for key in one:
        assert np.array_equal(one[key], many[key]), key


def test_vertex_normals_match_reference(mcrt):
    k = np.load(os.path.join(GOLDEN, "obj_kat.npz"))
    valid = (k["tri_v"] < len(k["vertices"])).all(axis=1)
    for threads in (1, 4):
        n = mcrt.vertex_normals(k["vertices"], k["tri_v"][valid], threads=threads)
        assert np.array_equal(n, k["generated_normals"], equal_nan=True)
    with pytest.raises(mcrt.McrtError):
        mcrt.vertex_normals(k["vertices"], [[0, 1, len(k["vertices"])]])    # Scene::generateVertexNormals' .at() throws
