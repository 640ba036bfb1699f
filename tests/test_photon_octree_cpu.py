"""CPU test of the host-side octree construction used by mcrt_photon_emit: fed the photons of the
reference's own maps (in shuffled order), it must reproduce the reference's LinearOctree exactly -
same octants in the same depth-first order (bounds, ranges, sibling links, leaf flags) and the same
photons in every leaf (as a set: the reference's order inside a leaf depends on its thread schedule)."""
import os

import numpy as np

from conftest import GOLDEN


def _leaf_sets(m):
    ph = m["photons"].reshape(-1, 8)
    sets = []
    for s, c, leaf in zip(m["octant_start"], m["octant_count"], m["octant_leaf"]):
        if leaf:
            block = ph[int(s):int(s) + int(c)]
            sets.append(block[np.lexsort(block.T[::-1])].tobytes())
    return sets


def test_octree_matches_reference(mcrt):
    scene = mcrt.Scene.from_pack(os.path.join(GOLDEN, "pm_hexagon_room_64.mcrtpack"))
    caustic, glob, _, _ = scene.photon_maps()
    params = scene.extra["photon_emit_params"]
    rng = np.random.default_rng(5)
    for ref_map in (caustic, glob):
        photons = ref_map["photons"].reshape(-1, 8)
        shuffled = photons[rng.permutation(len(photons))]
        built = mcrt.build_photon_octree(shuffled, int(params[2]), params[3:9])
        for key in ("octant_start", "octant_count", "octant_next", "octant_leaf", "octant_bounds"):
            assert np.array_equal(built[key], ref_map[key]), key
        assert _leaf_sets(built) == _leaf_sets(ref_map)


def test_parallel_build_path_matches_reference():
    """Same check through the multi-threaded front end (subtrees built by worker threads and
    concatenated): forced on for this small input with MCRT_OCTREE_PAR_MIN in a fresh process."""
    import subprocess, sys
    code = (
        "import importlib, os, sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "m = importlib.import_module('monte-carlo-ray-tracer_b200')\n"
        "scene = m.Scene.from_pack(%r)\n"
        "maps = scene.photon_maps(); p = scene.extra['photon_emit_params']\n"
        "for ref in maps[:2]:\n"
        "    ph = ref['photons'].reshape(-1, 8)[::-1]\n"
        "    b = m.build_photon_octree(ph, int(p[2]), p[3:9])\n"
        "    for k in ('octant_start', 'octant_count', 'octant_next', 'octant_leaf', 'octant_bounds'):\n"
        "        assert np.array_equal(b[k], ref[k]), k\n"
        "print('ok')\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(GOLDEN, "pm_hexagon_room_64.mcrtpack"))
    env = dict(os.environ, MCRT_OCTREE_PAR_MIN="300")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_octree_edge_cases(mcrt):
    bounds = np.array([0, 0, 0, 1, 1, 1], dtype=np.float64)
    empty = mcrt.build_photon_octree(np.zeros((0, 8), np.float32), 4, bounds)
    assert empty["octant_leaf"].size == 0 and empty["photons"].size == 0
    one = np.zeros((1, 8), np.float32); one[0, 3:6] = 0.25
    m = mcrt.build_photon_octree(one, 4, bounds)
    assert m["octant_leaf"].tolist() == [1] and m["octant_count"].tolist() == [1] and m["octant_next"].tolist() == [0xFFFFFFFF]
    # more coincident photons than a leaf may hold: the reference would recurse forever; we stop
    same = np.zeros((9, 8), np.float32); same[:, 3:6] = 0.3
    m = mcrt.build_photon_octree(same, 4, bounds)
    assert int(m["octant_count"][0]) == 9 and m["octant_leaf"][-1] == 1
