"""CPU test of the octree restatement in oracle/ (Octree<Photon> + LinearOctree::compact): fed the
photons of the reference's own maps (in shuffled order), it must reproduce the reference's
LinearOctree exactly - same octants in the same depth-first order (bounds, ranges, sibling links,
leaf flags) and the same photons in every leaf (as a set: the reference's order inside a leaf
depends on its thread schedule). The GPU builder is then compared with this restatement
array-for-array (tests/test_gpu_parity.py)."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import port


def oracle_octree(mcrt, photons, leaf, bounds):
    return port.build_photon_octree(photons, leaf, bounds, mcrt.PhotonMapDesc, mcrt._map_arrays)


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
        built = oracle_octree(mcrt, shuffled, int(params[2]), params[3:9])
        for key in ("octant_start", "octant_count", "octant_next", "octant_leaf", "octant_bounds"):
            assert np.array_equal(built[key], ref_map[key]), key
        assert _leaf_sets(built) == _leaf_sets(ref_map)


def test_octree_edge_cases(mcrt):
    bounds = np.array([0, 0, 0, 1, 1, 1], dtype=np.float64)
    empty = oracle_octree(mcrt, np.zeros((0, 8), np.float32), 4, bounds)
    assert empty["octant_leaf"].size == 0 and empty["photons"].size == 0
    one = np.zeros((1, 8), np.float32); one[0, 3:6] = 0.25
    m = oracle_octree(mcrt, one, 4, bounds)
    assert m["octant_leaf"].tolist() == [1] and m["octant_count"].tolist() == [1] and m["octant_next"].tolist() == [0xFFFFFFFF]
    # more coincident photons than a leaf may hold: the reference would recurse forever; we stop
    same = np.zeros((9, 8), np.float32); same[:, 3:6] = 0.3
    m = oracle_octree(mcrt, same, 4, bounds)
    assert int(m["octant_count"][0]) == 9 and m["octant_leaf"][-1] == 1
