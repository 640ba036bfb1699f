"""CPU test of host logic: the 4-wide float-box BVH that mcrt_scene_upload derives from the reference's tree for the order-free
closest-hit search (csrc/bvh4.cuh, built by buildBvh4 in csrc/abi.cu; mcrt_bvh4_host exposes it without a CUDA call).
Whatever its shape, the search is only correct if every primitive is reachable exactly once and every box contains what is below it."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases

LEAF = 0x80000000


def prim_boxes(scene):
    a = scene.a
    n = scene.n_prims
    lo, hi = np.zeros((n, 3)), np.zeros((n, 3))
    for i in range(n):
        t, idx = int(a["prim_type"][i]), int(a["prim_index"][i])
        if t == 0:
            v = np.stack([a["tri_v0"].reshape(-1, 3)[idx], a["tri_v1"].reshape(-1, 3)[idx], a["tri_v2"].reshape(-1, 3)[idx]])
            lo[i], hi[i] = v.min(axis=0), v.max(axis=0)
        elif t == 1:
            s = a["sphere_origin_radius"].reshape(-1, 4)[idx]
            lo[i], hi[i] = s[:3] - s[3], s[:3] + s[3]
        else:
            b = a["quadric_bounds"].reshape(-1, 6)[idx]
            lo[i], hi[i] = b[:3], b[3:]
    return lo, hi


@pytest.mark.parametrize("max_leaf", [0, 2, 0xFFFFFFFF])
@pytest.mark.parametrize("cid", [c for c in golden_cases() if c != "ior_test_nobvh_64"])
def test_bvh4_covers_every_primitive_once_with_containing_boxes(cid, max_leaf, mcrt):
    scene = mcrt.Scene.from_pack(os.path.join(GOLDEN, cid + ".mcrtpack"))
    nodes = mcrt.bvh4_host(scene, max_leaf)
    assert len(nodes) > 0 and nodes.dtype.itemsize == 128
    plo, phi = prim_boxes(scene)
    seen = np.zeros(scene.n_prims, dtype=int)
    visited = np.zeros(len(nodes), dtype=int)

    def walk(ni):
        """-> (lo, hi) float64 bounds of everything below node ni"""
        visited[ni] += 1
        node = nodes[ni]
        lo_all, hi_all = np.full(3, np.inf), np.full(3, -np.inf)
        assert (node["child"] != 0).any()
        for c in range(4):
            ref = int(node["child"][c])
            if ref == 0:
                assert (node["lo"][:, c] > node["hi"][:, c]).all()          # empty slot: inverted box, never hit
                continue
            if ref & LEAF:
                first, count = (ref >> 8) & 0x7FFFFF, ref & 0xFF
                assert count >= 1
                seen[first:first + count] += 1
                lo, hi = plo[first:first + count].min(axis=0), phi[first:first + count].max(axis=0)
            else:
                assert 0 < ref < len(nodes)
                lo, hi = walk(ref)
            # float boxes rounded outwards: never inside the float64 extent of what they hold
            assert (node["lo"][:, c].astype(np.float64) <= lo).all() and (node["hi"][:, c].astype(np.float64) >= hi).all(), (cid, ni, c)
            lo_all, hi_all = np.minimum(lo_all, lo), np.maximum(hi_all, hi)
        return lo_all, hi_all

    walk(0)
    assert (seen == 1).all()                    # every ordered primitive in exactly one leaf
    assert (visited == 1).all()                 # a tree: every node reached once from the root
    if max_leaf == 2:
        counts = [int(r) & 0xFF for r in nodes["child"].reshape(-1) if int(r) & LEAF]
        assert max(counts) <= 2


def test_bvh4_absent_without_a_tree(mcrt):
    scene = mcrt.Scene.from_pack(os.path.join(GOLDEN, "ior_test_nobvh_64.mcrtpack"))
    assert scene.n_nodes == 0 and len(mcrt.bvh4_host(scene)) == 0     # linear-scan scenes use the replay traversal
