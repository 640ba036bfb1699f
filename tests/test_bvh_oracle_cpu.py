"""CPU tests: the scalar restatement of BVH::BVH in oracle/ (binned SAH binary / quaternary, octree-derived,
arbitrarySplit, compact) against trees the unmodified reference built - the golden scene packs and
tests/golden/bvh_kat.npz (synthetic sphere sets that reach the builders' rare branches). The GPU builder is
compared with the same fixtures in tests/test_gpu_parity.py / tools/probe_bvh.py."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases
from oracle import port

TYPE_CODE = {"octree": 0, "binary_sah": 1, "quaternary_sah": 2}


def _equal(got, ref, order):
    assert np.array_equal(got["node_first_prim"], ref["node_first_prim"])
    assert np.array_equal(got["node_prim_count"], ref["node_prim_count"])
    assert np.array_equal(got["node_next_sibling"], ref["node_next_sibling"])
    assert np.array_equal(got["node_bounds"].reshape(-1), np.asarray(ref["node_bounds"]).reshape(-1))
    assert np.array_equal(got["prim_order"], order)


@pytest.mark.parametrize("cid", [c for c in golden_cases() if c != "ior_test_nobvh_64"])
def test_restatement_on_scene_packs(cid, mcrt):
    scene = mcrt.Scene.from_pack(os.path.join(GOLDEN, cid + ".mcrtpack"))
    _, bvh_type, bins = (int(v) for v in scene.extra["bvh_params"])
    flat = scene.unbuilt()
    got = port.bvh_build(flat.prim_bounds(), scene.extra["scene_bounds"], bvh_type, bins, mcrt.BvhDesc)
    _equal(got, scene.a, scene.extra["prim_original"])


def bvh_kat_cases():
    k = np.load(os.path.join(GOLDEN, "bvh_kat.npz"))
    cases = []
    for name in json.loads(str(k["sets"])):
        for t, b in json.loads(str(k["types"])):
            if f"{name}/{t}:{b}/node_first_prim" in k.files:
                cases.append((name, t, b))
    return cases


@pytest.mark.parametrize("name,bvh_type,bins", bvh_kat_cases())
def test_restatement_on_degenerate_sets(name, bvh_type, bins, mcrt):
    k = np.load(os.path.join(GOLDEN, "bvh_kat.npz"))
    key = f"{name}/{bvh_type}:{bins}"
    ref = {f: k[f"{key}/{f}"] for f in ("node_bounds", "node_first_prim", "node_prim_count", "node_next_sibling")}
    got = port.bvh_build(k[f"{name}/prim_bounds"], k[f"{name}/scene_bounds"], TYPE_CODE[bvh_type], bins, mcrt.BvhDesc)
    _equal(got, ref, k[f"{key}/prim_order"])
