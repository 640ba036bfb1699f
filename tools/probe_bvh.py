#!/usr/bin/env python3
"""GPU BVH builder vs the reference's trees (bench_data/bvh_<scene>.npz from tools/make_bvh_cases.py):
array-for-array comparison with diagnostics, and build time next to the reference's CPU time."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mcrt = importlib.import_module("monte-carlo-ray-tracer_b200")


def compare(got, ref, label):
    ok = True
    if len(got["node_first_prim"]) != len(ref["node_first_prim"]):
        print(f"  {label}: node count {len(got['node_first_prim'])} != {len(ref['node_first_prim'])}")
        ok = False
    m = min(len(got["node_first_prim"]), len(ref["node_first_prim"]))
    for k in ("node_first_prim", "node_prim_count", "node_next_sibling"):
        bad = np.nonzero(got[k][:m] != ref[k][:m])[0]
        if len(bad):
            ok = False
            print(f"  {label}: {k} differs at {len(bad)} nodes, first {bad[0]}: {got[k][bad[0]]} vs {ref[k][bad[0]]}")
    gb, rb = got["node_bounds"].reshape(-1, 6)[:m], ref["node_bounds"].reshape(-1, 6)[:m]
    bad = np.nonzero((gb != rb).any(axis=1))[0]
    if len(bad):
        ok = False
        print(f"  {label}: node_bounds differ at {len(bad)} nodes, first {bad[0]}: {gb[bad[0]]} vs {rb[bad[0]]}")
    bad = np.nonzero(got["prim_order"] != ref["prim_order"])[0]
    if len(bad):
        ok = False
        print(f"  {label}: prim_order differs at {len(bad)} positions, first {bad[0]}")
    return ok


def probe_kat(results):
    """tests/golden/bvh_kat.npz: sphere sets that reach arbitrarySplit / the quaternary->binary fall-back."""
    path = os.path.join(ROOT, "tests", "golden", "bvh_kat.npz")
    z = np.load(path)
    for name in json.loads(str(z["sets"])):
        for t, b in json.loads(str(z["types"])):
            key = f"{name}/{t}:{b}"
            if f"{key}/node_first_prim" not in z.files:
                continue
            ref = {k: z[f"{key}/{k}"] for k in ("node_bounds", "node_first_prim", "node_prim_count", "node_next_sibling", "prim_order")}
            got = mcrt.bvh_build(z[f"{name}/prim_bounds"], z[f"{name}/scene_bounds"], t, int(b))
            ok = compare(got, ref, key)
            print(f"kat {key}: {'IDENTICAL' if ok else 'DIFFERENT'} nodes {len(got['node_first_prim'])} rounds {got['rounds']}")
            results["kat " + key] = dict(identical=bool(ok), nodes=int(len(got["node_first_prim"])), rounds=got["rounds"])


if __name__ == "__main__":
    results = {}
    if "kat" in sys.argv[1:]:
        sys.argv.remove("kat")
        probe_kat(results)
    for name in (sys.argv[1:] or ["spaceship", "lego_bulldozer"]):
        path = os.path.join(ROOT, "bench_data", f"bvh_{name}.npz")
        if not os.path.exists(path):
            print("missing", path)
            continue
        z = np.load(path)
        bounds, sb = z["prim_bounds"], z["scene_bounds"]
        for case in [str(c) for c in z["cases"]]:
            t, b = case.split(":")
            ref = {k: z[f"{case}/{k}"] for k in ("node_bounds", "node_first_prim", "node_prim_count", "node_next_sibling", "prim_order")}
            best = None
            for rep in range(3):
                got = mcrt.bvh_build(bounds, sb, t, int(b))
                best = got["gpu_ms"] if best is None else min(best, got["gpu_ms"])
            ok = compare(got, ref, f"{name} {case}")
            cpu = float(z[f"{case}/cpu_seconds"])
            print(f"{name} {case}: {'IDENTICAL' if ok else 'DIFFERENT'}  prims {len(bounds)} nodes {len(got['node_first_prim'])} "
                  f"rounds {got['rounds']} launches {got['kernel_launches']}  GPU {best:.2f} ms  reference CPU {cpu * 1e3:.0f} ms  "
                  f"({cpu * 1e3 / best:.0f}x)")
            results[f"{name} {case}"] = dict(identical=bool(ok), prims=int(len(bounds)), nodes=int(len(got["node_first_prim"])),
                                             rounds=got["rounds"], gpu_ms=best, reference_cpu_ms=cpu * 1e3)
    out = os.path.join(ROOT, "gpurun_out", "bvh_build.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(results, open(out, "w"), indent=1)
