#!/usr/bin/env python3
"""ORACLE / TEST INFRASTRUCTURE — builds oracle/_ref/libmcrt_ref.so from the UNMODIFIED reference.

The reference's own translation units are compiled where they lie under /root/reference/source
(nothing is copied into this repository) with the flags of its CMakeLists.txt:16-21,30,40-41
(C++20, -O3, no -march, no fast-math, pthreads) plus what a shared object needs (-fPIC,
-fno-semantic-interposition so that same-TU inlining matches the executable build). Additions:
  * -include oracle/seed_pin.hpp      pins Sampler::global_seed (source/sampling/sampler.hpp:58)
  * -Wl,--wrap=<Scene::intersect>,<Integrator::sampleDirect>   ray counters in ref_driver.cpp
  * oracle/ref_driver.cpp + monte-carlo-ray-tracer_b200/host/exporter.cpp with -fno-access-control
The reference's CMake build is NOT run. Outputs go to oracle/_ref/ only (git-ignored, but shipped
to the GPU box by gpurun): libmcrt_ref.so and a copy of the scene JSONs / small data files the
CPU baseline needs at run time there.

Usage: python oracle/build_ref.py [--reference /root/reference] [--force] [--with-data spaceship,...]
"""
import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "oracle", "_ref")

WRAP = [
    "_ZNK5Scene9intersectERK3Ray",
    "_ZNK10Integrator12sampleDirectERK11InteractionRNS_11LightSampleE",
]


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr + "\n")
        raise SystemExit("oracle/_ref build failed")
    return r


def newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in (src, *extra))


def build(reference="/root/reference", force=False, with_data=()):
    src_root = os.path.join(reference, "source")
    if not os.path.isdir(src_root):
        raise SystemExit(f"reference sources not found at {src_root}")
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)

    seed_pin = os.path.join(ROOT, "oracle", "seed_pin.hpp")
    inc = [
        "-I", os.path.join(reference, "lib", "glm"),
        "-I", os.path.join(reference, "lib", "nlohmann"),
        "-I", src_root,
        "-I", os.path.join(ROOT, "include"),
        "-I", os.path.join(ROOT, "monte-carlo-ray-tracer_b200", "host"),
    ]
    base = ["g++", "-std=c++20", "-O3", "-fPIC", "-fno-semantic-interposition", "-pthread", "-w",
            "-include", seed_pin]

    ref_units = []
    for d, _, files in os.walk(src_root):
        for f in sorted(files):
            # main.cpp is the interactive CLI; octree*.cpp are template bodies that the reference
            # itself #includes from bvh.cpp / photon-mapper.cpp
            if f.endswith(".cpp") and f not in ("main.cpp",):
                ref_units.append(os.path.join(d, f))

    own_units = [
        os.path.join(ROOT, "oracle", "ref_driver.cpp"),
        os.path.join(ROOT, "monte-carlo-ray-tracer_b200", "host", "exporter.cpp"),
        os.path.join(ROOT, "monte-carlo-ray-tracer_b200", "host", "obj_loader.cpp"),
    ]
    own_deps = [seed_pin,
                os.path.join(ROOT, "monte-carlo-ray-tracer_b200", "host", "obj_loader.hpp"),
                os.path.join(ROOT, "monte-carlo-ray-tracer_b200", "host", "obj_adapter.hpp"),
                os.path.join(ROOT, "include", "mcrt_abi.h"),
                os.path.join(ROOT, "monte-carlo-ray-tracer_b200", "host", "exporter.hpp")]

    jobs = []
    objs = []
    for u in ref_units:
        o = os.path.join(OUT, "obj", "ref_" + os.path.relpath(u, src_root).replace("/", "_")[:-4] + ".o")
        objs.append(o)
        if force or newer(u, o, (seed_pin,)):
            jobs.append(base + inc + ["-c", u, "-o", o])
    for u in own_units:
        o = os.path.join(OUT, "obj", "own_" + os.path.basename(u)[:-4] + ".o")
        objs.append(o)
        if force or newer(u, o, own_deps):
            jobs.append(base + ["-fno-access-control"] + inc + ["-c", u, "-o", o])

    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(run, jobs))

    lib = os.path.join(OUT, "libmcrt_ref.so")
    if force or jobs or not os.path.exists(lib):
        link = ["g++", "-shared", "-pthread", "-o", lib] + objs + \
               ["-Wl,-Bsymbolic-functions"] + [f"-Wl,--wrap={s}" for s in WRAP]
        run(link)

    # run-time inputs of the CPU baseline on the GPU box: scene JSONs + small data files
    scenes_src = os.path.join(reference, "scenes")
    scenes_dst = os.path.join(OUT, "scenes")
    os.makedirs(os.path.join(scenes_dst, "data"), exist_ok=True)
    for f in sorted(os.listdir(scenes_src)):
        if f.endswith(".json"):
            shutil.copyfile(os.path.join(scenes_src, f), os.path.join(scenes_dst, f))
    small = ["spectral-distributions", "backwall.obj", "shelf.obj"] + list(with_data)
    for name in small:
        s = os.path.join(scenes_src, "data", name)
        d = os.path.join(scenes_dst, "data", name)
        if os.path.isdir(s):
            if not os.path.isdir(d):
                shutil.copytree(s, d)
        elif os.path.exists(s) and not os.path.exists(d):
            shutil.copyfile(s, d)
    return lib


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--with-data", default="")
    a = ap.parse_args()
    print(build(a.reference, a.force, tuple(x for x in a.with_data.split(",") if x)))
