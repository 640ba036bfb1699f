#!/usr/bin/env python3
"""Builds the reference-side host adapters (needs the reference's sources; build container only):

   host/_build/mcrt_gpu_render        the reference's main with Camera::sampleImage replaced (mcrt_gpu_render.cpp)
   host/_build/test_gpu_integrators   GpuPathTracer / GpuPhotonMapper substituted at camera.cpp:22-29 (test_gpu_integrators.cpp)

The reference's own translation units are compiled HERE, where they lie, with the flags of its
CMakeLists.txt:16-21,30 (C++20, -O3, pthreads) into host/_build/obj/ - nothing is taken from oracle/
(the test infrastructure's build directory) and Sampler::global_seed is the reference's own
std::random_device value (pin it per run with the MCRT_SEED environment variable of mcrt_gpu_render)."""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr + "\n")
        raise SystemExit("host adapter build failed")


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in (src, *extra))


def build(reference="/root/reference", force=False):
    out = os.path.join(HERE, "_build")
    objdir = os.path.join(out, "obj")
    os.makedirs(objdir, exist_ok=True)
    src_root = os.path.join(reference, "source")
    inc = ["-I", os.path.join(reference, "lib", "glm"), "-I", os.path.join(reference, "lib", "nlohmann"),
           "-I", src_root, "-I", os.path.join(ROOT, "include"), "-I", HERE]
    jobs, ref_objs = [], []
    for d, _, files in os.walk(src_root):
        for f in sorted(files):
            if f.endswith(".cpp") and f != "main.cpp":
                u = os.path.join(d, f)
                o = os.path.join(objdir, "ref_" + os.path.relpath(u, src_root).replace("/", "_")[:-4] + ".o")
                ref_objs.append(o)
                if force or _newer(u, o):
                    jobs.append(["g++", "-std=c++20", "-O3", "-pthread", "-w"] + inc + ["-c", u, "-o", o])
    own = ["exporter.cpp", "gpu_integrator.cpp", "gpu_integrators.cpp", "obj_loader.cpp", "mcrt_gpu_render.cpp", "test_gpu_integrators.cpp"]
    headers = [os.path.join(HERE, h) for h in os.listdir(HERE) if h.endswith(".hpp")] + [os.path.join(ROOT, "include", "mcrt_abi.h")]
    own_objs = {}
    for src in own:
        o = os.path.join(objdir, "own_" + src[:-4] + ".o")
        own_objs[src] = o
        if force or _newer(os.path.join(HERE, src), o, headers):
            jobs.append(["g++", "-std=c++20", "-O2", "-pthread", "-w", "-fno-access-control"] + inc + ["-c", os.path.join(HERE, src), "-o", o])
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(_run, jobs))
    common = [own_objs[s] for s in ("exporter.cpp", "gpu_integrator.cpp", "gpu_integrators.cpp", "obj_loader.cpp")] + ref_objs
    link = ["-L", PKG, "-lmcrt_b200", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN/../.."]
    exes = []
    for main_src, name in (("mcrt_gpu_render.cpp", "mcrt_gpu_render"), ("test_gpu_integrators.cpp", "test_gpu_integrators")):
        exe = os.path.join(out, name)
        _run(["g++", "-pthread", "-o", exe, own_objs[main_src]] + common + link)
        exes.append(exe)
    return exes[0]


if __name__ == "__main__":
    print(build(*(sys.argv[1:2])))
