#!/usr/bin/env python3
"""Builds the reference-side host adapter (needs /root/reference; container only):
   host/_build/mcrt_gpu_render = reference objects (as compiled by oracle/build_ref.py)
                               + exporter.cpp + gpu_integrator.cpp + mcrt_gpu_render.cpp
                               + libmcrt_b200.so"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)


def build(reference="/root/reference"):
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    ref_objs = sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "obj", "ref_*.o")))
    if not ref_objs:
        raise SystemExit("run oracle/build_ref.py first (it compiles the reference's translation units)")
    inc = ["-I", os.path.join(reference, "lib", "glm"), "-I", os.path.join(reference, "lib", "nlohmann"),
           "-I", os.path.join(reference, "source"), "-I", os.path.join(ROOT, "include"), "-I", HERE]
    objs = []
    for src in ("exporter.cpp", "gpu_integrator.cpp", "mcrt_gpu_render.cpp"):
        o = os.path.join(out, src[:-4] + ".o")
        cmd = ["g++", "-std=c++20", "-O2", "-w", "-fno-access-control", "-include", os.path.join(ROOT, "oracle", "seed_pin.hpp")] + inc + \
              ["-c", os.path.join(HERE, src), "-o", o]
        subprocess.check_call(cmd)
        objs.append(o)
    # pinnedSeed() lives in the oracle driver; the product binary gets its own tiny definition
    seed_src = os.path.join(out, "seed.cpp")
    with open(seed_src, "w") as f:
        f.write('#include <cstdlib>\nnamespace mcrt_oracle { unsigned pinnedSeed() { const char* e = std::getenv("MCRT_SEED"); '
                'return e ? (unsigned)std::strtoul(e, nullptr, 0) : 0x12345678u; } }\n')
    exe = os.path.join(out, "mcrt_gpu_render")
    subprocess.check_call(["g++", "-pthread", "-o", exe, seed_src] + objs + ref_objs +
                          ["-L", PKG, "-lmcrt_b200", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN/../.."])
    return exe


if __name__ == "__main__":
    print(build(*(sys.argv[1:2])))
