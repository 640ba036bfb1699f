"""ORACLE / TEST INFRASTRUCTURE - ctypes binding of the CPU restatement (oracle/mcrt_oracle.cpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module."""
import ctypes as C
import importlib
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libmcrt_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            import importlib.util
            spec = importlib.util.spec_from_file_location("build_oracle", os.path.join(HERE, "build_oracle.py"))
            mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
            mod.build()
        L = C.CDLL(LIB_PATH)
        L.oracle_scene_create.restype = C.c_void_p
        L.oracle_scene_create.argtypes = [C.c_void_p]
        L.oracle_scene_destroy.argtypes = [C.c_void_p]
        L.oracle_sampler_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p]
        L.oracle_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_sample_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
        L.oracle_render_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint64)]
        L.oracle_sample_pixels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p]
        L.oracle_trace_fast.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.oracle_trace_visible.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.oracle_render_rows_checked.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                 C.c_void_p, C.c_void_p]
        L.oracle_bvh_build.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p]
        L.oracle_bvh_free.argtypes = [C.c_void_p]
        L.oracle_octree_build.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]
        L.oracle_octree_free.argtypes = [C.c_void_p]
        L.oracle_image_tonemap.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_render_film.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def sampler_stream(pixel, sample, n_shuffles, seed):
    pixel = np.ascontiguousarray(pixel, dtype=np.uint32); sample = np.ascontiguousarray(sample, dtype=np.uint32)
    out = np.zeros((len(pixel), 7), dtype=np.uint32)
    lib().oracle_sampler_stream(_p(pixel), _p(sample), len(pixel), n_shuffles, seed, _p(out))
    return out


def bvh_build(prim_bounds, scene_bounds, bvh_type, bins_per_axis, desc_type):
    """BVH::BVH restated (scalar, host). bvh_type: 0 octree, 1 binary_sah, 2 quaternary_sah; desc_type: the
    product's BvhDesc ctypes struct (layout only). -> dict of node arrays + prim_order."""
    prim_bounds = np.ascontiguousarray(prim_bounds, dtype=np.float64).reshape(-1, 6)
    scene_bounds = np.ascontiguousarray(scene_bounds, dtype=np.float64).reshape(6)
    h, d = C.c_void_p(), desc_type()
    lib().oracle_bvh_build(_p(prim_bounds), len(prim_bounds), _p(scene_bounds), int(bvh_type), int(bins_per_axis), C.byref(h), C.addressof(d))

    def arr(ptr, count, dtype):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count * np.dtype(dtype).itemsize,)).view(dtype).copy() if count else np.zeros(0, dtype)
    out = dict(node_bounds=arr(d.node_bounds, d.n_nodes * 6, np.float64).reshape(-1, 6), node_first_prim=arr(d.node_first_prim, d.n_nodes, np.uint32),
               node_prim_count=arr(d.node_prim_count, d.n_nodes, np.uint32), node_next_sibling=arr(d.node_next_sibling, d.n_nodes, np.uint32),
               prim_order=arr(d.prim_order, d.n_prims, np.uint32))
    lib().oracle_bvh_free(h)
    return out


def build_photon_octree(photons, max_photons_per_octree_leaf, scene_bounds, desc_type, map_arrays):
    """Octree<Photon> + LinearOctree::compact restated (scalar, host). desc_type / map_arrays: the
    product's PhotonMapDesc struct and its numpy unpacker (layout only)."""
    photons = np.ascontiguousarray(photons, dtype=np.float32).reshape(-1, 8)
    bounds = np.ascontiguousarray(scene_bounds, dtype=np.float64)
    h, d = C.c_void_p(), desc_type()
    lib().oracle_octree_build(_p(photons), len(photons), int(max_photons_per_octree_leaf), _p(bounds), C.byref(h), C.addressof(d))
    out = map_arrays(d)
    lib().oracle_octree_free(h)
    return out


def image_tonemap(rgb, params):
    """Image::save restated; params = the product's ImageParams ctypes struct (layout only)."""
    rgb = np.ascontiguousarray(rgb, dtype=np.float64)
    h, w = rgb.shape[:2]
    out = np.zeros((h, w, 3), dtype=np.uint8)
    e, g = C.c_double(), C.c_double()
    lib().oracle_image_tonemap(_p(rgb), w, h, C.addressof(params), _p(out), C.byref(e), C.byref(g))
    return out, e.value, g.value


class PortScene:
    """The restatement's view of a flattened scene (takes the product package's Scene only as a
    container of the float64 arrays and the ctypes struct layout)."""

    def __init__(self, scene):
        self.scene = scene
        self._desc = scene.desc()
        self.h = lib().oracle_scene_create(C.addressof(self._desc))
        self.pkg = importlib.import_module("monte-carlo-ray-tracer_b200")

    def close(self):
        if self.h:
            lib().oracle_scene_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trace(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        hits = np.zeros(len(rays), dtype=self.pkg.HIT_DTYPE)
        lib().oracle_trace(self.h, _p(rays), len(rays), _p(hits))
        return hits

    def sample_rays(self, rays, pixel, sample, seed):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        pixel = np.ascontiguousarray(pixel, dtype=np.uint32); sample = np.ascontiguousarray(sample, dtype=np.uint32)
        out = np.zeros((len(rays), 3))
        lib().oracle_sample_rays(self.h, _p(rays), _p(pixel), _p(sample), len(rays), seed, _p(out))
        return out

    def render_film(self, camera, sqrtspp, seed):
        """Whole frame through Film::deposit with the camera's reconstruction filter."""
        rec = camera.film_rec()
        default_radius = [0.5, 2.0, 2.0, 1.39, 1.0, 1.71, 2.0]   # film.cpp:32-45
        if rec.radius <= 0.0:
            rec.radius = default_radius[rec.filter]
        out = np.zeros((camera.height, camera.width, 3))
        lib().oracle_render_film(self.h, C.addressof(camera.rec), C.addressof(rec), sqrtspp, seed, _p(out))
        return out

    def trace_fast(self, nodes, scene_scale, rays):
        """The product's order-free search restated on the CPU over `nodes` (mcrt.bvh4_host). -> (hits, ambiguous flags, box tests, primitive tests)"""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        nodes = np.ascontiguousarray(nodes)
        hits = np.zeros(len(rays), dtype=self.pkg.HIT_DTYPE)
        flags = np.zeros(len(rays), dtype=np.uint8)
        bt, pt = C.c_uint64(), C.c_uint64()
        lib().oracle_trace_fast(self.h, _p(nodes), len(nodes), float(scene_scale), _p(rays), len(rays), _p(hits), _p(flags), C.byref(bt), C.byref(pt))
        return hits, flags.astype(bool), bt.value, pt.value

    def trace_visible(self, nodes, scene_scale, rays, target):
        """The product's occlusion query restated on the CPU. -> (verdict 0 visible / 1 not / 2 tie-to-replay, t of the target)"""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        target = np.ascontiguousarray(target, dtype=np.uint32)
        nodes = np.ascontiguousarray(nodes)
        verdict = np.zeros(len(rays), dtype=np.uint8); t = np.zeros(len(rays))
        lib().oracle_trace_visible(self.h, _p(nodes), float(scene_scale), _p(rays), _p(target), len(rays), _p(verdict), _p(t))
        return verdict, t

    def sample_pixels(self, camera, pixel, sample, seed):
        """-> (rgb [n,3], camera rays [n,6]) for (pixel, sample) pairs"""
        pixel = np.ascontiguousarray(pixel, dtype=np.uint32); sample = np.ascontiguousarray(sample, dtype=np.uint32)
        rgb = np.zeros((len(pixel), 3)); rays = np.zeros((len(pixel), 6))
        lib().oracle_sample_pixels(self.h, C.addressof(camera.rec), _p(pixel), _p(sample), len(pixel), seed, _p(rgb), _p(rays))
        return rgb, rays

    def render_rows_checked(self, nodes, scene_scale, camera, y0, y1, sqrtspp, seed):
        """render_rows with every Scene::intersect call cross-checked against the order-free search -> (image, calls, flagged, mismatches)"""
        nodes = np.ascontiguousarray(nodes)
        out = np.zeros((y1 - y0, camera.width, 3))
        counts = np.zeros(3, dtype=np.uint64)
        lib().oracle_render_rows_checked(self.h, _p(nodes), float(scene_scale), C.addressof(camera.rec), y0, y1, sqrtspp, seed, _p(out), _p(counts))
        return out, int(counts[0]), int(counts[1]), int(counts[2])

    def render_rows(self, camera, y0, y1, sqrtspp, seed):
        out = np.zeros((y1 - y0, camera.width, 3))
        rays = C.c_uint64()
        lib().oracle_render_rows(self.h, C.addressof(camera.rec), y0, y1, sqrtspp, seed, _p(out), C.byref(rays))
        return out, rays.value
