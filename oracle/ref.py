"""ORACLE / TEST INFRASTRUCTURE — ctypes binding of oracle/_ref/libmcrt_ref.so (the unmodified
reference compiled by oracle/build_ref.py). Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product never does."""
import ctypes as C
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libmcrt_ref.so")
SCENES_DIR = os.path.join(HERE, "_ref", "scenes")
REFERENCE_SCENES_DIR = "/root/reference/scenes"

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def scenes_dir():
    """Full asset tree when the reference checkout is present, else the copies made by build_ref."""
    return REFERENCE_SCENES_DIR if os.path.isdir(REFERENCE_SCENES_DIR) else SCENES_DIR


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libmcrt_ref.so missing: run python oracle/build_ref.py")
        L = C.CDLL(LIB_PATH)
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
        L.ref_close.argtypes = [C.c_void_p]
        L.ref_set_seed.argtypes = [C.c_uint32]
        L.ref_get_seed.restype = C.c_uint32
        L.ref_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint32)] * 6
        L.ref_render.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p,
                                 C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.ref_sample_pixels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.ref_sample_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_sampler_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
        L.ref_knn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_export_pack.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_bvh_build.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint32)]
        L.ref_bvh_arrays.argtypes = [C.c_void_p] * 6
        L.ref_prim_bounds.argtypes = [C.c_void_p] * 3
        L.ref_obj_parse.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_double)]
        L.ref_obj_copy.argtypes = [C.c_void_p] * 5
        L.ref_obj_adapter_check.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.ref_vertex_normals.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_double)]
        L.ref_image_save.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_fresnel_dielectric.restype = C.c_double
        L.ref_fresnel_dielectric.argtypes = [C.c_double] * 3
        L.ref_fresnel_conductor.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.ref_ggx_reflection.restype = C.c_double
        L.ref_ggx_reflection.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.POINTER(C.c_double)]
        L.ref_ggx_transmission.restype = C.c_double
        L.ref_ggx_transmission.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]
        L.ref_ggx_visible_microfacet.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_double, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def set_seed(seed):
    lib().ref_set_seed(seed & 0xFFFFFFFF)


class RefScene:
    """A reference Camera (+ PathTracer or PhotonMapper) built from a scene JSON with overrides."""

    def __init__(self, scene_file, overrides=None, camera_idx=0, photon_map=False, scenes=None):
        err = C.create_string_buffer(512)
        self.h = lib().ref_open((scenes or scenes_dir()).encode(), scene_file.encode(),
                                json.dumps(overrides or {}).encode(), camera_idx, int(photon_map), err, 512)
        if not self.h:
            raise RuntimeError("ref_open failed: " + err.value.decode())
        v = [C.c_uint32() for _ in range(6)]
        lib().ref_info(self.h, *[C.byref(x) for x in v])
        self.width, self.height, self.sqrtspp, self.n_prims, self.n_nodes, self.n_lights = [x.value for x in v]

    def close(self):
        if self.h:
            lib().ref_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, threads=1, y0=0, y1=None):
        """-> (rgb float64 [rows, W, 3], seconds, total_rays, shadow_rays)"""
        y1 = self.height if y1 is None else y1
        out = np.zeros((y1 - y0, self.width, 3), dtype=np.float64)
        sec, rays, sh = C.c_double(), C.c_uint64(), C.c_uint64()
        rc = lib().ref_render(self.h, threads, y0, y1, _p(out), C.byref(sec), C.byref(rays), C.byref(sh))
        if rc:
            raise RuntimeError("ref_render failed")
        return out, sec.value, rays.value, sh.value

    def sample_pixels(self, pixel, sample, radiance=True, rays=True):
        pixel = np.ascontiguousarray(pixel, dtype=np.uint32)
        sample = np.ascontiguousarray(sample, dtype=np.uint32)
        n = len(pixel)
        rgb = np.zeros((n, 3)) if radiance else None
        r6 = np.zeros((n, 6)) if rays else None
        lib().ref_sample_pixels(self.h, _p(pixel), _p(sample), n, _p(rgb), _p(r6))
        return rgb, r6

    def sample_rays(self, rays6, pixel, sample):
        rays6 = np.ascontiguousarray(rays6, dtype=np.float64)
        pixel = np.ascontiguousarray(pixel, dtype=np.uint32)
        sample = np.ascontiguousarray(sample, dtype=np.uint32)
        rgb = np.zeros((len(pixel), 3))
        lib().ref_sample_rays(self.h, _p(rays6), _p(pixel), _p(sample), len(pixel), _p(rgb))
        return rgb

    def trace(self, rays6):
        rays6 = np.ascontiguousarray(rays6, dtype=np.float64)
        n = len(rays6)
        t = np.zeros(n); prim = np.zeros(n, dtype=np.uint32); uv = np.zeros((n, 2)); ip = np.zeros(n, dtype=np.uint8)
        lib().ref_trace(self.h, _p(rays6), n, _p(t), _p(prim), _p(uv), _p(ip))
        return t, prim, uv, ip

    def knn(self, which, points, k):
        points = np.ascontiguousarray(points, dtype=np.float64)
        n = len(points)
        ph = np.zeros((n, k, 8), dtype=np.float32); d2 = np.full((n, k), np.inf); cnt = np.zeros(n, dtype=np.uint32)
        rc = lib().ref_knn(self.h, which, _p(points), n, k, _p(ph), _p(d2), _p(cnt))
        if rc:
            raise RuntimeError("ref_knn: scene was not opened with photon_map=True")
        return ph, d2, cnt

    def prim_bounds(self):
        """Surface::BB() of Scene::surfaces in order, and Scene::BB()."""
        b = np.zeros((self.n_prims, 6)); sb = np.zeros(6)
        lib().ref_prim_bounds(self.h, _p(b), _p(sb))
        return b, sb

    def build_bvh(self, bvh_type, bins=0):
        """BVH::BVH on this scene's surfaces, timed -> dict(node arrays, prim_order, seconds)."""
        sec = C.c_double(); nn = C.c_uint32()
        lib().ref_bvh_build(self.h, bvh_type.encode(), int(bins), C.byref(sec), C.byref(nn))
        n = nn.value
        out = dict(node_bounds=np.zeros((n, 6)), node_first_prim=np.zeros(n, np.uint32), node_prim_count=np.zeros(n, np.uint32),
                   node_next_sibling=np.zeros(n, np.uint32), prim_order=np.zeros(self.n_prims, np.uint32))
        if lib().ref_bvh_arrays(self.h, _p(out["node_bounds"]), _p(out["node_first_prim"]), _p(out["node_prim_count"]),
                                _p(out["node_next_sibling"]), _p(out["prim_order"])):
            raise RuntimeError("ref_bvh_arrays failed")
        out["seconds"] = sec.value
        return out

    def parse_obj(self, path):
        """Scene::parseOBJ -> dict(vertices, normals, tri_v, tri_vt, tri_vn, seconds); raises if the reference threw."""
        counts = np.zeros(5, dtype=np.uint64); sec = C.c_double()
        if lib().ref_obj_parse(self.h, os.fsencode(path), _p(counts), C.byref(sec)):
            raise RuntimeError("Scene::parseOBJ threw")
        c = [int(x) for x in counts]
        out = dict(vertices=np.zeros((c[0], 3)), normals=np.zeros((c[1], 3)), tri_v=np.zeros((c[2], 3), np.uint64),
                   tri_vt=np.zeros((c[3], 3), np.uint64), tri_vn=np.zeros((c[4], 3), np.uint64))
        lib().ref_obj_copy(_p(out["vertices"]), _p(out["normals"]), _p(out["tri_v"]), _p(out["tri_vt"]), _p(out["tri_vn"]))
        out["seconds"] = sec.value
        return out

    def obj_adapter_check(self, path, with_normals=True):
        """host/obj_adapter.hpp (the drop-in bodies for Scene::parseOBJ / generateVertexNormals) vs the reference's."""
        return lib().ref_obj_adapter_check(self.h, os.fsencode(path), int(with_normals))

    def vertex_normals(self, vertices, tri_v):
        """Scene::generateVertexNormals -> ([n, 3], seconds)."""
        vertices = np.ascontiguousarray(vertices, dtype=np.float64); tri_v = np.ascontiguousarray(tri_v, dtype=np.uint64)
        out = np.zeros_like(vertices); sec = C.c_double()
        if lib().ref_vertex_normals(self.h, _p(vertices), len(vertices), _p(tri_v), len(tri_v), _p(out), C.byref(sec)):
            raise RuntimeError("Scene::generateVertexNormals threw")
        return out, sec.value

    def export_pack(self, path):
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        if lib().ref_export_pack(self.h, path.encode()):
            raise RuntimeError("ref_export_pack failed")
        return path


def image_save(rgb, image):
    """Image::save on an [H, W, 3] float64 image; `image` = the camera's "image" JSON object without the
    size. -> (bytes [H, W, 3] in B,G,R order, exposure factor, gain factor)."""
    rgb = np.ascontiguousarray(rgb, dtype=np.float64)
    h, w = rgb.shape[:2]
    out = np.zeros((h, w, 3), dtype=np.uint8)
    e, g = C.c_double(), C.c_double()
    rc = lib().ref_image_save(json.dumps(dict(image, width=w, height=h)).encode(), _p(rgb), _p(out), C.byref(e), C.byref(g))
    if rc:
        raise RuntimeError(f"ref_image_save failed: {rc}")
    return out, e.value, g.value


def sampler_stream(pixel, sample, n_shuffles):
    pixel = np.ascontiguousarray(pixel, dtype=np.uint32)
    sample = np.ascontiguousarray(sample, dtype=np.uint32)
    out = np.zeros((len(pixel), 7), dtype=np.uint32)
    lib().ref_sampler_stream(_p(pixel), _p(sample), len(pixel), n_shuffles, _p(out))
    return out
