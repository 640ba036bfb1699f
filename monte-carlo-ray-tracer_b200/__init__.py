"""B200-native path-tracing integrator for the ray/BVH/BSDF hot path of
linusmossberg/monte-carlo-ray-tracer — Python host mirror over the C ABI (include/mcrt_abi.h).

The names follow the reference's classes for this path:
    Scene            flattened Scene + BVH          (source/scene/scene.hpp, source/bvh/bvh.hpp)
    Camera           camera state + sampleImage     (source/camera/camera.cpp:20-145)
    PathTracer       Integrator::sampleRay, batched (source/integrator/path-tracer/path-tracer.cpp)
    PhotonMapper     same with photon maps          (source/integrator/photon-mapper/photon-mapper.cpp)
Everything that computes runs in libmcrt_b200.so on the GPU; this module only marshals buffers.
There is no CPU fallback: importing works anywhere (so that symbols can be checked), but any compute
call without the built library or without a CUDA device raises."""
import ctypes as C
import math
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCRT_LIB", os.path.join(HERE, "libmcrt_b200.so"))  # MCRT_LIB: tuning variants

INTEGRATOR_PATH, INTEGRATOR_PHOTON = 0, 1
PRECISION_F64, PRECISION_F32 = 0, 1
PRIM_TRIANGLE, PRIM_SPHERE, PRIM_QUADRIC = 0, 1, 2
NO_PRIM = 0xFFFFFFFF

ABI_SYMBOLS = [
    "mcrt_abi_version", "mcrt_init", "mcrt_destroy", "mcrt_last_error", "mcrt_scene_upload",
    "mcrt_photon_upload", "mcrt_photon_emit", "mcrt_photon_download", "mcrt_octree_build",
    "mcrt_octree_free", "mcrt_render_rows", "mcrt_render_rows_dev", "mcrt_render_rows_strided_dev",
    "mcrt_trace_closest",
    "mcrt_sample_rays", "mcrt_sampler_stream", "mcrt_knn_search", "mcrt_set_option", "mcrt_set_film",
    "mcrt_obj_load", "mcrt_obj_free", "mcrt_obj_vertex_normals",
    "mcrt_bvh_build", "mcrt_bvh_free", "mcrt_image_tonemap", "mcrt_image_tonemap_dev",
    "mcrt_render_rows_strided_peers", "mcrt_frame_alloc", "mcrt_frame_open", "mcrt_frame_close", "mcrt_frame_free",
    "mcrt_render_film_sums_strided_dev", "mcrt_film_resolve_dev",
    "mcrt_bvh4_host", "mcrt_bvh4_host_free",
    "mcrt_fp64_peak", "mcrt_photon_emit_total", "mcrt_photon_emit_range", "mcrt_photon_build_dev",
]


class McrtError(RuntimeError):
    pass


# ------------------------------------------------------------------------------------ ctypes structs
class MaterialRec(C.Structure):
    _fields_ = [("reflectance", C.c_double * 3), ("specular_reflectance", C.c_double * 3),
                ("transmittance", C.c_double * 3), ("emittance", C.c_double * 3),
                ("roughness", C.c_double), ("specular_roughness", C.c_double), ("ior", C.c_double),
                ("transparency", C.c_double), ("complex_ior_real", C.c_double * 3),
                ("complex_ior_imag", C.c_double * 3), ("A", C.c_double), ("B", C.c_double),
                ("a", C.c_double * 2), ("has_complex_ior", C.c_uint32), ("perfect_mirror", C.c_uint32),
                ("rough", C.c_uint32), ("rough_specular", C.c_uint32), ("opaque", C.c_uint32),
                ("emissive", C.c_uint32), ("dirac_delta", C.c_uint32), ("_pad", C.c_uint32)]


MATERIAL_DTYPE = np.dtype([
    ("reflectance", "<f8", 3), ("specular_reflectance", "<f8", 3), ("transmittance", "<f8", 3),
    ("emittance", "<f8", 3), ("roughness", "<f8"), ("specular_roughness", "<f8"), ("ior", "<f8"),
    ("transparency", "<f8"), ("complex_ior_real", "<f8", 3), ("complex_ior_imag", "<f8", 3),
    ("A", "<f8"), ("B", "<f8"), ("a", "<f8", 2), ("has_complex_ior", "<u4"), ("perfect_mirror", "<u4"),
    ("rough", "<u4"), ("rough_specular", "<u4"), ("opaque", "<u4"), ("emissive", "<u4"),
    ("dirac_delta", "<u4"), ("_pad", "<u4")])
assert MATERIAL_DTYPE.itemsize == C.sizeof(MaterialRec)


class SceneDesc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("n_nodes", C.c_uint32),
                ("node_bounds", C.c_void_p), ("node_first_prim", C.c_void_p),
                ("node_prim_count", C.c_void_p), ("node_next_sibling", C.c_void_p),
                ("n_prims", C.c_uint32),
                ("prim_type", C.c_void_p), ("prim_index", C.c_void_p), ("prim_material", C.c_void_p),
                ("prim_area", C.c_void_p),
                ("n_tris", C.c_uint32),
                ("tri_v0", C.c_void_p), ("tri_v1", C.c_void_p), ("tri_v2", C.c_void_p),
                ("tri_e1", C.c_void_p), ("tri_e2", C.c_void_p), ("tri_normal", C.c_void_p),
                ("tri_vn_index", C.c_void_p),
                ("n_vertex_normals", C.c_uint32), ("vertex_normals", C.c_void_p),
                ("n_spheres", C.c_uint32), ("sphere_origin_radius", C.c_void_p),
                ("n_quadrics", C.c_uint32),
                ("quadric_Q", C.c_void_p), ("quadric_G", C.c_void_p), ("quadric_bounds", C.c_void_p),
                ("n_materials", C.c_uint32), ("materials", C.c_void_p),
                ("n_lights", C.c_uint32), ("light_prim", C.c_void_p), ("light_cdf", C.c_void_p),
                ("scene_ior", C.c_double)]


class CameraRec(C.Structure):
    _fields_ = [("eye", C.c_double * 3), ("forward", C.c_double * 3), ("left", C.c_double * 3),
                ("up", C.c_double * 3), ("focal_length", C.c_double), ("sensor_width", C.c_double),
                ("aperture_radius", C.c_double), ("focus_distance", C.c_double),
                ("width", C.c_uint32), ("height", C.c_uint32), ("thin_lens", C.c_uint32), ("_pad", C.c_uint32)]


class FilmRec(C.Structure):
    _fields_ = [("filter", C.c_uint32), ("cache_size", C.c_uint32), ("radius", C.c_double)]


FILM_FILTERS = {"box": 0, "mitchell-netravali": 1, "catmull-rom": 2, "b-spline": 3, "hermite": 4, "gaussian": 5,
                "lanczos": 6}


class HitRec(C.Structure):
    _fields_ = [("t", C.c_double), ("u", C.c_double), ("v", C.c_double), ("prim", C.c_uint32),
                ("interpolate", C.c_uint32)]


HIT_DTYPE = np.dtype([("t", "<f8"), ("u", "<f8"), ("v", "<f8"), ("prim", "<u4"), ("interpolate", "<u4")])


class PhotonMapDesc(C.Structure):
    _fields_ = [("n_octants", C.c_uint32), ("octant_bounds", C.c_void_p), ("octant_start", C.c_void_p),
                ("octant_count", C.c_void_p), ("octant_next_sibling", C.c_void_p), ("octant_leaf", C.c_void_p),
                ("n_photons", C.c_uint64), ("photons", C.c_void_p)]


class ImageParams(C.Structure):
    """The camera's "image" object (image.cpp:10-35) without the size."""
    _fields_ = [("plain", C.c_uint32), ("tonemapper", C.c_uint32), ("exposure_scale", C.c_double),
                ("gain_scale", C.c_double)]

    @classmethod
    def from_json(cls, image):
        tm = str(image.get("tonemapper", "HABLE")).upper()
        return cls(int(bool(image.get("plain", False))), 1 if tm == "ACES" else 0,
                   math.pow(2, image.get("exposure_compensation", 0.0)), math.pow(2, image.get("gain_compensation", 0.0)))


class ObjMeshRec(C.Structure):
    _fields_ = [("n_vertices", C.c_uint64), ("n_normals", C.c_uint64), ("n_tri_v", C.c_uint64), ("n_tri_vt", C.c_uint64),
                ("n_tri_vn", C.c_uint64), ("vertices", C.c_void_p), ("normals", C.c_void_p), ("tri_v", C.c_void_p),
                ("tri_vt", C.c_void_p), ("tri_vn", C.c_void_p)]


class BvhDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("n_prims", C.c_uint32), ("node_bounds", C.c_void_p),
                ("node_first_prim", C.c_void_p), ("node_prim_count", C.c_void_p), ("node_next_sibling", C.c_void_p),
                ("prim_order", C.c_void_p), ("build_rounds", C.c_uint32), ("kernel_launches", C.c_uint32)]


BVH_TYPES = {"octree": 0, "binary_sah": 1, "quaternary_sah": 2}


class PhotonEmitParams(C.Structure):
    _fields_ = [("emissions", C.c_uint64), ("caustic_factor", C.c_double), ("max_photons_per_octree_leaf", C.c_uint32),
                ("k_nearest_photons", C.c_uint32), ("direct_visualization", C.c_uint32), ("global_seed", C.c_uint32),
                ("scene_bounds", C.c_double * 6)]


class Stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("extension_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("box_tests", C.c_uint64), ("prim_tests", C.c_uint64), ("knn_queries", C.c_uint64),
                ("wavefront_iterations", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("ior_stack_overflows", C.c_uint64), ("max_depth", C.c_uint32), ("_pad", C.c_uint32),
                ("gpu_ms_total", C.c_double), ("gpu_ms_generate", C.c_double), ("gpu_ms_extend", C.c_double),
                ("gpu_ms_shade", C.c_double), ("gpu_ms_shadow", C.c_double), ("gpu_ms_knn", C.c_double),
                ("extend_launches", C.c_uint64), ("shadow_launches", C.c_uint64),
                ("shadow_box_tests", C.c_uint64), ("shadow_prim_tests", C.c_uint64),
                ("extend_work_sum", C.c_uint64), ("extend_work_warpmax", C.c_uint64),
                ("replayed_rays", C.c_uint64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_ if not f.startswith("_")}


_lib = None


def lib():
    """The C-ABI library. Raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise McrtError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        L.mcrt_abi_version.restype = C.c_int
        L.mcrt_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.mcrt_destroy.argtypes = [C.c_void_p]
        L.mcrt_destroy.restype = None
        L.mcrt_last_error.argtypes = [C.c_void_p]
        L.mcrt_last_error.restype = C.c_char_p
        L.mcrt_scene_upload.argtypes = [C.c_void_p, C.POINTER(SceneDesc), C.POINTER(C.c_uint64)]
        L.mcrt_photon_upload.argtypes = [C.c_void_p, C.POINTER(PhotonMapDesc), C.POINTER(PhotonMapDesc),
                                         C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
        L.mcrt_photon_emit.argtypes = [C.c_void_p, C.POINTER(PhotonEmitParams), C.c_int, C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_uint64), C.POINTER(Stats)]
        L.mcrt_photon_download.argtypes = [C.c_void_p, C.c_int, C.POINTER(PhotonMapDesc)]
        L.mcrt_octree_build.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p),
                                        C.POINTER(PhotonMapDesc), C.POINTER(C.c_double)]
        L.mcrt_octree_free.argtypes = [C.c_void_p]
        L.mcrt_octree_free.restype = None
        render_args = [C.c_void_p, C.POINTER(CameraRec), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                       C.c_int, C.c_int, C.c_void_p, C.POINTER(Stats)]
        L.mcrt_render_rows.argtypes = render_args
        L.mcrt_render_rows_dev.argtypes = render_args
        L.mcrt_render_rows_strided_dev.argtypes = [C.c_void_p, C.POINTER(CameraRec), C.c_uint32, C.c_uint32, C.c_uint32,
                                                   C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.POINTER(Stats)]
        L.mcrt_render_rows_strided_peers.argtypes = [C.c_void_p, C.POINTER(CameraRec), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                     C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(Stats)]
        L.mcrt_frame_alloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.c_void_p]
        L.mcrt_frame_open.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.mcrt_frame_close.argtypes = [C.c_void_p, C.c_void_p]
        L.mcrt_frame_free.argtypes = [C.c_void_p, C.c_void_p]
        L.mcrt_fp64_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.mcrt_bvh4_host.argtypes = [C.POINTER(SceneDesc), C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]
        L.mcrt_bvh4_host_free.argtypes = [C.c_void_p]
        L.mcrt_bvh4_host_free.restype = None
        L.mcrt_render_film_sums_strided_dev.argtypes = [C.c_void_p, C.POINTER(CameraRec), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                        C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        L.mcrt_film_resolve_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.mcrt_photon_emit_total.argtypes = [C.c_void_p, C.POINTER(PhotonEmitParams), C.POINTER(C.c_uint64)]
        L.mcrt_photon_emit_range.argtypes = [C.c_void_p, C.POINTER(PhotonEmitParams), C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p),
                                             C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(Stats)]
        L.mcrt_photon_build_dev.argtypes = [C.c_void_p, C.POINTER(PhotonEmitParams), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                            C.POINTER(C.c_double)]
        L.mcrt_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(Stats)]
        L.mcrt_sample_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32,
                                       C.c_int, C.c_int, C.c_void_p, C.POINTER(Stats)]
        L.mcrt_sampler_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p]
        L.mcrt_knn_search.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.POINTER(Stats)]
        L.mcrt_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.mcrt_set_film.argtypes = [C.c_void_p, C.POINTER(FilmRec)]
        L.mcrt_bvh_build.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int,
                                     C.POINTER(C.c_void_p), C.POINTER(BvhDesc), C.POINTER(C.c_double)]
        L.mcrt_bvh_free.argtypes = [C.c_void_p]
        L.mcrt_bvh_free.restype = None
        L.mcrt_obj_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(ObjMeshRec), C.c_char_p, C.c_size_t]
        L.mcrt_obj_free.argtypes = [C.c_void_p]
        L.mcrt_obj_free.restype = None
        L.mcrt_obj_vertex_normals.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        tm_args = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(ImageParams), C.c_void_p,
                   C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.mcrt_image_tonemap.argtypes = tm_args
        L.mcrt_image_tonemap_dev.argtypes = tm_args
        if L.mcrt_abi_version() != 1:
            raise McrtError("libmcrt_b200.so ABI version mismatch")
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


# ------------------------------------------------------------------------------------ scene packs
_PACK_DTYPES = {0: np.uint8, 1: np.uint32, 2: np.int32, 3: np.uint64, 4: np.float32, 5: np.float64}


def read_pack(path):
    """Reads a scene pack written by host/exporter.cpp (PackWriter) → dict of numpy arrays."""
    if path.endswith(".xz"):   # large OBJ-scene packs are shipped xz-compressed (float64 arrays: ~5x)
        import lzma
        with lzma.open(path, "rb") as f:
            blob = f.read()
    else:
        with open(path, "rb") as f:
            blob = f.read()
    if blob[:8] != b"MCRTPK01":
        raise McrtError(f"{path}: not a scene pack")
    n, = struct.unpack_from("<I", blob, 8)
    out = {}
    for i in range(n):
        name, dtype, elem_size, count, offset = struct.unpack_from("<32sIIQQ", blob, 16 + 56 * i)
        name = name.split(b"\0")[0].decode()
        if dtype == 6:
            if name != "materials" or elem_size != MATERIAL_DTYPE.itemsize:
                raise McrtError(f"{path}: unexpected struct entry {name}")
            arr = np.frombuffer(blob, dtype=MATERIAL_DTYPE, count=count, offset=offset)
        else:
            arr = np.frombuffer(blob, dtype=_PACK_DTYPES[dtype], count=count, offset=offset)
        out[name] = arr.copy()
    return out


class Scene:
    """Flattened reference Scene (surfaces, materials, emissives, BVH) as float64 arrays."""

    _ARRAYS = ["node_bounds", "node_first_prim", "node_prim_count", "node_next_sibling", "prim_type",
               "prim_index", "prim_material", "prim_area", "tri_v0", "tri_v1", "tri_v2", "tri_e1", "tri_e2",
               "tri_normal", "tri_vn_index", "vertex_normals", "sphere_origin_radius", "quadric_Q",
               "quadric_G", "quadric_bounds", "materials", "light_prim", "light_cdf"]

    def __init__(self, arrays):
        self.a = {k: np.ascontiguousarray(arrays[k]) for k in self._ARRAYS}
        self.ior = float(np.asarray(arrays["scene_ior"]).reshape(-1)[0])
        self.extra = {k: v for k, v in arrays.items() if k not in self._ARRAYS}

    @classmethod
    def from_pack(cls, path):
        return cls(read_pack(path))

    @property
    def n_prims(self):
        return int(self.a["prim_type"].size)

    @property
    def n_nodes(self):
        return int(self.a["node_first_prim"].size)

    @property
    def n_lights(self):
        return int(self.a["light_prim"].size)

    def desc(self):
        a = self.a
        d = SceneDesc()
        d.abi_version = 1
        d.n_nodes = a["node_first_prim"].size
        d.n_prims = a["prim_type"].size
        d.n_tris = a["tri_vn_index"].size
        d.n_vertex_normals = a["vertex_normals"].size // 9
        d.n_spheres = a["sphere_origin_radius"].size // 4
        d.n_quadrics = a["quadric_bounds"].size // 6
        d.n_materials = a["materials"].size
        d.n_lights = a["light_prim"].size
        for k in self._ARRAYS:
            setattr(d, k, _ptr(a[k]))
        d.scene_ior = self.ior
        return d

    # -- inputs / outputs of BVH::BVH (mcrt_bvh_build)
    def prim_bounds(self):
        """Surface::Base::BB() of every primitive, [n_prims, 6] (triangle.cpp:115-122, sphere.cpp:56-62,
        quadric BB_)."""
        a = self.a
        out = np.zeros((self.n_prims, 6))
        t, i = a["prim_type"], a["prim_index"]
        tri = t == PRIM_TRIANGLE
        if tri.any():
            v = np.stack([a["tri_v0"].reshape(-1, 3), a["tri_v1"].reshape(-1, 3), a["tri_v2"].reshape(-1, 3)])[:, i[tri]]
            out[tri, :3] = v.min(axis=0); out[tri, 3:] = v.max(axis=0)
        sph = t == PRIM_SPHERE
        if sph.any():
            s = a["sphere_origin_radius"].reshape(-1, 4)[i[sph]]
            out[sph, :3] = s[:, :3] - s[:, 3:4]; out[sph, 3:] = s[:, :3] + s[:, 3:4]
        quad = t == PRIM_QUADRIC
        if quad.any():
            out[quad] = a["quadric_bounds"].reshape(-1, 6)[i[quad]]
        return out

    def reordered(self, order, bvh=None):
        """Scene whose primitive k is this scene's primitive order[k]; node arrays from `bvh` (a dict as
        returned by bvh_build) or none (Scene::intersect then scans all primitives, scene.cpp:159-171)."""
        order = np.asarray(order, dtype=np.int64)
        arrays = dict(self.a, **self.extra)
        arrays["scene_ior"] = np.array([self.ior])
        for k in ("prim_type", "prim_index", "prim_material", "prim_area"):
            arrays[k] = self.a[k][order]
        inverse = np.empty(len(order), dtype=np.int64)
        inverse[order] = np.arange(len(order))
        arrays["light_prim"] = inverse[self.a["light_prim"]].astype(np.uint32)
        if "prim_original" in arrays:
            arrays["prim_original"] = arrays["prim_original"][order]
        for k, dt in (("node_bounds", np.float64), ("node_first_prim", np.uint32), ("node_prim_count", np.uint32),
                      ("node_next_sibling", np.uint32)):
            arrays[k] = np.ascontiguousarray(bvh[k], dtype=dt).reshape(-1) if bvh is not None else np.zeros(0, dtype=dt)
        return Scene(arrays)

    def unbuilt(self):
        """The scene as BVH::BVH receives it: primitives in Scene::surfaces order, no hierarchy."""
        return self.reordered(np.argsort(self.extra["prim_original"], kind="stable"))

    def with_bvh(self, bvh):
        return self.reordered(bvh["prim_order"], bvh)

    def cameras(self):
        """Cameras stored in the pack (the exporter writes the one the scene was opened with)."""
        cams = []
        if "camera_f64" in self.extra:
            cams.append(Camera.from_pack_arrays(self.extra["camera_f64"], self.extra["camera_u32"],
                                                self.extra.get("camera_film_u32"), self.extra.get("camera_film_f64")))
        return cams

    def photon_maps(self):
        e = self.extra
        if "photon_params" not in e:
            return None
        maps = []
        for prefix in ("caustic", "global"):
            maps.append({k: e[f"{prefix}_{k}"] for k in
                         ("octant_bounds", "octant_start", "octant_count", "octant_next", "octant_leaf", "photons")})
        return maps[0], maps[1], int(e["photon_params"][0]), int(e["photon_params"][1])


class Camera:
    """Camera state after the reference's Camera::Camera (camera.cpp:20-64)."""

    def __init__(self, eye, forward, left, up, focal_length, sensor_width, width, height,
                 aperture_radius=-1.0, focus_distance=-1.0, thin_lens=False, sqrtspp=1, film=None):
        """film: None (default box film) or dict(filter=name|code, radius=None, cache_size=0), the camera's
        "film" object in the scene JSON (source/camera/film.cpp:19-59)."""
        self.film = dict(film) if film else None
        self.rec = CameraRec()
        for name, v in (("eye", eye), ("forward", forward), ("left", left), ("up", up)):
            for i in range(3):
                getattr(self.rec, name)[i] = float(v[i])
        self.rec.focal_length = focal_length
        self.rec.sensor_width = sensor_width
        self.rec.aperture_radius = aperture_radius
        self.rec.focus_distance = focus_distance
        self.rec.width, self.rec.height = int(width), int(height)
        self.rec.thin_lens = int(bool(thin_lens))
        self.sqrtspp = int(sqrtspp)

    @classmethod
    def from_pack_arrays(cls, f64, u32, film_u32=None, film_f64=None):
        film = None
        if film_u32 is not None and not (int(film_u32[0]) == 0 and float(film_f64[0]) == 0.5):
            film = dict(filter=int(film_u32[0]), cache_size=int(film_u32[1]), radius=float(film_f64[0]))
        return cls(f64[0:3], f64[3:6], f64[6:9], f64[9:12], f64[12], f64[13], u32[0], u32[1],
                   f64[14], f64[15], bool(u32[2]), int(u32[3]), film)

    def film_rec(self):
        """mcrt_film of this camera, or None for the default box film."""
        if not self.film:
            return None
        f = self.film.get("filter", "box")
        code = FILM_FILTERS[f.lower()] if isinstance(f, str) else int(f)
        radius = self.film.get("radius")
        return FilmRec(code, int(self.film.get("cache_size") or 0), float(radius) if radius else 0.0)

    @property
    def width(self):
        return self.rec.width

    @property
    def height(self):
        return self.rec.height

    def resized(self, width, height, sqrtspp=None):
        c = Camera(self.rec.eye, self.rec.forward, self.rec.left, self.rec.up, self.rec.focal_length,
                   self.rec.sensor_width, width, height, self.rec.aperture_radius, self.rec.focus_distance,
                   self.rec.thin_lens, self.sqrtspp if sqrtspp is None else sqrtspp, self.film)
        return c


class Integrator:
    """GPU context + uploaded scene. Mirrors class Integrator (source/integrator/integrator.hpp:7-30):
    owns the Scene, exposes sampleRay (batched) and is what Camera.sampleImage renders through."""

    kind = INTEGRATOR_PATH

    def __init__(self, scene, device=0, precision=PRECISION_F64, global_seed=0x12345678):
        self.scene = scene
        self.precision = precision
        self.global_seed = global_seed & 0xFFFFFFFF
        self.ctx = C.c_void_p()
        rc = lib().mcrt_init(device, C.byref(self.ctx))
        if rc:
            self.ctx = C.c_void_p()
            raise McrtError(f"mcrt_init(device={device}) failed with {rc}: no CUDA device? (there is no CPU fallback)")
        self.device = device
        self.h2d_bytes = 0
        self.upload_scene()
        self.last_stats = None

    def _check(self, rc):
        if rc:
            raise McrtError(f"mcrt error {rc}: {lib().mcrt_last_error(self.ctx).decode()}")

    def set_option(self, key, value):
        self._check(lib().mcrt_set_option(self.ctx, key.encode(), float(value)))

    def upload_scene(self):
        d = self.scene.desc()
        n = C.c_uint64()
        self._check(lib().mcrt_scene_upload(self.ctx, C.byref(d), C.byref(n)))
        self.h2d_bytes = n.value
        return n.value

    def close(self):
        if getattr(self, "ctx", None):
            lib().mcrt_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- Integrator::sampleRay, batched
    def sampleRay(self, rays, pixel, sample, precision=None):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        pixel = np.ascontiguousarray(pixel, dtype=np.uint32)
        sample = np.ascontiguousarray(sample, dtype=np.uint32)
        out = np.zeros((len(rays), 3))
        st = Stats()
        self._check(lib().mcrt_sample_rays(self.ctx, _ptr(rays), _ptr(pixel), _ptr(sample), len(rays),
                                           self.global_seed, self.kind,
                                           self.precision if precision is None else precision, _ptr(out), C.byref(st)))
        self.last_stats = st.as_dict()
        return out

    # -- Scene::intersect, batched
    def intersect(self, rays, precision=None):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        st = Stats()
        self._check(lib().mcrt_trace_closest(self.ctx, _ptr(rays), len(rays),
                                             self.precision if precision is None else precision, _ptr(hits), C.byref(st)))
        self.last_stats = st.as_dict()
        return hits

    # -- Camera::sampleImage for a block of rows (box film), host output
    def set_film(self, camera):
        """Film of the following renders = the camera's (Camera owns its Film, camera.cpp:34-37)."""
        rec = camera.film_rec()
        self._check(lib().mcrt_set_film(self.ctx, C.byref(rec) if rec is not None else None))

    def render_rows(self, camera, y0=0, y1=None, sqrtspp=None, precision=None, out=None):
        y1 = camera.height if y1 is None else y1
        self.set_film(camera)
        if out is None:
            out = np.zeros((y1 - y0, camera.width, 3))
        st = Stats()
        self._check(lib().mcrt_render_rows(self.ctx, C.byref(camera.rec), y0, y1,
                                           camera.sqrtspp if sqrtspp is None else sqrtspp, self.global_seed, self.kind,
                                           self.precision if precision is None else precision, _ptr(out), C.byref(st)))
        self.last_stats = st.as_dict()
        return out

    # -- same, framebuffer stays in HBM (raw device pointer, float64 [rows*W*3])
    def render_rows_dev(self, camera, out_dev_ptr, y0=0, y1=None, sqrtspp=None, precision=None):
        y1 = camera.height if y1 is None else y1
        self.set_film(camera)
        st = Stats()
        self._check(lib().mcrt_render_rows_dev(self.ctx, C.byref(camera.rec), y0, y1,
                                               camera.sqrtspp if sqrtspp is None else sqrtspp, self.global_seed,
                                               self.kind, self.precision if precision is None else precision,
                                               C.c_void_p(out_dev_ptr), C.byref(st)))
        self.last_stats = st.as_dict()
        return self.last_stats

    # -- interleaved rows y_first + k*y_step (multi-GPU sharding), framebuffer in HBM
    def render_rows_strided_peers(self, camera, frame_ptrs, y_first, y_step, n_rows, frame_is_float32=True, sqrtspp=None, precision=None):
        """Rows y_first + k*y_step rendered and resolved straight into the full-frame buffers `frame_ptrs` (this
        rank's and the peers', see distributed.PeerFrames): mcrt_render_rows_strided_peers."""
        self.set_film(camera)
        arr = (C.c_void_p * len(frame_ptrs))(*[C.c_void_p(int(q)) for q in frame_ptrs])
        st = Stats()
        self._check(lib().mcrt_render_rows_strided_peers(self.ctx, C.byref(camera.rec), y_first, y_step, n_rows,
                                                         camera.sqrtspp if sqrtspp is None else sqrtspp, self.global_seed, self.kind,
                                                         self.precision if precision is None else precision, arr, len(frame_ptrs),
                                                         1 if frame_is_float32 else 0, C.byref(st)))
        self.last_stats = st.as_dict()
        return self.last_stats

    def render_film_sums_strided_dev(self, camera, rgb_sum_ptr, weight_sum_ptr, y_first, y_step, n_rows, sqrtspp=None, precision=None):
        """Row shard of a render through the camera's reconstruction filter: unresolved whole-frame sums (device)."""
        self.set_film(camera)
        st = Stats()
        self._check(lib().mcrt_render_film_sums_strided_dev(self.ctx, C.byref(camera.rec), y_first, y_step, n_rows,
                                                            camera.sqrtspp if sqrtspp is None else sqrtspp, self.global_seed, self.kind,
                                                            self.precision if precision is None else precision,
                                                            C.c_void_p(rgb_sum_ptr), C.c_void_p(weight_sum_ptr), C.byref(st)))
        self.last_stats = st.as_dict()
        return self.last_stats

    def film_resolve_dev(self, rgb_sum_ptr, weight_sum_ptr, n_pixels, out_ptr):
        self._check(lib().mcrt_film_resolve_dev(self.ctx, C.c_void_p(rgb_sum_ptr), C.c_void_p(weight_sum_ptr), n_pixels, C.c_void_p(out_ptr)))

    def frame_alloc(self, nbytes):
        """-> (device pointer, 64-byte CUDA IPC handle) of a zero-filled buffer other ranks can map"""
        ptr = C.c_void_p(); h = (C.c_ubyte * 64)()
        self._check(lib().mcrt_frame_alloc(self.ctx, nbytes, C.byref(ptr), h))
        return ptr.value, bytes(h)

    def frame_open(self, handle):
        ptr = C.c_void_p(); h = (C.c_ubyte * 64)(*handle)
        self._check(lib().mcrt_frame_open(self.ctx, h, C.byref(ptr)))
        return ptr.value

    def frame_close(self, ptr):
        self._check(lib().mcrt_frame_close(self.ctx, C.c_void_p(ptr)))

    def frame_free(self, ptr):
        self._check(lib().mcrt_frame_free(self.ctx, C.c_void_p(ptr)))

    def fp64_peak(self):
        """Measured DFMA thread-instructions per second of this GPU (mcrt_fp64_peak)."""
        v = C.c_double()
        self._check(lib().mcrt_fp64_peak(self.ctx, C.byref(v)))
        return v.value

    def render_rows_strided_dev(self, camera, out_dev_ptr, y_first, y_step, n_rows, sqrtspp=None, precision=None):
        self.set_film(camera)
        st = Stats()
        self._check(lib().mcrt_render_rows_strided_dev(self.ctx, C.byref(camera.rec), y_first, y_step, n_rows,
                                                       camera.sqrtspp if sqrtspp is None else sqrtspp,
                                                       self.global_seed, self.kind,
                                                       self.precision if precision is None else precision,
                                                       C.c_void_p(out_dev_ptr), C.byref(st)))
        self.last_stats = st.as_dict()
        return self.last_stats

    # -- Image::save without the file: float64 [H, W, 3] -> bytes [H, W, 3] in B,G,R order (the .tga payload)
    def tonemap(self, rgb, image=None):
        params = image if isinstance(image, ImageParams) else ImageParams.from_json(image or {})
        rgb = np.ascontiguousarray(rgb, dtype=np.float64)
        h, w = rgb.shape[:2]
        out = np.zeros((h, w, 3), dtype=np.uint8)
        e, g = C.c_double(), C.c_double()
        self._check(lib().mcrt_image_tonemap(self.ctx, _ptr(rgb), w, h, C.byref(params), _ptr(out), C.byref(e), C.byref(g)))
        return out, e.value, g.value

    def tonemap_dev(self, rgb_dev_ptr, out_dev_ptr, width, height, image=None):
        params = image if isinstance(image, ImageParams) else ImageParams.from_json(image or {})
        e, g = C.c_double(), C.c_double()
        self._check(lib().mcrt_image_tonemap_dev(self.ctx, C.c_void_p(rgb_dev_ptr), width, height, C.byref(params),
                                                 C.c_void_p(out_dev_ptr), C.byref(e), C.byref(g)))
        return e.value, g.value

    def sampler_stream(self, pixel, sample, n_shuffles):
        pixel = np.ascontiguousarray(pixel, dtype=np.uint32)
        sample = np.ascontiguousarray(sample, dtype=np.uint32)
        out = np.zeros((len(pixel), 7), dtype=np.uint32)
        self._check(lib().mcrt_sampler_stream(self.ctx, _ptr(pixel), _ptr(sample), len(pixel), n_shuffles,
                                              self.global_seed, _ptr(out)))
        return out


class PathTracer(Integrator):
    kind = INTEGRATOR_PATH


class PhotonMapper(Integrator):
    """Renders with the caustic/global photon maps the CPU photon pass produced (first pass stays on
    the CPU); maps come from the scene pack or from explicit arrays."""
    kind = INTEGRATOR_PHOTON

    def __init__(self, scene, device=0, precision=PRECISION_F64, global_seed=0x12345678, photon_maps=None, emit=None):
        """photon_maps: (caustic, global, k, direct_visualization) built by the reference's CPU pass (default: the
        ones in the scene pack); emit: dict(emissions, caustic_factor, max_photons_per_octree_leaf, k_nearest_photons,
        direct_visualization, scene_bounds) to run the photon pass on the GPU instead (mcrt_photon_emit)."""
        super().__init__(scene, device, precision, global_seed)
        if emit is not None:
            self.emit(**emit)
            return
        maps = photon_maps or scene.photon_maps()
        if maps is None:
            raise McrtError("PhotonMapper needs photon maps (scene pack exported with photon_map=True) or emit=...")
        self._maps = maps
        self.upload_photons()

    def _emit_params(self, emissions, caustic_factor, max_photons_per_octree_leaf, k_nearest_photons, direct_visualization, scene_bounds):
        p = PhotonEmitParams()
        p.emissions = int(emissions); p.caustic_factor = float(caustic_factor)
        p.max_photons_per_octree_leaf = int(max_photons_per_octree_leaf); p.k_nearest_photons = int(k_nearest_photons)
        p.direct_visualization = int(bool(direct_visualization)); p.global_seed = self.global_seed
        b = scene_bounds if scene_bounds is not None else self.scene.a["node_bounds"][:6]
        for i in range(6):
            p.scene_bounds[i] = float(b[i])
        return p

    def emit_sharded(self, rank, world, emissions, caustic_factor, max_photons_per_octree_leaf=200, k_nearest_photons=50,
                     direct_visualization=False, scene_bounds=None, precision=None):
        """The photon pass over `world` GPUs: this rank emits its range of the emission index space
        (distributed.emission_range), the photon arrays are all-gathered (torch.distributed), and every rank builds the
        same two octrees from the concatenation. -> (n_caustic, n_global) of the whole maps."""
        import torch
        from . import distributed as mdist
        p = self._emit_params(emissions, caustic_factor, max_photons_per_octree_leaf, k_nearest_photons, direct_visualization, scene_bounds)
        prec = self.precision if precision is None else precision
        total = C.c_uint64()
        self._check(lib().mcrt_photon_emit_total(self.ctx, C.byref(p), C.byref(total)))
        first, count = mdist.emission_range(total.value, rank, world)
        ptr = [C.c_void_p(), C.c_void_p()]; n = [C.c_uint64(), C.c_uint64()]; st = Stats()
        self._check(lib().mcrt_photon_emit_range(self.ctx, C.byref(p), prec, first, count, C.byref(ptr[0]), C.byref(n[0]),
                                                 C.byref(ptr[1]), C.byref(n[1]), C.byref(st)))
        self.last_stats = st.as_dict()
        dev = torch.device("cuda", self.device)
        gathered = []
        for w in range(2):
            mine = mdist.device_view(ptr[w].value, n[w].value * 8, torch.float32, dev)
            gathered.append(mdist.all_gather_photons(mine, world, dev))
        ms = C.c_double()
        self._check(lib().mcrt_photon_build_dev(self.ctx, C.byref(p), C.c_void_p(gathered[0].data_ptr()), gathered[0].numel() // 8,
                                                C.c_void_p(gathered[1].data_ptr()), gathered[1].numel() // 8, C.byref(ms)))
        torch.cuda.synchronize(dev)
        self.last_stats["gpu_ms_knn"] = ms.value
        self.k_nearest = int(k_nearest_photons)
        self._host_maps = None
        self._emitted = (int(k_nearest_photons), int(bool(direct_visualization)))
        self.n_photons = (gathered[0].numel() // 8, gathered[1].numel() // 8)
        return self.n_photons

    def emit(self, emissions, caustic_factor, max_photons_per_octree_leaf=200, k_nearest_photons=50,
             direct_visualization=False, scene_bounds=None, precision=None):
        """PhotonMapper::PhotonMapper's first pass on the GPU; replaces the uploaded maps."""
        p = self._emit_params(emissions, caustic_factor, max_photons_per_octree_leaf, k_nearest_photons, direct_visualization, scene_bounds)
        nc, ng, st = C.c_uint64(), C.c_uint64(), Stats()
        self._check(lib().mcrt_photon_emit(self.ctx, C.byref(p), self.precision if precision is None else precision,
                                           C.byref(nc), C.byref(ng), C.byref(st)))
        self.last_stats = st.as_dict()
        self.k_nearest = int(k_nearest_photons)
        # the maps stay in HBM (octrees are built there too); host copies only when somebody asks
        self._host_maps = None
        self._emitted = (int(k_nearest_photons), int(bool(direct_visualization)))
        self.n_photons = (nc.value, ng.value)
        return nc.value, ng.value

    @property
    def _maps(self):
        if self._host_maps is None and getattr(self, "_emitted", None) is not None:
            maps = []
            for which in (0, 1):
                d = PhotonMapDesc()
                self._check(lib().mcrt_photon_download(self.ctx, which, C.byref(d)))
                maps.append(_map_arrays(d))
            self._host_maps = (maps[0], maps[1]) + self._emitted
        return self._host_maps

    @_maps.setter
    def _maps(self, value):
        self._host_maps = value
        self._emitted = None

    @staticmethod
    def _map_desc(m):
        d = PhotonMapDesc()
        d.n_octants = m["octant_leaf"].size
        d.octant_bounds = _ptr(m["octant_bounds"]); d.octant_start = _ptr(m["octant_start"])
        d.octant_count = _ptr(m["octant_count"]); d.octant_next_sibling = _ptr(m["octant_next"])
        d.octant_leaf = _ptr(m["octant_leaf"])
        d.n_photons = m["photons"].size // 8
        d.photons = _ptr(m["photons"])
        return d

    def upload_photons(self):
        caustic, glob, k, dv = self._maps
        dc, dg = self._map_desc(caustic), self._map_desc(glob)
        n = C.c_uint64()
        self._check(lib().mcrt_photon_upload(self.ctx, C.byref(dc), C.byref(dg), k, dv, C.byref(n)))
        self.k_nearest = k
        return n.value

    def knn(self, which, points):
        points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        n, k = len(points), self.k_nearest
        idx = np.full((n, k), NO_PRIM, dtype=np.uint32); d2 = np.full((n, k), np.inf); cnt = np.zeros(n, dtype=np.uint32)
        st = Stats()
        self._check(lib().mcrt_knn_search(self.ctx, which, _ptr(points), n, _ptr(idx), _ptr(d2), _ptr(cnt), C.byref(st)))
        return idx, d2, cnt


def _map_arrays(d):
    """PhotonMapDesc (host pointers) -> dict of numpy copies."""
    n, npn = d.n_octants, d.n_photons

    def arr(ptr, ctype, count, dtype):
        if not ptr or count == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).astype(dtype, copy=True)
    return {"octant_bounds": arr(d.octant_bounds, C.c_double, 6 * n, np.float64),
            "octant_start": arr(d.octant_start, C.c_uint64, n, np.uint64),
            "octant_count": arr(d.octant_count, C.c_uint64, n, np.uint64),
            "octant_next": arr(d.octant_next_sibling, C.c_uint32, n, np.uint32),
            "octant_leaf": arr(d.octant_leaf, C.c_uint8, n, np.uint8),
            "photons": arr(d.photons, C.c_float, 8 * npn, np.float32)}


def build_photon_octree(photons, max_photons_per_octree_leaf, scene_bounds, device=0):
    """mcrt_octree_build: the octree construction of mcrt_photon_emit on the GPU, on caller photons.
    -> (dict of octant_bounds, octant_start, octant_count, octant_next, octant_leaf, photons; gpu_ms)."""
    photons = np.ascontiguousarray(photons, dtype=np.float32).reshape(-1, 8)
    bounds = np.ascontiguousarray(scene_bounds, dtype=np.float64)
    ctx = C.c_void_p()
    rc = lib().mcrt_init(device, C.byref(ctx))
    if rc:
        raise McrtError(f"mcrt_init({device}) failed: {rc} (no CUDA device? there is no CPU fallback)")
    try:
        h, d, ms = C.c_void_p(), PhotonMapDesc(), C.c_double()
        rc = lib().mcrt_octree_build(ctx, _ptr(photons), len(photons), int(max_photons_per_octree_leaf), _ptr(bounds),
                                     C.byref(h), C.byref(d), C.byref(ms))
        if rc:
            raise McrtError(f"mcrt_octree_build failed ({rc}): {lib().mcrt_last_error(ctx).decode()}")
        out = _map_arrays(d)
        lib().mcrt_octree_free(h)
        return out, ms.value
    finally:
        lib().mcrt_destroy(ctx)


def load_obj(path, threads=0):
    """mcrt_obj_load: Scene::parseOBJ (scene.cpp:238-324), parallel. -> dict(vertices [n,3], normals [n,3],
    tri_v / tri_vt / tri_vn [m,3] uint64)."""
    h, d = C.c_void_p(), ObjMeshRec()
    err = C.create_string_buffer(512)
    rc = lib().mcrt_obj_load(os.fsencode(path), int(threads), C.byref(h), C.byref(d), err, 512)
    if rc:
        raise McrtError(err.value.decode() or f"mcrt_obj_load failed: {rc}")

    def arr(ptr, count, dtype):
        if not ptr or count == 0:
            return np.zeros((0, 3), dtype=dtype)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count * 3 * 8,)).view(dtype).reshape(-1, 3).copy()
    out = dict(vertices=arr(d.vertices, d.n_vertices, np.float64), normals=arr(d.normals, d.n_normals, np.float64),
               tri_v=arr(d.tri_v, d.n_tri_v, np.uint64), tri_vt=arr(d.tri_vt, d.n_tri_vt, np.uint64),
               tri_vn=arr(d.tri_vn, d.n_tri_vn, np.uint64))
    lib().mcrt_obj_free(h)
    return out


def vertex_normals(vertices, tri_v, threads=0):
    """mcrt_obj_vertex_normals: Scene::generateVertexNormals (scene.cpp:326-355). -> [n_vertices, 3]."""
    vertices = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
    tri_v = np.ascontiguousarray(tri_v, dtype=np.uint64).reshape(-1, 3)
    out = np.zeros_like(vertices)
    rc = lib().mcrt_obj_vertex_normals(_ptr(vertices), len(vertices), _ptr(tri_v), len(tri_v), int(threads), _ptr(out))
    if rc:
        raise McrtError("mcrt_obj_vertex_normals: triangle index out of range")
    return out


def bvh_build(prim_bounds, scene_bounds, bvh_type, bins_per_axis=0, device=0):
    """mcrt_bvh_build: the reference's BVH (bvh.cpp:13-78) over primitive boxes, built on the GPU.
    -> dict(node_bounds [n,6], node_first_prim, node_prim_count, node_next_sibling, prim_order, gpu_ms, rounds)."""
    prim_bounds = np.ascontiguousarray(prim_bounds, dtype=np.float64).reshape(-1, 6)
    scene_bounds = np.ascontiguousarray(scene_bounds, dtype=np.float64).reshape(6)
    code = BVH_TYPES[bvh_type.lower()] if isinstance(bvh_type, str) else int(bvh_type)
    ctx = C.c_void_p()
    rc = lib().mcrt_init(device, C.byref(ctx))
    if rc:
        raise McrtError(f"mcrt_init({device}) failed: {rc} (no CUDA device? there is no CPU fallback)")
    try:
        h, d, ms = C.c_void_p(), BvhDesc(), C.c_double()
        rc = lib().mcrt_bvh_build(ctx, _ptr(prim_bounds), len(prim_bounds), _ptr(scene_bounds), code, int(bins_per_axis),
                                  C.byref(h), C.byref(d), C.byref(ms))
        if rc:
            raise McrtError(f"mcrt_bvh_build failed ({rc}): {lib().mcrt_last_error(ctx).decode()}")

        def arr(ptr, count, dtype):
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count * np.dtype(dtype).itemsize,)).view(dtype).copy()
        out = dict(node_bounds=arr(d.node_bounds, d.n_nodes * 6, np.float64).reshape(-1, 6),
                   node_first_prim=arr(d.node_first_prim, d.n_nodes, np.uint32),
                   node_prim_count=arr(d.node_prim_count, d.n_nodes, np.uint32),
                   node_next_sibling=arr(d.node_next_sibling, d.n_nodes, np.uint32),
                   prim_order=arr(d.prim_order, d.n_prims, np.uint32),
                   gpu_ms=ms.value, rounds=int(d.build_rounds), kernel_launches=int(d.kernel_launches))
        lib().mcrt_bvh_free(h)
        return out
    finally:
        lib().mcrt_destroy(ctx)


BVH4_NODE_DTYPE = np.dtype([("lo", "<f4", (3, 4)), ("hi", "<f4", (3, 4)), ("child", "<u4", (4,)), ("pad", "<u4", (4,))])


def bvh4_host(scene, max_leaf=0xFFFFFFFF):
    """The 4-wide float-box BVH of the order-free search as mcrt_scene_upload builds it (host only) -> structured array."""
    d = scene.desc()
    h, p, n = C.c_void_p(), C.c_void_p(), C.c_uint32()
    rc = lib().mcrt_bvh4_host(C.byref(d), max_leaf, C.byref(h), C.byref(p), C.byref(n))
    if rc:
        raise McrtError(f"mcrt_bvh4_host failed with {rc}")
    try:
        if n.value == 0:
            return np.zeros(0, BVH4_NODE_DTYPE)
        buf = (C.c_uint8 * (128 * n.value)).from_address(p.value)
        return np.frombuffer(buf, dtype=BVH4_NODE_DTYPE).copy()
    finally:
        lib().mcrt_bvh4_host_free(h)


def shard_rows(height, rank, world_size):
    """Row block of `rank`: contiguous blocks, remainder spread over the first ranks."""
    base, rem = divmod(height, world_size)
    y0 = rank * base + min(rank, rem)
    return y0, y0 + base + (1 if rank < rem else 0)
