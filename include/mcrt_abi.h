/*
 * mcrt_abi.h — C ABI of the B200-native path-tracing integrator.
 *
 * Drop-in boundary for the ray/BVH/BSDF hot path of linusmossberg/monte-carlo-ray-tracer.
 * The reference has no FFI; the interface this replaces is the C++ virtual
 *     glm::dvec3 Integrator::sampleRay(Ray)            (source/integrator/integrator.hpp:20)
 * and its only caller, the per-pixel loop in
 *     Camera::samplePixel / Camera::sampleImage        (source/camera/camera.cpp:66-145).
 * Because one scalar ray per virtual call cannot feed a GPU, the batch boundary sits at
 * the body of Camera::sampleImage ("render these rows of this camera"), plus batched
 * sampleRay / Scene::intersect entry points with the reference's per-call semantics.
 *
 * Conventions
 *   - plain C, no torch / C++ types; every pointer is a HOST pointer unless the name ends
 *     in _dev; the caller owns host buffers, the context owns device memory.
 *   - return 0 on success, a negative mcrt_status on failure; mcrt_last_error() gives text.
 *   - one context per GPU, driven from one host thread at a time.
 *   - all scene arrays are the *built* state of the reference's Scene/BVH/Material objects
 *     in float64 (the reference computes in double), flattened by the exporter
 *     (monte-carlo-ray-tracer_b200/host/exporter.cpp). Layouts for the device (FP64 parity
 *     arrays, FP32 wide-node arrays) are derived inside mcrt_scene_upload.
 */
#ifndef MCRT_ABI_H
#define MCRT_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCRT_ABI_VERSION 1

typedef enum mcrt_status {
    MCRT_OK = 0,
    MCRT_ERR_INVALID = -1,   /* bad argument / inconsistent scene description        */
    MCRT_ERR_CUDA = -2,      /* CUDA runtime error (text in mcrt_last_error)          */
    MCRT_ERR_NO_SCENE = -3,  /* render/trace called before mcrt_scene_upload          */
    MCRT_ERR_UNSUPPORTED = -4, /* feature outside the hot path (e.g. non-box film)     */
    MCRT_ERR_NO_PHOTONS = -5 /* photon-mapped render without mcrt_photon_upload       */
} mcrt_status;

/* Primitive type tags: Surface::{Triangle,Sphere,Quadric} (source/surface/surface.hpp:54-116). */
enum { MCRT_PRIM_TRIANGLE = 0, MCRT_PRIM_SPHERE = 1, MCRT_PRIM_QUADRIC = 2 };

/* Integrator kinds: PathTracer / PhotonMapper (source/camera/camera.cpp:22-29). */
enum { MCRT_INTEGRATOR_PATH = 0, MCRT_INTEGRATOR_PHOTON = 1 };

/* Arithmetic of the device path.
 *   MCRT_PRECISION_F64: parity mode — double, operation order of the reference, no FMA
 *                       contraction, reference's best-first traversal order.
 *   MCRT_PRECISION_F32: fast mode — float, wide-node stack traversal, scale-aware offsets. */
enum { MCRT_PRECISION_F64 = 0, MCRT_PRECISION_F32 = 1 };

/* Material record: the built state of class Material (source/material/material.hpp:37-54)
 * including the derived flags/constants of Material::computeProperties (material.cpp:97-111). */
typedef struct mcrt_material {
    double reflectance[3];
    double specular_reflectance[3];
    double transmittance[3];
    double emittance[3];          /* radiosity after Scene::generateEmissives (scene.cpp:202) */
    double roughness, specular_roughness, ior, transparency;
    double complex_ior_real[3], complex_ior_imag[3];
    double A, B;                  /* Oren–Nayar constants (material.cpp:106-108)              */
    double a[2];                  /* GGX alpha (material.cpp:110)                             */
    uint32_t has_complex_ior, perfect_mirror;
    uint32_t rough, rough_specular, opaque, emissive, dirac_delta;
    uint32_t _pad;
} mcrt_material;

/* Flattened Scene + BVH (source/scene/scene.hpp, source/bvh/bvh.hpp:68-108). */
typedef struct mcrt_scene_desc {
    uint32_t abi_version;             /* MCRT_ABI_VERSION */
    /* BVH::linear_tree in depth-first order; n_nodes == 0 means "no bvh object" and the
     * linear-scan branch of Scene::intersect (scene.cpp:159-173) is taken. */
    uint32_t n_nodes;
    const double* node_bounds;        /* [n_nodes][6]  min.xyz, max.xyz                       */
    const uint32_t* node_first_prim;  /* [n_nodes]     LinearNode::start_surface              */
    const uint32_t* node_prim_count;  /* [n_nodes]     LinearNode::num_surfaces (0 = inner)   */
    const uint32_t* node_next_sibling;/* [n_nodes]     0 = none                               */
    /* BVH::ordered_surfaces (or Scene::surfaces without a BVH). */
    uint32_t n_prims;
    const uint8_t* prim_type;         /* [n_prims] MCRT_PRIM_*                                */
    const uint32_t* prim_index;       /* [n_prims] index into the per-type arrays             */
    const uint32_t* prim_material;    /* [n_prims] index into materials                       */
    const double* prim_area;          /* [n_prims] Surface::Base::area_                       */
    /* triangles (source/surface/surface.hpp:92-96) */
    uint32_t n_tris;
    const double* tri_v0;             /* [n_tris][3] */
    const double* tri_v1;
    const double* tri_v2;
    const double* tri_e1;             /* stored, not recomputed: keeps the reference's rounding */
    const double* tri_e2;
    const double* tri_normal;         /* face normal_ */
    const int32_t* tri_vn_index;      /* [n_tris] index into vertex_normals or -1             */
    uint32_t n_vertex_normals;
    const double* vertex_normals;     /* [n_vertex_normals][9]  n0,n1,n2                      */
    /* spheres (surface.hpp:68-69) */
    uint32_t n_spheres;
    const double* sphere_origin_radius; /* [n_spheres][4] */
    /* quadrics (surface.hpp:114-115, clip box = Base::BB_) */
    uint32_t n_quadrics;
    const double* quadric_Q;          /* [n_quadrics][16] column-major dmat4                  */
    const double* quadric_G;          /* [n_quadrics][12] column-major dmat4x3                */
    const double* quadric_bounds;     /* [n_quadrics][6]                                      */
    /* materials: one entry per distinct Material object */
    uint32_t n_materials;
    const mcrt_material* materials;
    /* Scene::emissives / cumulative_emissives_importance (scene.cpp:178-209) */
    uint32_t n_lights;
    const uint32_t* light_prim;       /* [n_lights] ordered-primitive index of each emissive  */
    const double* light_cdf;          /* [n_lights]                                           */
    double scene_ior;                 /* Scene::ior */
} mcrt_scene_desc;

/* Camera state after Camera::Camera (source/camera/camera.cpp:20-64). */
typedef struct mcrt_camera {
    double eye[3], forward[3], left[3], up[3];
    double focal_length, sensor_width, aperture_radius, focus_distance;
    uint32_t width, height;
    uint32_t thin_lens;
    uint32_t _pad;
} mcrt_camera;

/* Ray as the reference's Ray(start, direction, medium_ior) (source/ray/ray.cpp:13-14). */
typedef struct mcrt_ray {
    double origin[3];
    double direction[3];
} mcrt_ray;

/* Result of Scene::intersect (source/ray/intersection.hpp:9-23). prim == 0xFFFFFFFF: miss. */
typedef struct mcrt_hit {
    double t, u, v;
    uint32_t prim;
    uint32_t interpolate;
} mcrt_hit;

/* LinearOctree<Photon> (source/octree/linear-octree.hpp:19-29, photon.hpp:35-37). */
typedef struct mcrt_photon_map_desc {
    uint32_t n_octants;
    const double* octant_bounds;        /* [n_octants][6]                                     */
    const uint64_t* octant_start;       /* start_data                                         */
    const uint64_t* octant_count;       /* contained_data                                     */
    const uint32_t* octant_next_sibling;/* 0xFFFFFFFF = none                                  */
    const uint8_t* octant_leaf;
    uint64_t n_photons;
    const float* photons;               /* [n_photons][8] flux.xyz, pos.xyz, phi, theta       */
} mcrt_photon_map_desc;

typedef struct mcrt_stats {
    uint64_t paths;            /* camera paths started                                        */
    uint64_t extension_rays;   /* closest-hit queries for path extension (incl. primary)      */
    uint64_t shadow_rays;      /* closest-hit queries for next-event estimation               */
    uint64_t box_tests;        /* AABB slab tests executed by the traversal kernels           */
    uint64_t prim_tests;       /* primitive intersection tests executed                       */
    uint64_t knn_queries;      /* photon-map k-NN queries                                     */
    uint64_t wavefront_iterations;
    uint64_t kernel_launches;
    uint64_t ior_stack_overflows;
    uint32_t max_depth;
    uint32_t _pad;
    double gpu_ms_total;       /* CUDA-event time of the render, first launch → last          */
    /* per-stage sums of CUDA-event times; filled only with option "stage_timing" = 1 */
    double gpu_ms_generate, gpu_ms_extend, gpu_ms_shade, gpu_ms_shadow, gpu_ms_knn;
    uint64_t extend_launches, shadow_launches;
    uint64_t shadow_box_tests, shadow_prim_tests; /* k_shadow's share of box_tests / prim_tests */
    /* k_extend warp-tail diagnostic: sum over rays of (box+prim tests) / sum over warps of
     * 32*max over the warp's rays = the lane utilisation lost to uneven ray lengths alone */
    uint64_t extend_work_sum, extend_work_warpmax; /* builds with -DMCRT_TAIL_DIAGNOSTIC only, else 0 */
    /* parity mode: closest-hit queries whose two nearest candidates lay within rounding distance of each
     * other and were therefore re-traced in the reference's own visiting order (csrc/bvh4.cuh) */
    uint64_t replayed_rays;
} mcrt_stats;

typedef struct mcrt_ctx mcrt_ctx;

int mcrt_abi_version(void);

/* Create a context on CUDA device `device`. */
int mcrt_init(int device, mcrt_ctx** out_ctx);
void mcrt_destroy(mcrt_ctx* ctx);
const char* mcrt_last_error(const mcrt_ctx* ctx);

/* Replaces the Scene held by Integrator (integrator.hpp:25): copies the description to the
 * device and derives both device layouts. Returns bytes copied host→device in *h2d_bytes. */
int mcrt_scene_upload(mcrt_ctx* ctx, const mcrt_scene_desc* scene, uint64_t* h2d_bytes);

/* Replaces PhotonMapper's caustic_map/global_map members (photon-mapper.hpp:27-28). */
int mcrt_photon_upload(mcrt_ctx* ctx, const mcrt_photon_map_desc* caustic_map,
                       const mcrt_photon_map_desc* global_map, uint32_t k_nearest,
                       uint32_t direct_visualization, uint64_t* h2d_bytes);

/* Film reconstruction filter, the "film" object of a camera in the scene JSON
 * (source/camera/film.cpp:19-59). Default (never set): box, radius 0.5, no cache. */
enum { MCRT_FILM_BOX = 0, MCRT_FILM_MITCHELL_NETRAVALI = 1, MCRT_FILM_CATMULL_ROM = 2, MCRT_FILM_B_SPLINE = 3,
       MCRT_FILM_HERMITE = 4, MCRT_FILM_GAUSSIAN = 5, MCRT_FILM_LANCZOS = 6 };
typedef struct mcrt_film {
    uint32_t filter;        /* MCRT_FILM_* */
    uint32_t cache_size;    /* 0: evaluate the filter function, else nearest-neighbour lookup table */
    double radius;          /* <= 0: the filter's default radius (film.cpp:32-43) */
} mcrt_film;

/* SURVEY.md §8f-4 ("next"): Film::deposit with the reference's reconstruction filters for the
 * following renders of this context (source/camera/film.cpp:61-113, filter.hpp). With a filter other
 * than the default box, samples splat into neighbouring pixels, so mcrt_render_rows* must then
 * cover the whole frame (rows cannot be sharded without halo exchange). NULL restores the default. */
int mcrt_set_film(mcrt_ctx* ctx, const mcrt_film* film);

/* Parameters of the photon pass, the "photon_map" object of the scene JSON
 * (source/integrator/photon-mapper/photon-mapper.cpp:28-38). */
typedef struct mcrt_photon_emit_params {
    uint64_t emissions;                    /* "emissions" (before the caustic_factor scaling)  */
    double caustic_factor;
    uint32_t max_photons_per_octree_leaf;
    uint32_t k_nearest_photons;
    uint32_t direct_visualization;
    uint32_t global_seed;
    double scene_bounds[6];                /* Scene::BB() = root box of both octrees            */
} mcrt_photon_emit_params;

/* SURVEY.md §8f-1 ("next"): replaces the first pass of PhotonMapper::PhotonMapper
 * (photon-mapper.cpp:24-223) — photon emission + tracing (emitPhoton, :225-277) on the GPU, then
 * Octree<Photon> construction + LinearOctree::compact (octree.cpp:34-81, linear-octree.cpp:201-244)
 * — and installs the two maps in the context exactly as mcrt_photon_upload would. The photon
 * multiset and the octree structure equal the reference's (photons inside one leaf may be stored
 * in a different order: the reference's order depends on its thread schedule). The octrees are
 * built on the device from the emission buffers (the photons never visit the host); in `stats`,
 * gpu_ms_total is the emission wavefront and gpu_ms_knn the octree construction. */
int mcrt_photon_emit(mcrt_ctx* ctx, const mcrt_photon_emit_params* params, int precision,
                     uint64_t* n_caustic, uint64_t* n_global, mcrt_stats* stats);

/* The photon pass in pieces, for sharding it over GPUs (emission sharded by ranges of the
 * reference's EmissionWork index space, photon-mapper.cpp:61-78, photons all-gathered, every rank builds the
 * same octrees):
 *   mcrt_photon_emit_total   size of the emission index space (sum over lights of their emission counts)
 *   mcrt_photon_emit_range   emits work items [work_first, work_first + work_count); the photons stay in device
 *                            buffers of the context - 8 floats per photon {flux.xyz, pos.xyz, phi, theta} - whose
 *                            addresses are returned (valid until the next emission / mcrt_destroy)
 *   mcrt_photon_build_dev    builds both octrees from device photon arrays (this rank's, or the concatenation
 *                            of all ranks') and installs them like mcrt_photon_upload
 * mcrt_photon_emit == total, range(0, total), build. */
int mcrt_photon_emit_total(mcrt_ctx* ctx, const mcrt_photon_emit_params* params, uint64_t* total_emissions);
int mcrt_photon_emit_range(mcrt_ctx* ctx, const mcrt_photon_emit_params* params, int precision,
                           uint64_t work_first, uint64_t work_count, const float** caustic_dev, uint64_t* n_caustic,
                           const float** global_dev, uint64_t* n_global, mcrt_stats* stats);
int mcrt_photon_build_dev(mcrt_ctx* ctx, const mcrt_photon_emit_params* params, const float* caustic_dev,
                          uint64_t n_caustic, const float* global_dev, uint64_t n_global, double* build_ms);

/* Host view of the maps built by mcrt_photon_emit (which: 0 caustic, 1 global). The pointers stay
 * valid until the next mcrt_photon_emit / mcrt_destroy. */
int mcrt_photon_download(mcrt_ctx* ctx, int which, mcrt_photon_map_desc* out);

/* The octree construction step of mcrt_photon_emit alone, on caller photons: Octree<Photon>
 * insertion + LinearOctree::compact (octree.cpp:34-81, linear-octree.cpp:201-244) on the GPU.
 * photons: HOST, [n][8] floats {flux.xyz, pos.xyz, phi, theta} (copied to the device); the result
 * (same octants as the reference's LinearOctree; photons of a leaf in input order) is copied back:
 * *out points into host memory owned by *handle, release it with mcrt_octree_free. */
int mcrt_octree_build(mcrt_ctx* ctx, const float* photons, uint64_t n, uint32_t max_photons_per_octree_leaf,
                      const double* scene_bounds6, void** handle, mcrt_photon_map_desc* out, double* gpu_ms);
void mcrt_octree_free(void* handle);

/* SURVEY.md §8f-3 ("next"): scene ingest. Scene::parseOBJ (source/scene/scene.cpp:238-324) as a
 * parallel mmap-based reader with identical results: "v" / "vn" lines and the first three corners
 * of every "f" line, zero-based (idx - 1 in size_t arithmetic). tri_vt / tri_vn hold only the faces
 * whose three corners all carry that index, exactly as the reference's separate lists do. Host
 * code, no GPU involved. *out points into memory owned by *handle (mcrt_obj_free). A missing file
 * or a negative index (the reference prints / throws) returns MCRT_ERR_INVALID with the message. */
typedef struct mcrt_obj_mesh {
    uint64_t n_vertices, n_normals, n_tri_v, n_tri_vt, n_tri_vn;
    const double* vertices;     /* [n_vertices][3] */
    const double* normals;      /* [n_normals][3] */
    const uint64_t* tri_v;      /* [n_tri_v][3] */
    const uint64_t* tri_vt;     /* [n_tri_vt][3] */
    const uint64_t* tri_vn;     /* [n_tri_vn][3] */
} mcrt_obj_mesh;
int mcrt_obj_load(const char* path, int threads, void** handle, mcrt_obj_mesh* out, char* err, size_t errlen);
void mcrt_obj_free(void* handle);
/* Scene::generateVertexNormals (scene.cpp:326-355): normalised sum of face normal x area x corner
 * angle over the incident triangles, added in triangle order (bit-identical sums), parallel over
 * vertices. out_normals: [n_vertices][3]. */
int mcrt_obj_vertex_normals(const double* vertices, uint64_t n_vertices, const uint64_t* tri_v, uint64_t n_triangles,
                            int threads, double* out_normals);

/* SURVEY.md §8f-4 ("next", image half): Image::save (source/camera/image.cpp:37-51) without the file:
 * auto exposure (getExposure, image.cpp:63-73: histogram median -> 0.5), tone-mapping operator
 * (pixel-operators.cpp:7-51), auto gain (getGain, image.cpp:78-88: 99th percentile -> 0.99), sRGB
 * gamma (srgb.hpp:54-62) and truncation to bytes in B,G,R order. `params` = the camera's "image"
 * object (image.cpp:10-35). rgb: [height][width][3] float64 as mcrt_render_rows returns it;
 * out_bgr: [height][width][3] bytes = the reference's .tga after its 18-byte header.
 * mcrt_image_tonemap takes HOST buffers, mcrt_image_tonemap_dev DEVICE pointers (e.g. the
 * framebuffer of mcrt_render_rows_dev, so that 6 MB of bytes leave the GPU instead of 50 MB). */
enum { MCRT_TONEMAP_HABLE = 0, MCRT_TONEMAP_ACES = 1, MCRT_TONEMAP_LINEAR = 2 };
typedef struct mcrt_image_params {
    uint32_t plain;                 /* "plain": no tone mapping, no auto exposure / gain */
    uint32_t tonemapper;            /* MCRT_TONEMAP_HABLE (default) | MCRT_TONEMAP_ACES */
    double exposure_scale;          /* 2^exposure_compensation, Image::exposure_scale (image.cpp:21) */
    double gain_scale;              /* 2^gain_compensation, Image::gain_scale (image.cpp:22) */
} mcrt_image_params;
int mcrt_image_tonemap(mcrt_ctx* ctx, const double* rgb, uint32_t width, uint32_t height, const mcrt_image_params* params,
                       uint8_t* out_bgr, double* exposure_factor, double* gain_factor);
int mcrt_image_tonemap_dev(mcrt_ctx* ctx, const double* rgb_dev, uint32_t width, uint32_t height,
                           const mcrt_image_params* params, uint8_t* out_bgr_dev, double* exposure_factor, double* gain_factor);

/* SURVEY.md §8f-2 ("next"): BVH construction on the GPU. Replaces BVH::BVH (source/bvh/bvh.cpp:13-78):
 * the binned-SAH builders recursiveBuildBinarySAH / recursiveBuildQuaternarySAH (bvh.cpp:165-432),
 * the octree-derived hierarchy (bvh.cpp:130-163, octree.cpp:34-81), arbitrarySplit and compact
 * (bvh.cpp:434-474). The result is the reference's tree node for node (same boxes, same
 * depth-first order, same ordered_surfaces), because every decision of those builders depends
 * only on counts and min/max unions.
 *   prim_bounds   [n_prims][6] = Surface::Base::BB() {min.xyz, max.xyz} in Scene::surfaces order
 *   scene_bounds  Scene::BB() (the root box, bvh.cpp:20)
 *   type          the "bvh" object's "type" (bvh.cpp:24-56), bins_per_axis its "bins_per_axis"
 *                 (<= 0: the reference's default, 16 binary / 8 quaternary)
 * *out points into memory owned by *handle (release with mcrt_bvh_free): node arrays in the layout
 * mcrt_scene_desc takes, and prim_order[i] = index into the caller's primitives of ordered
 * primitive i. gpu_ms (optional): device time of the build. */
enum { MCRT_BVH_OCTREE = 0, MCRT_BVH_BINARY_SAH = 1, MCRT_BVH_QUATERNARY_SAH = 2 };
typedef struct mcrt_bvh_desc {
    uint32_t n_nodes, n_prims;
    const double* node_bounds;            /* [n_nodes][6] */
    const uint32_t* node_first_prim;      /* LinearNode::start_surface */
    const uint32_t* node_prim_count;      /* LinearNode::num_surfaces (0 = inner node) */
    const uint32_t* node_next_sibling;    /* LinearNode::next_sibling */
    const uint32_t* prim_order;           /* [n_prims] */
    uint32_t build_rounds, kernel_launches;
} mcrt_bvh_desc;
int mcrt_bvh_build(mcrt_ctx* ctx, const double* prim_bounds, uint32_t n_prims, const double* scene_bounds6, int type,
                   int bins_per_axis, void** handle, mcrt_bvh_desc* out, double* gpu_ms);
void mcrt_bvh_free(void* handle);

/* Replaces Camera::sampleImage (camera.cpp:101-145) for rows [y0, y1) with the default box
 * film (film.cpp:13-17): out_rgb[(y-y0)*W+x][3] = mean over sqrtspp² samples of
 * Integrator::sampleRay, clamped at 0 (film.cpp:112). Sample s of pixel p uses
 * Sampler::initiate(p) / setIndex(s) with `global_seed` (sampler.hpp:30-44,58).
 * out_rgb is a HOST buffer of (y1-y0)*W*3 doubles. */
int mcrt_render_rows(mcrt_ctx* ctx, const mcrt_camera* camera, uint32_t y0, uint32_t y1,
                     uint32_t sqrtspp, uint32_t global_seed, int integrator_kind,
                     int precision, double* out_rgb, mcrt_stats* stats);

/* Same, but the framebuffer stays resident in HBM (device pointer, doubles); used by the
 * multi-GPU all-gather and by the HBM-resident throughput measurement. */
int mcrt_render_rows_dev(mcrt_ctx* ctx, const mcrt_camera* camera, uint32_t y0, uint32_t y1,
                         uint32_t sqrtspp, uint32_t global_seed, int integrator_kind,
                         int precision, double* out_rgb_dev, mcrt_stats* stats);

/* Interleaved row sharding for multi-GPU renders: renders image rows
 * y_first + k*y_step for k in [0, n_rows) into out_rgb_dev[k*W + x][3] (device pointer). Every
 * rank gets statistically identical rows, so per-rank cost is balanced. */
int mcrt_render_rows_strided_dev(mcrt_ctx* ctx, const mcrt_camera* camera, uint32_t y_first,
                                 uint32_t y_step, uint32_t n_rows, uint32_t sqrtspp,
                                 uint32_t global_seed, int integrator_kind, int precision,
                                 double* out_rgb_dev, mcrt_stats* stats);

/* Row-sharded render with the frame exchange fused into the film resolve (replaces the NCCL
 * all-gather of the framebuffer): the rows y_first + k*y_step this rank renders are written straight into
 * frames[0..n_frames) - the FULL-frame buffers [height][width][3] of every rank, float32 ("float3
 * framebuffer") or float64 - at their final position. frames[] are device pointers valid on this device:
 * the rank's own buffer from mcrt_frame_alloc and the peers' buffers mapped with mcrt_frame_open (CUDA IPC;
 * the stores cross NVLink). The caller synchronises the ranks (barrier) before reading a frame and before
 * the next render overwrites it. */
int mcrt_render_rows_strided_peers(mcrt_ctx* ctx, const mcrt_camera* camera, uint32_t y_first,
                                   uint32_t y_step, uint32_t n_rows, uint32_t sqrtspp,
                                   uint32_t global_seed, int integrator_kind, int precision,
                                   void* const* frames, uint32_t n_frames, int frame_is_float32,
                                   mcrt_stats* stats);
/* Reconstruction filters across row shards (film.cpp:61-79): with a filter set by mcrt_set_film a sample splats
 * into neighbouring rows, so a rank that renders rows y_first + k*y_step accumulates into whole-frame buffers and
 * returns them UNRESOLVED: rgb_sum_dev[height*width][3] and weight_sum_dev[height*width] (device, float64). The
 * host adds them over the ranks (an all-reduce) and calls mcrt_film_resolve_dev, Film::Splat::get (film.cpp:106-113). */
int mcrt_render_film_sums_strided_dev(mcrt_ctx* ctx, const mcrt_camera* camera, uint32_t y_first, uint32_t y_step,
                                      uint32_t n_rows, uint32_t sqrtspp, uint32_t global_seed, int integrator_kind,
                                      int precision, double* rgb_sum_dev, double* weight_sum_dev, mcrt_stats* stats);
int mcrt_film_resolve_dev(mcrt_ctx* ctx, const double* rgb_sum_dev, const double* weight_sum_dev, uint64_t n_pixels,
                          double* out_rgb_dev);

/* A device buffer that other processes on the node can map: *dev_ptr (zero-filled) and its 64-byte CUDA IPC
 * handle, to be sent to the peers by whatever channel the host uses (torch.distributed in this repository). */
int mcrt_frame_alloc(mcrt_ctx* ctx, uint64_t bytes, void** dev_ptr, unsigned char ipc_handle[64]);
int mcrt_frame_open(mcrt_ctx* ctx, const unsigned char ipc_handle[64], void** dev_ptr);
int mcrt_frame_close(mcrt_ctx* ctx, void* peer_ptr);
int mcrt_frame_free(mcrt_ctx* ctx, void* dev_ptr);

/* Test hook (host only, no CUDA call): the 4-wide float-box BVH mcrt_scene_upload derives from the scene's tree for the
 * order-free closest-hit search (csrc/bvh4.cuh). nodes128: n_nodes records of 128 bytes {float lo[3][4], hi[3][4];
 * uint32 child[4], pad[4]}; child = 0 empty | inner node index | 0x80000000 | first_prim << 8 | count. max_leaf: 0 keeps the
 * reference's leaves, n cuts larger leaves into runs of n, 0xFFFFFFFF = the upload's own rule. */
int mcrt_bvh4_host(const mcrt_scene_desc* scene, uint32_t max_leaf, void** handle, const void** nodes128, uint32_t* n_nodes);
void mcrt_bvh4_host_free(void* handle);

/* Measured FP64 issue rate of this GPU (independent DFMA chains on every SM), thread-instructions per second:
 * the denominator of bench.py's FP64 roofline for the float64 kernels. */
int mcrt_fp64_peak(mcrt_ctx* ctx, double* dfma_per_second);

/* Batched Scene::intersect (scene.cpp:151-176): closest hit per ray. `medium_ior` is not
 * needed by the query. Host buffers. */
int mcrt_trace_closest(mcrt_ctx* ctx, const mcrt_ray* rays, size_t n, int precision,
                       mcrt_hit* hits, mcrt_stats* stats);

/* Batched Integrator::sampleRay (integrator.hpp:20): ray i is traced as sample `sample[i]`
 * of pixel `pixel[i]` (the sampler state the reference keeps thread_local), medium = scene
 * ior; out_rgb[i][3] receives the radiance estimate. Host buffers. */
int mcrt_sample_rays(mcrt_ctx* ctx, const mcrt_ray* rays, const uint32_t* pixel,
                     const uint32_t* sample, size_t n, uint32_t global_seed,
                     int integrator_kind, int precision, double* out_rgb, mcrt_stats* stats);

/* Test hook for the Owen-scrambled Sobol sampler (sampler.hpp:13-91): for each i, after
 * initiate(pixel[i]), setIndex(sample[i]) and `n_shuffles` calls of shuffle(), writes the
 * seven dimensions get<0,7>() as raw uint32 (before the 2^-32 scaling). Host buffers. */
int mcrt_sampler_stream(mcrt_ctx* ctx, const uint32_t* pixel, const uint32_t* sample, size_t n,
                        uint32_t n_shuffles, uint32_t global_seed, uint32_t* out_u32x7);

/* Batched LinearOctree<Photon>::knnSearch (linear-octree.cpp:24-117) on the uploaded map
 * (which: 0 caustic, 1 global). out_index[i][k] photon indices (ordered_data order, unsorted,
 * 0xFFFFFFFF padded), out_dist2[i][k]; out_count[i] results found. Host buffers. */
int mcrt_knn_search(mcrt_ctx* ctx, int which, const double* points_xyz, size_t n,
                    uint32_t* out_index, double* out_dist2, uint32_t* out_count,
                    mcrt_stats* stats);

/* Tunables (pool size etc.); unknown keys return MCRT_ERR_INVALID. */
int mcrt_set_option(mcrt_ctx* ctx, const char* key, double value);

#ifdef __cplusplus
}
#endif
#endif /* MCRT_ABI_H */
