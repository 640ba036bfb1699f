// Flattens the reference's *built* Scene / BVH / Material / Camera / photon-map objects into the
// POD arrays of include/mcrt_abi.h. Host-side C++20, compiled against the reference's own headers
// (-I/root/reference/source ... -fno-access-control: the fields read here are private in
// source/bvh/bvh.hpp:105-108, source/surface/surface.hpp:68-115, source/material/material.hpp:51-54).
// Nothing here re-implements reference logic: it only copies the state the reference's loader,
// BVH builders and photon pass produced, so load-time quirks are inherited unchanged.
#pragma once

#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "mcrt_abi.h"

class Scene;
class Camera;
class PhotonMapper;
namespace Surface { class Base; }

namespace mcrt_host
{
    struct FlatScene
    {
        std::vector<double> node_bounds;
        std::vector<uint32_t> node_first_prim, node_prim_count, node_next_sibling;
        std::vector<uint8_t> prim_type;
        std::vector<uint32_t> prim_index, prim_material;
        std::vector<double> prim_area;
        std::vector<double> tri_v0, tri_v1, tri_v2, tri_e1, tri_e2, tri_normal;
        std::vector<int32_t> tri_vn_index;
        std::vector<double> vertex_normals;
        std::vector<double> sphere_origin_radius;
        std::vector<double> quadric_Q, quadric_G, quadric_bounds;
        std::vector<mcrt_material> materials;
        std::vector<uint32_t> light_prim;
        std::vector<double> light_cdf;
        double scene_ior = 1.0;
        // inputs of BVH::BVH (bvh.cpp:13-15), for mcrt_bvh_build: position of every ordered primitive
        // in Scene::surfaces, and Scene::BB()
        std::vector<uint32_t> prim_original;
        double scene_bounds[6] = {0, 0, 0, 0, 0, 0};

        // ordered-primitive index of every Surface::Base object (pointer identity is what the
        // reference compares in integrator.cpp:70,101)
        std::unordered_map<const Surface::Base*, uint32_t> prim_of_surface;

        mcrt_scene_desc desc() const;
    };

    struct FlatPhotonMap
    {
        std::vector<double> octant_bounds;
        std::vector<uint64_t> octant_start, octant_count;
        std::vector<uint32_t> octant_next_sibling;
        std::vector<uint8_t> octant_leaf;
        std::vector<float> photons;

        mcrt_photon_map_desc desc() const;
    };

    void flattenScene(const Scene& scene, FlatScene& out);
    // Inputs of BVH::BVH (bvh.cpp:13-15) for mcrt_bvh_build: Surface::Base::BB() of Scene::surfaces in
    // order ([n][6]) and Scene::BB().
    void primitiveBounds(const Scene& scene, std::vector<double>& prim_bounds, double scene_bounds[6]);
    // Re-orders a scene flattened WITHOUT a hierarchy (Scene built from a JSON without "bvh") into
    // the order of a tree built by mcrt_bvh_build and attaches its node arrays.
    void applyBvh(FlatScene& flat, const mcrt_bvh_desc& bvh);
    mcrt_camera flattenCamera(const Camera& camera);
    // the camera's Film (filter, radius, cache size; source/camera/film.cpp:19-59)
    mcrt_film flattenFilm(const Camera& camera);
    // which: 0 caustic_map, 1 global_map
    void flattenPhotonMap(const PhotonMapper& pm, int which, FlatPhotonMap& out);
    void photonMapParams(const PhotonMapper& pm, uint32_t& k_nearest, uint32_t& direct_visualization);

    // "Scene pack": the flattened arrays in one little-endian file so that a scene built here can
    // be rendered on a machine that has neither the reference sources nor its OBJ assets.
    // Layout: magic "MCRTPK01", u32 n_entries, then n_entries × {char name[32]; u32 dtype;
    // u32 elem_size; u64 count; u64 offset}, then 64-byte aligned payloads.
    // dtype: 0 u8, 1 u32, 2 i32, 3 u64, 4 f32, 5 f64, 6 raw struct.
    struct PackWriter
    {
        struct Entry { std::string name; uint32_t dtype, elem_size; uint64_t count; const void* data; };
        std::vector<Entry> entries;
        std::vector<std::vector<uint8_t>> owned;

        template <class T> void add(const std::string& name, uint32_t dtype, const std::vector<T>& v)
        {
            entries.push_back({ name, dtype, (uint32_t)sizeof(T), v.size(), v.data() });
        }
        void addScalars(const std::string& name, const std::vector<double>& v);
        void addScalarsU32(const std::string& name, const std::vector<uint32_t>& v);
        bool write(const std::string& path) const;
    };

    void addSceneToPack(PackWriter& w, const FlatScene& s);
    void addCameraToPack(PackWriter& w, const std::string& prefix, const mcrt_camera& c, uint32_t sqrtspp);
    void addFilmToPack(PackWriter& w, const std::string& prefix, const mcrt_film& f);
    void addPhotonMapToPack(PackWriter& w, const std::string& prefix, const FlatPhotonMap& m);
}
