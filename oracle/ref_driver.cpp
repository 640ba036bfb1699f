// ORACLE / TEST INFRASTRUCTURE — never linked into the product library (libmcrt_b200.so).
//
// Headless C driver around the UNMODIFIED reference sources in /root/reference/source, compiled
// where they lie into oracle/_ref/libmcrt_ref.so by oracle/build_ref.py. It gives tests, bench.py's
// cpu_baseline / --impl reference leg and the scene-pack tool access to:
//   * the reference's Camera/PathTracer/PhotonMapper with a pinned sampler seed (seed_pin.hpp),
//     raw float64 Film::scan output instead of the 8-bit TGA (source/camera/image.cpp:37-51),
//     a steady_clock timer around the worker threads only (the reference joins a 1-s polling
//     printer thread first, source/camera/camera.cpp:131), and ray counters obtained with
//     ld --wrap on Scene::intersect / Integrator::sampleDirect (no source edits);
//   * per-function hooks (Scene::intersect, Integrator::sampleRay, Sampler, Fresnel, GGX, Material,
//     LinearOctree::knnSearch) used as known-answer generators;
//   * the exporter (monte-carlo-ray-tracer_b200/host/exporter.cpp) to write scene packs.
// Compiled with -fno-access-control: Camera::samplePixel, Camera::integrator, Sampler::global_seed
// etc. are private in the reference.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <unistd.h>
#include <thread>
#include <vector>

#include <nlohmann/json.hpp>

#include "obj_adapter.hpp"
#include "bvh/bvh.hpp"
#include "camera/camera.hpp"
#include "common/option.hpp"
#include "integrator/integrator.hpp"
#include "integrator/path-tracer/path-tracer.hpp"
#include "integrator/photon-mapper/photon-mapper.hpp"
#include "material/fresnel.hpp"
#include "material/ggx.hpp"
#include "material/material.hpp"
#include "ray/interaction.hpp"
#include "sampling/sampler.hpp"
#include "sampling/sampling.hpp"
#include "scene/scene.hpp"
#include "surface/surface.hpp"

#include "exporter.hpp"

namespace mcrt_oracle
{
    unsigned pinnedSeed()
    {
        static unsigned seed = []() -> unsigned
        {
            const char* e = std::getenv("MCRT_ORACLE_SEED");
            return e ? (unsigned)std::strtoul(e, nullptr, 0) : 0x12345678u;
        }();
        return seed;
    }
}

// ---------------------------------------------------------------- ray counters (ld --wrap)
static thread_local uint64_t tl_intersect_calls = 0;
static thread_local uint64_t tl_shadow_calls = 0;
static thread_local int tl_in_sample_direct = 0;

extern "C" Intersection __real__ZNK5Scene9intersectERK3Ray(const Scene*, const Ray&);
extern "C" Intersection __wrap__ZNK5Scene9intersectERK3Ray(const Scene* self, const Ray& ray)
{
    tl_intersect_calls++;
    if (tl_in_sample_direct) tl_shadow_calls++;
    return __real__ZNK5Scene9intersectERK3Ray(self, ray);
}

extern "C" glm::dvec3 __real__ZNK10Integrator12sampleDirectERK11InteractionRNS_11LightSampleE(
    const Integrator*, const Interaction&, Integrator::LightSample&);
extern "C" glm::dvec3 __wrap__ZNK10Integrator12sampleDirectERK11InteractionRNS_11LightSampleE(
    const Integrator* self, const Interaction& ia, Integrator::LightSample& ls)
{
    tl_in_sample_direct++;
    glm::dvec3 r = __real__ZNK10Integrator12sampleDirectERK11InteractionRNS_11LightSampleE(self, ia, ls);
    tl_in_sample_direct--;
    return r;
}

namespace
{
    struct Handle
    {
        std::unique_ptr<Camera> camera;
        mcrt_host::FlatScene flat;
        bool flat_valid = false;
        bool photon_map = false;
        nlohmann::json scene_json;
        int camera_idx = 0;
        std::unique_ptr<BVH> extra_bvh;   // ref_bvh_build
    };

    void setError(char* err, size_t errlen, const std::string& msg)
    {
        if (err && errlen)
        {
            std::strncpy(err, msg.c_str(), errlen - 1);
            err[errlen - 1] = 0;
        }
    }

    // silence the reference's progress chatter
    struct CoutSilencer
    {
        std::streambuf* old;
        std::ostringstream sink;
        CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) { }
        ~CoutSilencer() { std::cout.rdbuf(old); }
    };

    const mcrt_host::FlatScene& flatOf(Handle* h)
    {
        if (!h->flat_valid)
        {
            mcrt_host::flattenScene(h->camera->integrator->scene, h->flat);
            h->flat_valid = true;
        }
        return h->flat;
    }

    // Camera::samplePixel body for one sample (source/camera/camera.cpp:66-96), sampler already set.
    Ray cameraRay(const Camera& c, size_t x, size_t y, glm::dvec2& px)
    {
        double pixel_size = c.sensor_width / c.image.width;
        glm::dvec2 half_dim = glm::dvec2(c.image.width, c.image.height) * 0.5;
        auto u = Sampler::get<Dim::PIXEL, 2>();
        px = glm::dvec2(x + u[0], y + u[1]);
        glm::dvec2 local = pixel_size * (half_dim - px);
        glm::dvec3 direction = glm::normalize(c.forward * c.focal_length + c.left * local.x + c.up * local.y);
        Ray ray(c.eye, direction, c.integrator->scene.ior);
        if (c.thin_lens)
        {
            auto ul = Sampler::get<Dim::LENS, 2>();
            glm::dvec2 aperture_sample = Sampling::uniformDisk(ul[0], ul[1]) * c.aperture_radius;
            glm::dvec3 focus_point = ray(c.focus_distance / glm::dot(ray.direction, c.forward));
            glm::dvec3 start = c.eye + c.left * aperture_sample.x + c.up * aperture_sample.y;
            ray = Ray(start, glm::normalize(focus_point - start), c.integrator->scene.ior);
        }
        return ray;
    }
}

extern "C"
{

// overrides_json keys (all optional): width, height, sqrtspp, bvh_type ("octree"|"binary_sah"|
// "quaternary_sah"|"none"), bins_per_axis, emissions, caustic_factor, k_nearest_photons,
// max_photons_per_octree_leaf, num_render_threads (used by the photon pass), film (the camera's "film" object).
void* ref_open(const char* scenes_dir, const char* scene_file, const char* overrides_json,
               int camera_idx, int photon_map, char* err, size_t errlen)
{
    try
    {
        CoutSilencer quiet;
        std::filesystem::path dir(scenes_dir);
        Scene::path = dir;
        std::ifstream in(dir / scene_file);
        if (!in) throw std::runtime_error(std::string("cannot open scene ") + scene_file);
        nlohmann::json j;
        in >> j;

        nlohmann::json o = nlohmann::json::object();
        if (overrides_json && overrides_json[0]) o = nlohmann::json::parse(overrides_json);

        auto& cam = j.at("cameras").at(camera_idx);
        if (o.contains("width")) cam["image"]["width"] = o["width"];
        if (o.contains("height")) cam["image"]["height"] = o["height"];
        if (o.contains("sqrtspp")) cam["sqrtspp"] = o["sqrtspp"];
        if (o.contains("film")) cam["film"] = o["film"];   // {"filter": ..., "radius": ..., "cache_size": ...}
        if (o.contains("bvh_type"))
        {
            std::string t = o["bvh_type"];
            if (t == "none") j.erase("bvh");
            else
            {
                if (!j.contains("bvh")) j["bvh"] = nlohmann::json::object();
                j["bvh"]["type"] = t;
                j["bvh"].erase("bins_per_axis");
            }
        }
        if (o.contains("bins_per_axis")) j["bvh"]["bins_per_axis"] = o["bins_per_axis"];
        if (o.contains("num_render_threads")) j["num_render_threads"] = o["num_render_threads"];
        for (const char* k : { "emissions", "caustic_factor", "k_nearest_photons", "max_photons_per_octree_leaf" })
        {
            if (o.contains(k)) j["photon_map"][k] = o[k];
        }

        auto h = std::make_unique<Handle>();
        h->photon_map = photon_map != 0;
        h->camera_idx = camera_idx;
        Option option(dir / scene_file, "", camera_idx, photon_map != 0);
        h->camera = std::make_unique<Camera>(j, option);
        h->scene_json = std::move(j);
        return h.release();
    }
    catch (const std::exception& e)
    {
        setError(err, errlen, e.what());
        return nullptr;
    }
}

void ref_close(void* handle)
{
    delete static_cast<Handle*>(handle);
}

// The seed the reference keeps in a private `inline static const` (sampler.hpp:58). It has a dynamic
// initialiser, so it lives in writable storage; tests re-pin it per case.
void ref_set_seed(uint32_t seed)
{
    const_cast<uint32_t&>(Sampler::global_seed) = seed;
}

uint32_t ref_get_seed()
{
    return Sampler::global_seed;
}

int ref_info(void* handle, uint32_t* width, uint32_t* height, uint32_t* sqrtspp, uint32_t* n_prims,
             uint32_t* n_nodes, uint32_t* n_lights)
{
    Handle* h = static_cast<Handle*>(handle);
    const auto& f = flatOf(h);
    *width = (uint32_t)h->camera->image.width;
    *height = (uint32_t)h->camera->image.height;
    *sqrtspp = (uint32_t)h->camera->sqrtspp;
    *n_prims = (uint32_t)f.prim_type.size();
    *n_nodes = (uint32_t)f.node_first_prim.size();
    *n_lights = (uint32_t)f.light_prim.size();
    return 0;
}

int ref_hardware_threads()
{
    return (int)std::thread::hardware_concurrency();
}

// Camera::sampleImage (camera.cpp:101-145) for rows [y0,y1): same 32×32 buckets and the reference's
// own Camera::samplePixel, but buckets are handed out in raster order through an atomic counter and
// no printer thread runs. out_rgb[(y-y0)*W+x][3] = Film::scan.
int ref_render(void* handle, int threads, uint32_t y0, uint32_t y1, double* out_rgb,
               double* seconds, uint64_t* total_rays, uint64_t* shadow_rays)
{
    Handle* h = static_cast<Handle*>(handle);
    Camera& c = *h->camera;
    const size_t W = c.image.width, H = c.image.height;
    if (y1 > H) y1 = (uint32_t)H;
    if (y0 >= y1) return -1;
    if (threads < 1) threads = (int)std::thread::hardware_concurrency();

    {
        // fresh accumulators, same filter as Camera::Camera chose (camera.cpp:34-37)
        const auto& cj = h->scene_json.at("cameras").at(h->camera_idx);
        if (cj.find("film") != cj.end()) c.film = Film(W, H, cj.at("film"));
        else c.film = Film(W, H);
    }

    struct Bucket { size_t x0, y0, x1, y1; };
    std::vector<Bucket> buckets;
    const size_t bs = 32;
    for (size_t y = y0; y < y1; y += bs)
        for (size_t x = 0; x < W; x += bs)
            buckets.push_back({ x, y, std::min(x + bs, W), std::min(y + bs, (size_t)y1) });

    std::atomic<size_t> next(0);
    std::atomic<uint64_t> rays(0), shadows(0);

    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
    {
        pool.emplace_back([&]()
        {
            tl_intersect_calls = 0; tl_shadow_calls = 0;
            size_t i;
            while ((i = next.fetch_add(1)) < buckets.size())
            {
                const Bucket& b = buckets[i];
                for (size_t y = b.y0; y < b.y1; y++)
                    for (size_t x = b.x0; x < b.x1; x++)
                        c.samplePixel(x, y);
            }
            rays += tl_intersect_calls;
            shadows += tl_shadow_calls;
        });
    }
    for (auto& t : pool) t.join();
    auto t1 = std::chrono::steady_clock::now();

    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (total_rays) *total_rays = rays.load();
    if (shadow_rays) *shadow_rays = shadows.load();

    if (out_rgb)
    {
        for (size_t y = y0; y < y1; y++)
            for (size_t x = 0; x < W; x++)
            {
                glm::dvec3 v = c.film.scan(x, y);
                double* o = out_rgb + ((y - y0) * W + x) * 3;
                o[0] = v.x; o[1] = v.y; o[2] = v.z;
            }
    }
    return 0;
}

// One (pixel, sample) at a time: camera ray (out_ray6, may be null) and Integrator::sampleRay.
int ref_sample_pixels(void* handle, const uint32_t* pixel, const uint32_t* sample, size_t n,
                      double* out_rgb, double* out_ray6)
{
    Handle* h = static_cast<Handle*>(handle);
    Camera& c = *h->camera;
    const size_t W = c.image.width;
    for (size_t i = 0; i < n; i++)
    {
        Sampler::initiate(pixel[i]);
        Sampler::setIndex(sample[i]);
        glm::dvec2 px;
        Ray ray = cameraRay(c, pixel[i] % W, pixel[i] / W, px);
        if (out_ray6)
        {
            for (int k = 0; k < 3; k++) { out_ray6[i * 6 + k] = ray.start[k]; out_ray6[i * 6 + 3 + k] = ray.direction[k]; }
        }
        if (out_rgb)
        {
            glm::dvec3 v = c.integrator->sampleRay(ray);
            for (int k = 0; k < 3; k++) out_rgb[i * 3 + k] = v[k];
        }
    }
    return 0;
}

// Integrator::sampleRay on caller-supplied rays (medium = scene ior).
int ref_sample_rays(void* handle, const double* rays6, const uint32_t* pixel, const uint32_t* sample,
                    size_t n, double* out_rgb)
{
    Handle* h = static_cast<Handle*>(handle);
    Camera& c = *h->camera;
    for (size_t i = 0; i < n; i++)
    {
        Sampler::initiate(pixel[i]);
        Sampler::setIndex(sample[i]);
        Ray ray(glm::dvec3(rays6[i * 6], rays6[i * 6 + 1], rays6[i * 6 + 2]),
                glm::dvec3(rays6[i * 6 + 3], rays6[i * 6 + 4], rays6[i * 6 + 5]), c.integrator->scene.ior);
        glm::dvec3 v = c.integrator->sampleRay(ray);
        for (int k = 0; k < 3; k++) out_rgb[i * 3 + k] = v[k];
    }
    return 0;
}

// Scene::intersect on caller-supplied rays; prim = ordered-primitive index or 0xFFFFFFFF.
int ref_trace(void* handle, const double* rays6, size_t n, double* out_t, uint32_t* out_prim,
              double* out_uv, uint8_t* out_interpolate)
{
    Handle* h = static_cast<Handle*>(handle);
    const Scene& scene = h->camera->integrator->scene;
    const auto& f = flatOf(h);
    for (size_t i = 0; i < n; i++)
    {
        Ray ray(glm::dvec3(rays6[i * 6], rays6[i * 6 + 1], rays6[i * 6 + 2]),
                glm::dvec3(rays6[i * 6 + 3], rays6[i * 6 + 4], rays6[i * 6 + 5]), scene.ior);
        Intersection is = scene.intersect(ray);
        if (is)
        {
            out_t[i] = is.t;
            out_prim[i] = f.prim_of_surface.at(is.surface.get());
            out_uv[i * 2] = is.interpolate ? is.uv.x : 0.0;
            out_uv[i * 2 + 1] = is.interpolate ? is.uv.y : 0.0;
            out_interpolate[i] = is.interpolate;
        }
        else
        {
            out_t[i] = is.t;
            out_prim[i] = 0xFFFFFFFFu;
            out_uv[i * 2] = out_uv[i * 2 + 1] = 0.0;
            out_interpolate[i] = 0;
        }
    }
    return 0;
}

// Sampler streams (sampler.hpp:19-52): raw 32-bit values of get<0,7>() after n_shuffles shuffles.
int ref_sampler_stream(const uint32_t* pixel, const uint32_t* sample, size_t n, uint32_t n_shuffles,
                       uint32_t* out_u32x7)
{
    for (size_t i = 0; i < n; i++)
    {
        Sampler::initiate(pixel[i]);
        Sampler::setIndex(sample[i]);
        for (uint32_t s = 0; s < n_shuffles; s++) Sampler::shuffle();
        auto u = Sampler::get<0, 7>();
        for (int d = 0; d < 7; d++) out_u32x7[i * 7 + d] = (uint32_t)(u[d] * 0x1p32);
    }
    return 0;
}

// LinearOctree<Photon>::knnSearch; results sorted by distance2 ascending, photon = 8 floats.
int ref_knn(void* handle, int which, const double* points, size_t n, uint32_t k, float* out_photons,
            double* out_dist2, uint32_t* out_count)
{
    Handle* h = static_cast<Handle*>(handle);
    auto* pm = dynamic_cast<PhotonMapper*>(h->camera->integrator.get());
    if (!pm) return -1;
    const LinearOctree<Photon>& map = which == 0 ? pm->caustic_map : pm->global_map;
    PriorityQueue<SearchResult<Photon>> result;
    for (size_t i = 0; i < n; i++)
    {
        map.knnSearch(glm::dvec3(points[i * 3], points[i * 3 + 1], points[i * 3 + 2]), k, result);
        std::vector<SearchResult<Photon>> v(result.begin(), result.end());
        std::sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return a.distance2 < b.distance2; });
        out_count[i] = (uint32_t)v.size();
        for (size_t r = 0; r < v.size(); r++)
        {
            out_dist2[i * k + r] = v[r].distance2;
            float* o = out_photons + (i * k + r) * 8;
            for (int c = 0; c < 3; c++) { o[c] = v[r].data.flux_[c]; o[3 + c] = v[r].data.position_[c]; }
            o[6] = v[r].data.phi; o[7] = v[r].data.theta;
        }
    }
    return 0;
}

// Scene::parseOBJ (scene.cpp:238-324) on any file, through the opened scene object; results are kept
// in the handle and copied out by ref_obj_copy. Returns 0, or -1 if the reference threw.
namespace
{
    struct ParsedObj
    {
        std::vector<glm::dvec3> v, n;
        std::vector<std::vector<size_t>> tv, tvt, tvn;
    } g_obj;
}

int ref_obj_parse(void* handle, const char* path, uint64_t counts[5], double* seconds)
{
    Handle* h = static_cast<Handle*>(handle);
    const Scene& scene = h->camera->integrator->scene;
    g_obj = ParsedObj();
    try
    {
        CoutSilencer quiet;
        auto t0 = std::chrono::steady_clock::now();
        scene.parseOBJ(path, g_obj.v, g_obj.n, g_obj.tv, g_obj.tvt, g_obj.tvn);
        auto t1 = std::chrono::steady_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    }
    catch (const std::exception&)
    {
        return -1;
    }
    counts[0] = g_obj.v.size(); counts[1] = g_obj.n.size(); counts[2] = g_obj.tv.size(); counts[3] = g_obj.tvt.size(); counts[4] = g_obj.tvn.size();
    return 0;
}

void ref_obj_copy(double* v, double* n, uint64_t* tv, uint64_t* tvt, uint64_t* tvn)
{
    for (size_t i = 0; i < g_obj.v.size(); i++) for (int c = 0; c < 3; c++) v[3 * i + c] = g_obj.v[i][c];
    for (size_t i = 0; i < g_obj.n.size(); i++) for (int c = 0; c < 3; c++) n[3 * i + c] = g_obj.n[i][c];
    auto copy = [](const std::vector<std::vector<size_t>>& src, uint64_t* dst) { for (size_t i = 0; i < src.size(); i++) for (int c = 0; c < 3; c++) dst[3 * i + c] = src[i][c]; };
    copy(g_obj.tv, tv); copy(g_obj.tvt, tvt); copy(g_obj.tvn, tvn);
}

// The drop-in bodies of host/obj_adapter.hpp (what a maintainer puts into Scene::parseOBJ /
// generateVertexNormals) against the reference's own: 1 = every container equal.
int ref_obj_adapter_check(void* handle, const char* path, int with_normals)
{
    Handle* h = static_cast<Handle*>(handle);
    const Scene& scene = h->camera->integrator->scene;
    ParsedObj a, b;
    try
    {
        CoutSilencer quiet;
        scene.parseOBJ(path, a.v, a.n, a.tv, a.tvt, a.tvn);
        mcrt_host::parseOBJ(path, b.v, b.n, b.tv, b.tvt, b.tvn);
    }
    catch (const std::exception&)
    {
        return -1;
    }
    if (!(a.v == b.v && a.n == b.n && a.tv == b.tv && a.tvt == b.tvt && a.tvn == b.tvn)) return 0;
    if (with_normals)
    {
        std::vector<glm::dvec3> na, nb;
        try { scene.generateVertexNormals(na, a.v, a.tv); mcrt_host::generateVertexNormals(nb, b.v, b.tv); }
        catch (const std::exception&) { return -2; }
        if (na.size() != nb.size()) return 0;
        for (size_t i = 0; i < na.size(); i++)
            for (int c = 0; c < 3; c++)
                if (!(na[i][c] == nb[i][c] || (na[i][c] != na[i][c] && nb[i][c] != nb[i][c]))) return 0;
    }
    return 1;
}

// Scene::generateVertexNormals (scene.cpp:326-355)
int ref_vertex_normals(void* handle, const double* vertices, uint64_t n_vertices, const uint64_t* tri_v, uint64_t n_tris, double* out, double* seconds)
{
    Handle* h = static_cast<Handle*>(handle);
    const Scene& scene = h->camera->integrator->scene;
    std::vector<glm::dvec3> v(n_vertices), n;
    for (size_t i = 0; i < n_vertices; i++) v[i] = glm::dvec3(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]);
    std::vector<std::vector<size_t>> t(n_tris, std::vector<size_t>(3));
    for (size_t i = 0; i < n_tris; i++) for (int c = 0; c < 3; c++) t[i][c] = tri_v[3 * i + c];
    try
    {
        auto t0 = std::chrono::steady_clock::now();
        scene.generateVertexNormals(n, v, t);
        auto t1 = std::chrono::steady_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    }
    catch (const std::exception&)
    {
        return -1;
    }
    for (size_t i = 0; i < n.size(); i++) for (int c = 0; c < 3; c++) out[3 * i + c] = n[i][c];
    return 0;
}

// Image::save (image.cpp:37-51) on caller pixels: `image_json` is a camera's "image" object; the bytes
// are what the reference writes after the TGA header. Also returns getExposure()*exposure_scale
// and getGain()*gain_scale.
int ref_image_save(const char* image_json, const double* rgb, uint8_t* out_bgr, double* exposure_factor, double* gain_factor)
{
    try
    {
        Image img(nlohmann::json::parse(image_json));
        for (size_t i = 0; i < img.num_pixels; i++) img.blob[i] = glm::dvec3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
        const double e = img.plain ? 1.0 : img.getExposure() * img.exposure_scale;
        const double g = img.plain ? 1.0 : img.getGain(e) * img.gain_scale;
        if (exposure_factor) *exposure_factor = e;
        if (gain_factor) *gain_factor = g;
        char name[] = "/tmp/mcrt_ref_image_XXXXXX";
        int fd = mkstemp(name);
        if (fd < 0) return -2;
        close(fd);
        img.save(name);                       // writes name + ".tga"
        std::string tga = std::string(name) + ".tga";
        std::ifstream in(tga, std::ios::binary);
        in.seekg(18);
        in.read(reinterpret_cast<char*>(out_bgr), (std::streamsize)(img.num_pixels * 3));
        const bool ok = (size_t)in.gcount() == img.num_pixels * 3;
        in.close();
        std::remove(tga.c_str()); std::remove(name);
        return ok ? 0 : -3;
    }
    catch (const std::exception&)
    {
        return -1;
    }
}

// BVH::BVH on the opened scene's surfaces (bvh.cpp:13-78), timed: the oracle of mcrt_bvh_build.
// type: "octree" | "binary_sah" | "quaternary_sah"; bins <= 0 = the reference's default.
int ref_bvh_build(void* handle, const char* type, int bins, double* seconds, uint32_t* n_nodes)
{
    Handle* h = static_cast<Handle*>(handle);
    const Scene& scene = h->camera->integrator->scene;
    nlohmann::json j = nlohmann::json::object();
    j["type"] = std::string(type);
    if (bins > 0) j["bins_per_axis"] = bins;
    CoutSilencer quiet;
    auto t0 = std::chrono::steady_clock::now();
    h->extra_bvh = std::make_unique<BVH>(scene.BB(), scene.surfaces, j);
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (n_nodes) *n_nodes = (uint32_t)h->extra_bvh->linear_tree.size();
    return 0;
}

// arrays of the BVH built by ref_bvh_build: LinearNode fields + ordered_surfaces as indices into Scene::surfaces
int ref_bvh_arrays(void* handle, double* node_bounds, uint32_t* first, uint32_t* count, uint32_t* next, uint32_t* order)
{
    Handle* h = static_cast<Handle*>(handle);
    if (!h->extra_bvh) return -1;
    const BVH& b = *h->extra_bvh;
    const Scene& scene = h->camera->integrator->scene;
    for (size_t i = 0; i < b.linear_tree.size(); i++)
    {
        const auto& n = b.linear_tree[i];
        for (int c = 0; c < 3; c++) { node_bounds[6 * i + c] = n.BB.min[c]; node_bounds[6 * i + 3 + c] = n.BB.max[c]; }
        first[i] = n.start_surface; count[i] = n.num_surfaces; next[i] = n.next_sibling;
    }
    std::unordered_map<const Surface::Base*, uint32_t> original;
    for (size_t i = 0; i < scene.surfaces.size(); i++) original.emplace(scene.surfaces[i].get(), (uint32_t)i);
    for (size_t i = 0; i < b.ordered_surfaces.size(); i++) order[i] = original.at(b.ordered_surfaces[i].get());
    return 0;
}

// Surface::Base::BB() of Scene::surfaces in order + Scene::BB()
int ref_prim_bounds(void* handle, double* bounds, double* scene_bounds)
{
    Handle* h = static_cast<Handle*>(handle);
    const Scene& scene = h->camera->integrator->scene;
    for (size_t i = 0; i < scene.surfaces.size(); i++)
    {
        const BoundingBox bb = scene.surfaces[i]->BB();
        for (int c = 0; c < 3; c++) { bounds[6 * i + c] = bb.min[c]; bounds[6 * i + 3 + c] = bb.max[c]; }
    }
    const BoundingBox sb = scene.BB();
    for (int c = 0; c < 3; c++) { scene_bounds[c] = sb.min[c]; scene_bounds[3 + c] = sb.max[c]; }
    return 0;
}

// Scene pack: scene + selected camera (+ photon maps when opened with photon_map).
int ref_export_pack(void* handle, const char* path)
{
    Handle* h = static_cast<Handle*>(handle);
    const auto& f = flatOf(h);
    mcrt_host::PackWriter w;
    mcrt_host::addSceneToPack(w, f);
    mcrt_host::addCameraToPack(w, "camera", mcrt_host::flattenCamera(*h->camera), (uint32_t)h->camera->sqrtspp);
    mcrt_host::addFilmToPack(w, "camera", mcrt_host::flattenFilm(*h->camera));
    {
        // the "bvh" object the reference built this scene's BVH from (bvh.cpp:24-56): {has, type, bins_per_axis}
        uint32_t has = 0, type = MCRT_BVH_OCTREE, bins = 0;
        if (h->scene_json.contains("bvh"))
        {
            has = 1;
            const auto& bj = h->scene_json.at("bvh");
            std::string t = bj.value("type", std::string("OCTREE"));
            std::transform(t.begin(), t.end(), t.begin(), ::toupper);
            if (t == "QUATERNARY_SAH") { type = MCRT_BVH_QUATERNARY_SAH; bins = bj.value("bins_per_axis", 8); }
            else if (t == "BINARY_SAH") { type = MCRT_BVH_BINARY_SAH; bins = bj.value("bins_per_axis", 16); }
        }
        w.addScalarsU32("bvh_params", { has, type, bins });
    }
    mcrt_host::FlatPhotonMap caustic, global;
    if (auto* pm = dynamic_cast<PhotonMapper*>(h->camera->integrator.get()))
    {
        mcrt_host::flattenPhotonMap(*pm, 0, caustic);
        mcrt_host::flattenPhotonMap(*pm, 1, global);
        mcrt_host::addPhotonMapToPack(w, "caustic", caustic);
        mcrt_host::addPhotonMapToPack(w, "global", global);
        uint32_t k, dv;
        mcrt_host::photonMapParams(*pm, k, dv);
        w.addScalarsU32("photon_params", { k, dv });
        // parameters of the photon pass (scene JSON "photon_map") + Scene::BB(), for mcrt_photon_emit
        const auto& pmj = h->scene_json.at("photon_map");
        const BoundingBox bb = h->camera->integrator->scene.BB();
        w.addScalars("photon_emit_params", { (double)pmj.at("emissions").get<size_t>(), pmj.at("caustic_factor").get<double>(),
                                             (double)pm->max_node_data, bb.min.x, bb.min.y, bb.min.z, bb.max.x, bb.max.y, bb.max.z });
    }
    return w.write(path) ? 0 : -1;
}

// ------------------------------------------------------------ per-function known-answer hooks
double ref_fresnel_dielectric(double n1, double n2, double cos_theta)
{
    return Fresnel::dielectric(n1, n2, cos_theta);
}

void ref_fresnel_conductor(double n1, const double* real3, const double* imag3, double cos_theta, double* out3)
{
    ComplexIOR ior(glm::dvec3(real3[0], real3[1], real3[2]), glm::dvec3(imag3[0], imag3[1], imag3[2]));
    glm::dvec3 r = Fresnel::conductor(n1, &ior, cos_theta);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

double ref_ggx_reflection(const double* wi, const double* wo, double alpha, double* pdf)
{
    return GGX::reflection(glm::dvec3(wi[0], wi[1], wi[2]), glm::dvec3(wo[0], wo[1], wo[2]), glm::dvec2(alpha), *pdf);
}

double ref_ggx_transmission(const double* wi, const double* wo, double n1, double n2, double alpha, double* pdf)
{
    return GGX::transmission(glm::dvec3(wi[0], wi[1], wi[2]), glm::dvec3(wo[0], wo[1], wo[2]), n1, n2, glm::dvec2(alpha), *pdf);
}

void ref_ggx_visible_microfacet(double u, double v, const double* wo, double alpha, double* out3)
{
    glm::dvec3 m = GGX::visibleMicrofacet(u, v, glm::dvec3(wo[0], wo[1], wo[2]), glm::dvec2(alpha));
    out3[0] = m.x; out3[1] = m.y; out3[2] = m.z;
}

} // extern "C"
