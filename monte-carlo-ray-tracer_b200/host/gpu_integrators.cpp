// See gpu_integrators.hpp. Compile with -fno-access-control against /root/reference/{source,lib/*}.
#include "gpu_integrators.hpp"

#include <stdexcept>
#include <string>

#include "camera/camera.hpp"
#include "common/util.hpp"
#include "sampling/sampler.hpp"

#include "exporter.hpp"

namespace
{
    constexpr uint32_t inverseOdd(uint32_t a)   // a * inverseOdd(a) == 1 (mod 2^32)
    {
        uint32_t x = a;
        for (int i = 0; i < 5; i++) x *= 2u - a * x;
        return x;
    }

    // inverse of Sampler::hash (sampler.hpp:77-86): xorshift-15 and odd multiplications are bijections
    uint32_t unhash(uint32_t x)
    {
        x ^= x >> 15; x ^= x >> 30;
        x *= inverseOdd(0xaf723597u);
        x ^= x >> 15; x ^= x >> 30;
        x *= inverseOdd(0xd168aaadu);
        x ^= x >> 15; x ^= x >> 30;
        return x;
    }
}

namespace mcrt_host
{
    void currentSamplerPixelAndSample(uint32_t& pixel, uint32_t& sample)
    {
        // base_seed = hashCombine(global_seed, hash(start_seed)) = gs ^ (hash(start_seed) + 0x9e3779b9 + (gs << 6) + (gs >> 2))
        const uint32_t gs = Sampler::global_seed;
        const uint32_t hashed = (Sampler::base_seed ^ gs) - 0x9e3779b9u - (gs << 6) - (gs >> 2);
        pixel = unhash(hashed);
        sample = Sobol::reverseBits(Sampler::bit_reversed_index);
    }
}

void GpuPathTracer::check(int rc, const char* what) const
{
    if (rc != MCRT_OK)
        throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + (ctx_ ? mcrt_last_error(ctx_) : "no context"));
}

GpuPathTracer::GpuPathTracer(const nlohmann::json& j, int device, int precision)
    : Integrator(j), precision_(precision)       // reference code: scene load + BVH build on the CPU
{
    if (mcrt_init(device, &ctx_) != MCRT_OK) throw std::runtime_error("mcrt_init failed: no CUDA device (there is no CPU fallback)");
    mcrt_host::FlatScene flat;
    mcrt_host::flattenScene(scene, flat);
    mcrt_scene_desc desc = flat.desc();
    uint64_t bytes = 0;
    check(mcrt_scene_upload(ctx_, &desc, &bytes), "mcrt_scene_upload");
}

GpuPathTracer::~GpuPathTracer()
{
    mcrt_destroy(ctx_);
}

glm::dvec3 GpuPathTracer::sampleRay(Ray ray)
{
    if (ray.depth != 0 || Sampler::sequence != 0u)
        throw std::logic_error("GpuPathTracer::sampleRay continues no half-traced path: pass the camera ray of a fresh sample");
    uint32_t pixel, sample;
    mcrt_host::currentSamplerPixelAndSample(pixel, sample);
    mcrt_ray r;
    for (int c = 0; c < 3; c++) { r.origin[c] = ray.start[c]; r.direction[c] = ray.direction[c]; }
    double rgb[3] = { 0.0, 0.0, 0.0 };
    std::lock_guard<std::mutex> lock(mutex_);
    check(mcrt_sample_rays(ctx_, &r, &pixel, &sample, 1, Sampler::global_seed, kind_, precision_, rgb, &stats_), "mcrt_sample_rays");
    return glm::dvec3(rgb[0], rgb[1], rgb[2]);
}

std::vector<double> GpuPathTracer::sampleRows(const Camera& camera, uint32_t y0, uint32_t y1)
{
    const mcrt_film film = mcrt_host::flattenFilm(camera);
    mcrt_camera cam = mcrt_host::flattenCamera(camera);
    std::vector<double> out((size_t)cam.width * (y1 - y0) * 3);
    std::lock_guard<std::mutex> lock(mutex_);
    check(mcrt_set_film(ctx_, &film), "mcrt_set_film");
    check(mcrt_render_rows(ctx_, &cam, y0, y1, (uint32_t)camera.sqrtspp, Sampler::global_seed, kind_, precision_, out.data(), &stats_),
          "mcrt_render_rows");
    return out;
}

GpuPhotonMapper::GpuPhotonMapper(const nlohmann::json& j, int device, int precision) : GpuPathTracer(j, device, precision)
{
    // PhotonMapper::PhotonMapper's first pass (photon-mapper.cpp:24-277) on the GPU: emission + both octrees
    const nlohmann::json& pm = j.at("photon_map");
    mcrt_photon_emit_params p{};
    p.emissions = pm.at("emissions").get<uint64_t>();
    p.caustic_factor = pm.at("caustic_factor").get<double>();
    p.max_photons_per_octree_leaf = getOptional(pm, "max_photons_per_octree_leaf", 200u);
    p.k_nearest_photons = getOptional(pm, "k_nearest_photons", 50u);
    p.direct_visualization = getOptional(pm, "direct_visualization", false) ? 1u : 0u;
    p.global_seed = Sampler::global_seed;
    const BoundingBox bb = scene.BB();            // the octrees' root box (photon-mapper.cpp:181-183)
    for (int c = 0; c < 3; c++) { p.scene_bounds[c] = bb.min[c]; p.scene_bounds[3 + c] = bb.max[c]; }
    check(mcrt_photon_emit(ctx_, &p, precision_, &n_caustic_, &n_global_, &stats_), "mcrt_photon_emit");
    kind_ = MCRT_INTEGRATOR_PHOTON;
}
