// See gpu_integrator.hpp. Compile with -fno-access-control against /root/reference/{source,lib/*}.
#include "gpu_integrator.hpp"

#include <cctype>

#include <fstream>

#include "camera/camera.hpp"
#include "camera/pixel-operators.hpp"
#include "integrator/integrator.hpp"
#include "integrator/photon-mapper/photon-mapper.hpp"
#include "sampling/sampler.hpp"

namespace mcrt_host
{
    void GpuRenderer::check(int rc, const char* what) const
    {
        if (rc != MCRT_OK)
        {
            throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " +
                                     (ctx_ ? mcrt_last_error(ctx_) : "no context"));
        }
    }

    GpuBvh GpuBvh::fromTypeName(std::string type, int bins_per_axis)
    {
        for (auto& c : type) c = (char)std::toupper((unsigned char)c);
        GpuBvh b;
        b.type = type == "QUATERNARY_SAH" ? MCRT_BVH_QUATERNARY_SAH : (type == "BINARY_SAH" ? MCRT_BVH_BINARY_SAH : MCRT_BVH_OCTREE);
        b.bins_per_axis = bins_per_axis;
        return b;
    }

    GpuRenderer::GpuRenderer(const Camera& camera, int device, int precision, const GpuBvh* gpu_bvh)
        : precision_(precision), integrator_kind_(MCRT_INTEGRATOR_PATH), global_seed_(Sampler::global_seed)
    {
        int rc = mcrt_init(device, &ctx_);
        if (rc != MCRT_OK) throw std::runtime_error("mcrt_init failed: no CUDA device (there is no CPU fallback)");

        FlatScene flat;
        flattenScene(camera.integrator->scene, flat);
        if (gpu_bvh)
        {
            // BVH::BVH (bvh.cpp:13-78) on the GPU: same tree as the reference would have built
            if (camera.integrator->scene.bvh) throw std::runtime_error("GpuRenderer: scene already has a CPU-built BVH");
            std::vector<double> prim_bounds;
            double scene_bounds[6];
            primitiveBounds(camera.integrator->scene, prim_bounds, scene_bounds);
            void* handle = nullptr;
            mcrt_bvh_desc bvh{};
            check(mcrt_bvh_build(ctx_, prim_bounds.data(), (uint32_t)(prim_bounds.size() / 6), scene_bounds, gpu_bvh->type,
                                 gpu_bvh->bins_per_axis, &handle, &bvh, &bvh_build_ms_), "mcrt_bvh_build");
            applyBvh(flat, bvh);
            mcrt_bvh_free(handle);
        }
        mcrt_scene_desc desc = flat.desc();
        uint64_t bytes = 0;
        check(mcrt_scene_upload(ctx_, &desc, &bytes), "mcrt_scene_upload");
        h2d_bytes_ += bytes;

        if (auto* pm = dynamic_cast<const PhotonMapper*>(camera.integrator.get()))
        {
            FlatPhotonMap caustic, global;
            flattenPhotonMap(*pm, 0, caustic);
            flattenPhotonMap(*pm, 1, global);
            uint32_t k = 0, dv = 0;
            photonMapParams(*pm, k, dv);
            mcrt_photon_map_desc dc = caustic.desc(), dg = global.desc();
            check(mcrt_photon_upload(ctx_, &dc, &dg, k, dv, &bytes), "mcrt_photon_upload");
            h2d_bytes_ += bytes;
            integrator_kind_ = MCRT_INTEGRATOR_PHOTON;
        }
    }

    GpuRenderer::~GpuRenderer()
    {
        mcrt_destroy(ctx_);
    }

    std::vector<double> GpuRenderer::renderRows(const Camera& camera, uint32_t y0, uint32_t y1)
    {
        // the camera's reconstruction filter; anything but the default box needs the whole frame
        // in one call (the ABI reports MCRT_ERR_UNSUPPORTED otherwise)
        const mcrt_film film = flattenFilm(camera);
        check(mcrt_set_film(ctx_, &film), "mcrt_set_film");
        mcrt_camera cam = flattenCamera(camera);
        std::vector<double> out((size_t)cam.width * (y1 - y0) * 3);
        check(mcrt_render_rows(ctx_, &cam, y0, y1, (uint32_t)camera.sqrtspp, global_seed_, integrator_kind_,
                               precision_, out.data(), &stats_), "mcrt_render_rows");
        return out;
    }

    void GpuRenderer::sampleImage(Camera& camera)
    {
        const uint32_t W = (uint32_t)camera.image.width, H = (uint32_t)camera.image.height;
        std::vector<double> rgb = renderRows(camera, 0, H);
        for (uint32_t y = 0; y < H; y++)
            for (uint32_t x = 0; x < W; x++)
            {
                const double* p = &rgb[((size_t)y * W + x) * 3];
                camera.image(x, y) = glm::dvec3(p[0], p[1], p[2]); // what film.scan(x, y) would return
            }
    }

    void GpuRenderer::emitPhotons(const Camera& camera, uint64_t emissions, double caustic_factor, uint32_t max_photons_per_octree_leaf,
                                  uint32_t k_nearest_photons, bool direct_visualization)
    {
        mcrt_photon_emit_params p{};
        p.emissions = emissions;
        p.caustic_factor = caustic_factor;
        p.max_photons_per_octree_leaf = max_photons_per_octree_leaf;
        p.k_nearest_photons = k_nearest_photons;
        p.direct_visualization = direct_visualization ? 1u : 0u;
        p.global_seed = global_seed_;
        const BoundingBox bb = camera.integrator->scene.BB();   // the octrees' root box (photon-mapper.cpp:181-183)
        for (int c = 0; c < 3; c++) { p.scene_bounds[c] = bb.min[c]; p.scene_bounds[3 + c] = bb.max[c]; }
        uint64_t n_caustic = 0, n_global = 0;
        check(mcrt_photon_emit(ctx_, &p, precision_, &n_caustic, &n_global, &stats_), "mcrt_photon_emit");
        integrator_kind_ = MCRT_INTEGRATOR_PHOTON;
    }

    void GpuRenderer::saveImage(const Camera& camera)
    {
        const Image& img = camera.image;
        typedef glm::dvec3 (*Op)(const glm::dvec3&);
        const Op* op = img.tonemap.target<Op>();
        if (!op) throw std::runtime_error("GpuRenderer::saveImage: unknown tone mapping operator");
        mcrt_image_params prm{};
        prm.plain = img.plain ? 1u : 0u;
        prm.tonemapper = *op == static_cast<Op>(filmicACES) ? MCRT_TONEMAP_ACES : MCRT_TONEMAP_HABLE;
        prm.exposure_scale = img.exposure_scale;
        prm.gain_scale = img.gain_scale;
        std::vector<uint8_t> bytes(img.num_pixels * 3);
        static_assert(sizeof(glm::dvec3) == 3 * sizeof(double), "Image::blob must be packed doubles");
        check(mcrt_image_tonemap(ctx_, reinterpret_cast<const double*>(img.blob.data()), (uint32_t)img.width, (uint32_t)img.height,
                                 &prm, bytes.data(), nullptr, nullptr), "mcrt_image_tonemap");
        // Image::HeaderTGA (image.hpp:37-48): uncompressed 24 bpp true colour, top-left origin
        uint8_t header[18] = {0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 24, 32};
        header[12] = (uint8_t)(img.width & 0xFF); header[13] = (uint8_t)(img.width >> 8);
        header[14] = (uint8_t)(img.height & 0xFF); header[15] = (uint8_t)(img.height >> 8);
        std::ofstream out(camera.savename + ".tga", std::ios::binary);
        out.write(reinterpret_cast<const char*>(header), sizeof(header));
        out.write(reinterpret_cast<const char*>(bytes.data()), (std::streamsize)bytes.size());
    }

    std::vector<double> GpuRenderer::sampleRays(const std::vector<mcrt_ray>& rays, const std::vector<uint32_t>& pixel,
                                                const std::vector<uint32_t>& sample)
    {
        if (rays.size() != pixel.size() || rays.size() != sample.size()) throw std::invalid_argument("sampleRays: size mismatch");
        std::vector<double> out(rays.size() * 3);
        check(mcrt_sample_rays(ctx_, rays.data(), pixel.data(), sample.data(), rays.size(), global_seed_,
                               integrator_kind_, precision_, out.data(), &stats_), "mcrt_sample_rays");
        return out;
    }
}
