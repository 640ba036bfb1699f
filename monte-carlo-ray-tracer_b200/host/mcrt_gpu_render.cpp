// Headless stand-in for the reference's main.cpp (source/main.cpp:10-61) that renders through the
// GPU path: same scene directory / JSON / camera index / integrator choice, same Camera object,
// same Image::save — only Camera::sampleImage is replaced by GpuRenderer::sampleImage.
// usage: mcrt_gpu_render <scenes_dir> <scene.json> [camera_idx] [photon_map 0|1] [f64|f32] [gpu_bvh 0|1] [gpu_image 0|1] [gpu_photons 0|1]
// gpu_bvh = 1: the scene's "bvh" object is taken out of the JSON (the reference then builds no
// hierarchy) and the same tree is built by mcrt_bvh_build. gpu_image = 1: Image::save's exposure /
// tone mapping / gamma run on the GPU too (mcrt_image_tonemap) instead of camera.saveImage().
// gpu_photons = 1 (with photon_map = 1): the camera is built with a PathTracer, so the reference runs no
// CPU photon pass, and emission + octrees run on the GPU (mcrt_photon_emit).
#include <chrono>
#include <fstream>
#include <iostream>

#include <nlohmann/json.hpp>

#include "camera/camera.hpp"
#include "common/option.hpp"
#include "scene/scene.hpp"

#include "sampling/sampler.hpp"

#include "gpu_integrator.hpp"

int main(int argc, char* argv[])
{
    if (argc < 3)
    {
        std::cerr << "usage: mcrt_gpu_render <scenes_dir> <scene.json> [camera_idx] [photon_map] [f64|f32] [gpu_bvh] [gpu_image] [gpu_photons]\n";
        return 2;
    }
    try
    {
        // the reference seeds its sampler from std::random_device (sampler.hpp:58); MCRT_SEED makes a run repeatable
        if (const char* e = std::getenv("MCRT_SEED")) const_cast<uint32_t&>(Sampler::global_seed) = (uint32_t)std::strtoul(e, nullptr, 0);
        std::filesystem::path dir(argv[1]);
        Scene::path = dir;
        int camera_idx = argc > 3 ? std::atoi(argv[3]) : 0;
        bool photon_map = argc > 4 && std::atoi(argv[4]) != 0;
        int precision = (argc > 5 && std::string(argv[5]) == "f32") ? MCRT_PRECISION_F32 : MCRT_PRECISION_F64;

        std::ifstream in(dir / argv[2]);
        nlohmann::json j;
        in >> j;
        const bool gpu_photons = photon_map && argc > 8 && std::atoi(argv[8]) != 0;
        // (the reference's CPU photon pass traverses the reference's own BVH, so the two only combine with gpu_photons)
        const bool gpu_bvh = argc > 6 && std::atoi(argv[6]) != 0 && j.contains("bvh") && (!photon_map || gpu_photons);
        mcrt_host::GpuBvh bvh;
        if (gpu_bvh)
        {
            const auto& b = j.at("bvh");
            bvh = mcrt_host::GpuBvh::fromTypeName(b.value("type", std::string("OCTREE")), b.value("bins_per_axis", 0));
            j.erase("bvh");
        }
        Option option(dir / argv[2], "", camera_idx, photon_map && !gpu_photons);
        Camera camera(j, option);                    // reference: scene load, (BVH build,) photon pass

        mcrt_host::GpuRenderer gpu(camera, 0, precision, gpu_bvh ? &bvh : nullptr);
        if (gpu_bvh) std::cout << "bvh built on the GPU in " << gpu.bvhBuildMs() << " ms" << std::endl;
        if (gpu_photons)
        {
            const auto& pm = j.at("photon_map");
            gpu.emitPhotons(camera, pm.at("emissions").get<uint64_t>(), pm.at("caustic_factor").get<double>(),
                            pm.value("max_photons_per_octree_leaf", 200u), pm.value("k_nearest_photons", 50u),
                            pm.value("direct_visualization", false));
            std::cout << "photon pass on the GPU: " << gpu.lastStats().gpu_ms_total << " ms emission + "
                      << gpu.lastStats().gpu_ms_knn << " ms octrees" << std::endl;
        }
        auto t0 = std::chrono::steady_clock::now();
        gpu.sampleImage(camera);                     // GPU: the hot path
        auto t1 = std::chrono::steady_clock::now();
        if (argc > 7 && std::atoi(argv[7]) != 0) gpu.saveImage(camera);   // GPU: exposure, tonemap, gamma; same TGA
        else camera.saveImage();                     // reference: exposure, tonemap, TGA

        const mcrt_stats& st = gpu.lastStats();
        double rays = double(st.extension_rays + st.shadow_rays);
        std::cout << "paths " << st.paths << " rays " << (uint64_t)rays << " gpu_ms " << st.gpu_ms_total
                  << " Mray/s " << rays / st.gpu_ms_total / 1e3 << " wall_ms "
                  << std::chrono::duration<double, std::milli>(t1 - t0).count() << std::endl;
    }
    catch (const std::exception& e)
    {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
