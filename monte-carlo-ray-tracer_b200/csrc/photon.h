// Photon-mapped rendering support (PhotonMapper::sampleRay + LinearOctree::knnSearch on the device).
// Host-side interface used by abi.cu; kernels live in photon_f64.cu / photon_f32.cu.
#pragma once

#include <string>
#include <vector>

#include "mcrt_abi.h"
#include "integrator.cuh"

namespace mcrt
{
    struct DevicePhotonMap
    {
        const double* octant_bounds = nullptr;   // [n][6]
        const uint2* octant_range32 = nullptr;   // unused placeholder
        const unsigned long long* octant_start = nullptr;
        const unsigned long long* octant_count = nullptr;
        const uint32_t* octant_next = nullptr;
        const uint8_t* octant_leaf = nullptr;
        const float4* photons = nullptr;         // 2 float4 per photon: flux.xyz,pos.x | pos.yz,phi,theta
        uint32_t n_octants = 0;
        unsigned long long n_photons = 0;
    };

    struct PhotonMaps
    {
        bool valid = false;
        DevicePhotonMap map[2]; // 0 caustic, 1 global
        uint32_t k_nearest = 0;
        uint32_t direct_visualization = 0;
        std::vector<void*> allocs;
        // k-NN query queue (allocated per precision on demand)
        std::vector<void*> queue_allocs;
        void* queue = nullptr;
        uint32_t queue_capacity = 0;
        int queue_precision = -1;
    };

    template <class R> struct PhotonLaunchArgs
    {
        DevicePhotonMap map[2];
        uint32_t k_nearest, direct_visualization;
        void* queue;
    };

    void photonFree(PhotonMaps& pm);
    int photonUpload(PhotonMaps& pm, const mcrt_photon_map_desc& caustic, const mcrt_photon_map_desc& global,
                     uint32_t k_nearest, uint32_t direct_visualization, cudaStream_t s, uint64_t& bytes, std::string& err);
    template <class R> int photonEnsureQueue(PhotonMaps& pm, uint32_t capacity, std::string& err);
    template <class R> PhotonLaunchArgs<R> photonLaunchArgs(const PhotonMaps& pm);
    template <class R> void photonShade(const WaveParams<R>& p, const PhotonLaunchArgs<R>& a, int cur, int grid, cudaStream_t s);
    int photonKnnUser(PhotonMaps& pm, int which, const double* points, size_t n, uint32_t* out_index, double* out_dist2,
                      uint32_t* out_count, int sm_count, cudaStream_t s, std::string& err);
}
