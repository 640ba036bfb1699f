// BVH construction on the device (widened scope). See bvh_build.cu.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

namespace mcrt
{
    struct BvhBuildResult
    {
        // BVH::linear_tree in depth-first order (bvh.hpp:68-82) + BVH::ordered_surfaces as indices
        // into the caller's primitive array
        std::vector<double> node_bounds;
        std::vector<uint32_t> node_first_prim, node_prim_count, node_next_sibling, prim_order;
        double gpu_ms = 0.0;          // first kernel to last kernel, copies of the result excluded
        uint32_t iterations = 0;      // breadth-first build rounds
        uint32_t kernel_launches = 0;
    };

    // type: MCRT_BVH_*; returns an mcrt_status code, message in `error`
    int buildBvhOnDevice(const double* prim_bounds_host, uint32_t n_prims, const double scene_bounds[6], int type,
                         int bins_per_axis, int sm_count, cudaStream_t stream, BvhBuildResult& out, std::string& error);
}
