// Image pipeline on the device: see image.cu.
#pragma once

#include <cstdint>
#include <string>

#include <cuda_runtime.h>

struct mcrt_image_params;

namespace mcrt
{
    // d_rgb: [height][width][3] float64 linear radiance; d_out_bgr: [height][width][3] bytes (TGA payload)
    int imageTonemapOnDevice(const double* d_rgb, uint32_t width, uint32_t height, const mcrt_image_params& params,
                             uint8_t* d_out_bgr, int sm_count, cudaStream_t stream, double* exposure_factor,
                             double* gain_factor, std::string& error);
}
