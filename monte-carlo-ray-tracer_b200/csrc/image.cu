// Image pipeline on the device (widened scope, SURVEY.md §8f-4, second half): Image::save
// (source/camera/image.cpp:37-51) = auto exposure (Image::getExposure, image.cpp:63-73) ->
// tone-mapping operator (source/camera/pixel-operators.cpp:7-51) -> auto gain (Image::getGain,
// image.cpp:78-88) -> sRGB gamma (source/color/srgb.hpp:54-62) -> 8-bit truncation in B,G,R order.
// Both histogram passes are the reference's Histogram (source/common/histogram.cpp): 65536 bins of
// width max/65536, level(p) = upper edge of the first bin at which the running count reaches
// floor(N*p). Counts are integers and the bin index is one IEEE division, so the device histogram
// equals the CPU one; the float expressions keep the reference's operation order (this TU is
// compiled with --fmad=false). The output is the byte payload of the reference's TGA file.
#include "image.h"

#include <cmath>
#include <cstring>

#include "../../include/mcrt_abi.h"

namespace mcrt
{
namespace
{
    constexpr uint32_t HIST_BINS = 65536;   // image.cpp:70,85

    struct ImageState
    {
        unsigned long long max_key;   // brightness maximum as a sortable key (values are >= 0 when used)
        uint32_t negative;            // Histogram::Histogram returns early on a negative value
        uint32_t level_bin;           // first bin reaching the requested count (HIST_BINS: none)
    };

    __device__ inline void hable(const double in[3], double out[3])
    {
        // pixel-operators.cpp:9-20
        constexpr double A = 0.15, B = 0.50, C = 0.10, D = 0.20, E = 0.02, F = 0.30, W = 11.2;
        auto f = [&](double x) { return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F; };
        const double fw = f(W);
        for (int c = 0; c < 3; c++) out[c] = f(in[c]) / fw;
    }

    __device__ inline void aces(const double in[3], double out[3])
    {
        // pixel-operators.cpp:22-42 (column-major glm::dmat3 times vector)
        const double v[3] = {0.59719 * in[0] + 0.35458 * in[1] + 0.04823 * in[2],
                             0.07600 * in[0] + 0.90834 * in[1] + 0.01566 * in[2],
                             0.02840 * in[0] + 0.13383 * in[1] + 0.83777 * in[2]};
        double r[3];
        for (int c = 0; c < 3; c++)
        {
            const double a = v[c] * (v[c] + 0.0245786) - 0.000090537;
            const double b = v[c] * (0.983729 * v[c] + 0.4329510) + 0.238081;
            r[c] = a / b;
        }
        const double o[3] = {1.60475 * r[0] + -0.53108 * r[1] + -0.07367 * r[2],
                             -0.10208 * r[0] + 1.10813 * r[1] + -0.00605 * r[2],
                             -0.00327 * r[0] + -0.07276 * r[1] + 1.07602 * r[2]};
        for (int c = 0; c < 3; c++) out[c] = fmin(fmax(o[c], 0.0), 1.0);   // glm::clamp = min(max(x, lo), hi)
    }

    __device__ inline void tonemap(uint32_t op, const double in[3], double out[3])
    {
        if (op == MCRT_TONEMAP_LINEAR) { out[0] = in[0]; out[1] = in[1]; out[2] = in[2]; }
        else if (op == MCRT_TONEMAP_ACES) aces(in, out);
        else hable(in, out);
    }

    // brightness of pass 0 (getExposure: the pixel) or pass 1 (getGain: tonemap(pixel * exposure))
    __device__ inline double brightness(const double* rgb, size_t i, int pass, uint32_t op, double exposure)
    {
        double p[3] = {rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]};
        if (pass == 1)
        {
            const double q[3] = {p[0] * exposure, p[1] * exposure, p[2] * exposure};
            tonemap(op, q, p);
        }
        return ((0.0 + p[0]) + p[1] + p[2]) / 3.0;   // glm::compAdd / 3.0
    }

    __global__ void k_image_max(const double* rgb, size_t n, int pass, uint32_t op, double exposure, ImageState* st)
    {
        unsigned long long local = 0;
        bool neg = false;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        {
            const double v = brightness(rgb, i, pass, op, exposure);
            if (v < 0.0) neg = true;
            else if (v > 0.0) { const unsigned long long k = (unsigned long long)__double_as_longlong(v); local = k > local ? k : local; }
        }
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long w = __shfl_xor_sync(0xFFFFFFFFu, local, o); local = w > local ? w : local; }
        if ((threadIdx.x & 31) == 0 && local) atomicMax(&st->max_key, local);
        if (neg) st->negative = 1u;
    }

    __global__ void k_image_hist(const double* rgb, size_t n, int pass, uint32_t op, double exposure, const ImageState* st, uint32_t* counts)
    {
        if (st->negative || st->max_key == 0) return;
        const double bin_size = __longlong_as_double((long long)st->max_key) / (double)HIST_BINS;   // histogram.cpp:17
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        {
            const double v = brightness(rgb, i, pass, op, exposure);
            unsigned long long b = (unsigned long long)(v / bin_size);
            if (b > HIST_BINS - 1) b = HIST_BINS - 1;
            atomicAdd(&counts[b], 1u);
        }
    }

    // Histogram::level (histogram.cpp:25-40): first bin where the running count reaches `num`
    __global__ void __launch_bounds__(1024) k_image_level(const uint32_t* counts, unsigned long long num, ImageState* st)
    {
        __shared__ unsigned long long partial[1024];
        constexpr uint32_t PER = HIST_BINS / 1024;
        const uint32_t t = threadIdx.x;
        unsigned long long sum = 0;
        for (uint32_t k = 0; k < PER; k++) sum += counts[t * PER + k];
        partial[t] = sum;
        __syncthreads();
        if (t == 0)
        {
            unsigned long long run = 0;
            for (uint32_t k = 0; k < 1024; k++) { const unsigned long long v = partial[k]; partial[k] = run; run += v; }
        }
        __syncthreads();
        unsigned long long run = partial[t];
        uint32_t found = HIST_BINS;
        for (uint32_t k = 0; k < PER; k++)
        {
            run += counts[t * PER + k];
            if (run >= num) { found = t * PER + k; break; }
        }
        if (found != HIST_BINS) atomicMin(&st->level_bin, found);
    }

    __global__ void k_image_map(const double* rgb, size_t n, uint32_t op, double exposure, double gain, uint8_t* out_bgr)
    {
        const double top = 255.99999999999997;   // std::nextafter(256.0, 0.0)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        {
            const double q[3] = {rgb[3 * i] * exposure, rgb[3 * i + 1] * exposure, rgb[3 * i + 2] * exposure};
            double t[3];
            tonemap(op, q, t);
            uint8_t byte[3];
            for (int c = 0; c < 3; c++)
            {
                const double x = t[c] * gain;
                const double g = x <= 0.0031308 ? 12.92 * x : 1.055 * pow(x, 1.0 / 2.4) - 0.055;   // sRGB::gammaCompress
                const double v = fmin(fmax(g, 0.0), 1.0) * top;                                      // truncate, pixel-operators.cpp:47-51
                byte[c] = (uint8_t)v;
            }
            out_bgr[3 * i] = byte[2]; out_bgr[3 * i + 1] = byte[1]; out_bgr[3 * i + 2] = byte[0];
        }
    }
}

#define IK(call)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) { error = std::string(#call) + ": " + cudaGetErrorString(e_); return MCRT_ERR_CUDA; } \
    } while (0)

// One histogram pass: returns Histogram(...).level(pct) for the brightness of `pass`.
static int histogramLevel(const double* d_rgb, size_t n, int pass, uint32_t op, double exposure, double pct, ImageState* d_state,
                          uint32_t* d_counts, int grid, cudaStream_t s, double& level, std::string& error)
{
    ImageState init; init.max_key = 0; init.negative = 0; init.level_bin = HIST_BINS;
    IK(cudaMemcpyAsync(d_state, &init, sizeof(init), cudaMemcpyHostToDevice, s));
    IK(cudaMemsetAsync(d_counts, 0, HIST_BINS * sizeof(uint32_t), s));
    k_image_max<<<grid, 256, 0, s>>>(d_rgb, n, pass, op, exposure, d_state);
    k_image_hist<<<grid, 256, 0, s>>>(d_rgb, n, pass, op, exposure, d_state, d_counts);
    const unsigned long long num = (unsigned long long)((double)n * pct);   // static_cast<size_t>(data_size * count_percentage)
    k_image_level<<<1, 1024, 0, s>>>(d_counts, num, d_state);
    ImageState st;
    IK(cudaMemcpyAsync(&st, d_state, sizeof(st), cudaMemcpyDeviceToHost, s));
    IK(cudaStreamSynchronize(s));
    IK(cudaGetLastError());
    level = 0.0;
    if (st.negative || st.max_key == 0 || st.level_bin == HIST_BINS) return MCRT_OK;   // empty histogram / all black: level 0
    double mx; std::memcpy(&mx, &st.max_key, 8);
    const double bin_size = mx / (double)HIST_BINS;
    level = (double)(st.level_bin + 1) * bin_size;   // (i + 1) * bin_size
    return MCRT_OK;
}

int imageTonemapOnDevice(const double* d_rgb, uint32_t width, uint32_t height, const mcrt_image_params& prm, uint8_t* d_out_bgr,
                         int sm_count, cudaStream_t s, double* exposure_out, double* gain_out, std::string& error)
{
    const size_t n = (size_t)width * height;
    const int grid = sm_count * 8;
    const uint32_t op = prm.plain ? (uint32_t)MCRT_TONEMAP_LINEAR : (prm.tonemapper == MCRT_TONEMAP_ACES ? (uint32_t)MCRT_TONEMAP_ACES : (uint32_t)MCRT_TONEMAP_HABLE);
    double exposure = 1.0, gain = 1.0;
    ImageState* d_state = nullptr; uint32_t* d_counts = nullptr;
    if (!prm.plain)
    {
        IK(cudaMalloc((void**)&d_state, sizeof(ImageState)));
        if (cudaMalloc((void**)&d_counts, HIST_BINS * sizeof(uint32_t)) != cudaSuccess) { cudaFree(d_state); error = "cudaMalloc failed"; return MCRT_ERR_CUDA; }
        const double exposure_scale = prm.exposure_scale, gain_scale = prm.gain_scale;   // image.cpp:21-22, 39-40
        double L = 0.0;
        int rc = histogramLevel(d_rgb, n, 0, op, 1.0, 0.5, d_state, d_counts, grid, s, L, error);
        if (rc == MCRT_OK)
        {
            exposure = (L > 0.0 ? 0.5 / L : 1.0) * exposure_scale;
            rc = histogramLevel(d_rgb, n, 1, op, exposure, 0.99, d_state, d_counts, grid, s, L, error);
            if (rc == MCRT_OK) gain = (L > 0.0 ? 0.99 / L : 1.0) * gain_scale;
        }
        cudaFree(d_state); cudaFree(d_counts);
        if (rc != MCRT_OK) return rc;
    }
    k_image_map<<<grid, 256, 0, s>>>(d_rgb, n, op, exposure, gain, d_out_bgr);
    IK(cudaStreamSynchronize(s));
    IK(cudaGetLastError());
    if (exposure_out) *exposure_out = exposure;
    if (gain_out) *gain_out = gain;
    return MCRT_OK;
}
}
