// Shared body of kernels_f64.cu / kernels_f32.cu: defines Launch<MCRT_REAL>.
#include "launch.h"

namespace mcrt
{
    template <> void Launch<MCRT_REAL>::generate(const WaveParams<MCRT_REAL>& p, int next, int grid, cudaStream_t s)
    {
        if (p.filmp.is_default_box) k_generate<MCRT_REAL, false><<<grid, 256, 0, s>>>(p, next);
        else k_generate<MCRT_REAL, true><<<grid, 256, 0, s>>>(p, next);
    }
    template <> void Launch<MCRT_REAL>::extend(const WaveParams<MCRT_REAL>& p, int cur, int grid, cudaStream_t s)
    {
        k_extend<MCRT_REAL><<<grid, 256, 0, s>>>(p, cur);
    }
    template <> void Launch<MCRT_REAL>::shade(const WaveParams<MCRT_REAL>& p, int cur, int grid, cudaStream_t s)
    {
        if (p.filmp.is_default_box) k_shade<MCRT_REAL, 0, false><<<grid * 2, 128, 0, s>>>(p, cur);
        else k_shade<MCRT_REAL, 0, true><<<grid * 2, 128, 0, s>>>(p, cur);
    }
    template <> void Launch<MCRT_REAL>::shadePhoton(const WaveParams<MCRT_REAL>& p, int cur, int grid, cudaStream_t s)
    {
        if (p.filmp.is_default_box) k_shade<MCRT_REAL, 1, false><<<grid * 2, 128, 0, s>>>(p, cur);
        else k_shade<MCRT_REAL, 1, true><<<grid * 2, 128, 0, s>>>(p, cur);
    }
    template <> void Launch<MCRT_REAL>::knn(const WaveParams<MCRT_REAL>& p, int grid, cudaStream_t s)
    {
        const dim3 g(grid * 2), b(32 * KNN_WARPS_PER_BLOCK);
        const size_t smem = knnSharedBytes(p.pm.k_nearest);
        // the photon maps hold at least k photons in every render that matters; if a map is smaller the
        // search clamps k itself and the register slots are simply not all used
        if (!p.filmp.is_default_box)
        {
            // filtered film: the rare configuration, one generic instantiation
            k_knn<MCRT_REAL, 0, true><<<g, b, smem, s>>>(p);
            return;
        }
        switch (knnSlotsFor(p.pm.k_nearest))
        {
            case 1: k_knn<MCRT_REAL, 1, false><<<g, b, smem, s>>>(p); break;
            case 2: k_knn<MCRT_REAL, 2, false><<<g, b, smem, s>>>(p); break;
            case 4: k_knn<MCRT_REAL, 4, false><<<g, b, smem, s>>>(p); break;
            case 8: k_knn<MCRT_REAL, 8, false><<<g, b, smem, s>>>(p); break;
            default: k_knn<MCRT_REAL, 0, false><<<g, b, smem, s>>>(p); break;
        }
    }
    template <> void Launch<MCRT_REAL>::shadow(const WaveParams<MCRT_REAL>& p, int grid, cudaStream_t s)
    {
        if (p.filmp.is_default_box) k_shadow<MCRT_REAL, false><<<grid, 256, 0, s>>>(p);
        else k_shadow<MCRT_REAL, true><<<grid, 256, 0, s>>>(p);
    }
    template <> void Launch<MCRT_REAL>::emitGenerate(const WaveParams<MCRT_REAL>& p, int next, int grid, cudaStream_t s)
    {
        k_emit_generate<MCRT_REAL><<<grid, 256, 0, s>>>(p, next);
    }
    template <> void Launch<MCRT_REAL>::emitShade(const WaveParams<MCRT_REAL>& p, int cur, int grid, cudaStream_t s)
    {
        k_emit_shade<MCRT_REAL><<<grid * 2, 128, 0, s>>>(p, cur);
    }
    template <> void Launch<MCRT_REAL>::traceUser(const DeviceScene<MCRT_REAL>& sc, const double* rays6, size_t n,
                                                  double* out_tuv, uint32_t* out_prim, Counters* c, int grid, cudaStream_t s)
    {
        k_trace_user<MCRT_REAL><<<grid, 256, 0, s>>>(sc, rays6, n, out_tuv, out_prim, c);
    }
}
