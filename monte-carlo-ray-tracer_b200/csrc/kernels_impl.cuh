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
        // triangle-only scenes (every OBJ scene) run the traversal without sphere / quadric code
        if constexpr (Mode<MCRT_REAL>::parity)
        {
            if (p.scene.bvh4 && p.scene.dynamic_fetch)
            {
                if (p.scene.prims_class == PRIMS_TRI) k_extend<MCRT_REAL, PRIMS_TRI, 2><<<grid, 256, fastStackSharedBytes(256), s>>>(p, cur);
                else if (p.scene.prims_class == PRIMS_TRI_SPHERE) k_extend<MCRT_REAL, PRIMS_TRI_SPHERE, 2><<<grid, 256, fastStackSharedBytes(256), s>>>(p, cur);
                else k_extend<MCRT_REAL, PRIMS_ALL, 2><<<grid, 256, fastStackSharedBytes(256), s>>>(p, cur);
                return;
            }
            if (p.scene.bvh4)
            {
                if (p.scene.prims_class == PRIMS_TRI) k_extend<MCRT_REAL, PRIMS_TRI, 1><<<grid, 256, fastStackSharedBytes(256), s>>>(p, cur);
                else if (p.scene.prims_class == PRIMS_TRI_SPHERE) k_extend<MCRT_REAL, PRIMS_TRI_SPHERE, 1><<<grid, 256, fastStackSharedBytes(256), s>>>(p, cur);
                else k_extend<MCRT_REAL, PRIMS_ALL, 1><<<grid, 256, fastStackSharedBytes(256), s>>>(p, cur);
                return;
            }
        }
        if (p.scene.prims_class == PRIMS_TRI) k_extend<MCRT_REAL, PRIMS_TRI, 0><<<grid, 256, 0, s>>>(p, cur);
        else if (p.scene.prims_class == PRIMS_TRI_SPHERE) k_extend<MCRT_REAL, PRIMS_TRI_SPHERE, 0><<<grid, 256, 0, s>>>(p, cur);
        else k_extend<MCRT_REAL, PRIMS_ALL, 0><<<grid, 256, 0, s>>>(p, cur);
    }
    template <> void Launch<MCRT_REAL>::shade(const WaveParams<MCRT_REAL>& p, int cur, int grid, cudaStream_t s)
    {
        // scenes whose materials use no Oren-Nayar / GGX / conductor Fresnel run the instantiation without that code
        const bool lite = (p.scene.material_flags_any & ~SHADE_FEATS_LITE) == 0;
        if (!p.filmp.is_default_box) k_shade<MCRT_REAL, 0, true, SHADE_FEATS_ALL><<<grid * 2, 128, 0, s>>>(p, cur);
        else if (lite) k_shade<MCRT_REAL, 0, false, SHADE_FEATS_LITE><<<grid * 2, 128, 0, s>>>(p, cur);
        else k_shade<MCRT_REAL, 0, false, SHADE_FEATS_ALL><<<grid * 2, 128, 0, s>>>(p, cur);
    }
    template <> void Launch<MCRT_REAL>::shadePhoton(const WaveParams<MCRT_REAL>& p, int cur, int grid, cudaStream_t s)
    {
        const bool lite = (p.scene.material_flags_any & ~SHADE_FEATS_LITE) == 0;
        if (!p.filmp.is_default_box) k_shade<MCRT_REAL, 1, true, SHADE_FEATS_ALL><<<grid * 2, 128, 0, s>>>(p, cur);
        else if (lite) k_shade<MCRT_REAL, 1, false, SHADE_FEATS_LITE><<<grid * 2, 128, 0, s>>>(p, cur);
        else k_shade<MCRT_REAL, 1, false, SHADE_FEATS_ALL><<<grid * 2, 128, 0, s>>>(p, cur);
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
            static bool attr_set = false;
            if (!attr_set) { cudaFuncSetAttribute(k_knn<MCRT_REAL, 0, true, SHADE_FEATS_ALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)knnSharedBytes(1024)); attr_set = true; }
            k_knn<MCRT_REAL, 0, true, SHADE_FEATS_ALL><<<g, b, smem, s>>>(p);
            return;
        }
        const bool lite = (p.scene.material_flags_any & ~SHADE_FEATS_LITE) == 0;
        const int slots = knnSlotsFor(p.pm.k_nearest);
        // k > 768 needs more than the default 48 KB of dynamic shared memory (knnSharedBytes)
        #define MCRT_KNN_LAUNCH1(SL, FE) \
            do { static bool attr_set = false; \
                 if (!attr_set) { cudaFuncSetAttribute(k_knn<MCRT_REAL, SL, false, FE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)knnSharedBytes(1024)); attr_set = true; } \
                 k_knn<MCRT_REAL, SL, false, FE><<<g, b, smem, s>>>(p); } while (0)
        #define MCRT_KNN_LAUNCH(SL) \
            do { if (lite) MCRT_KNN_LAUNCH1(SL, SHADE_FEATS_LITE); else MCRT_KNN_LAUNCH1(SL, SHADE_FEATS_ALL); } while (0)
        switch (slots)
        {
            case 1: MCRT_KNN_LAUNCH(1); break;
            case 2: MCRT_KNN_LAUNCH(2); break;
            case 4: MCRT_KNN_LAUNCH(4); break;
            case 8: MCRT_KNN_LAUNCH(8); break;
            default: MCRT_KNN_LAUNCH(0); break;
        }
        #undef MCRT_KNN_LAUNCH
        #undef MCRT_KNN_LAUNCH1
    }
    template <> void Launch<MCRT_REAL>::shadow(const WaveParams<MCRT_REAL>& p, int grid, cudaStream_t s)
    {
        if constexpr (Mode<MCRT_REAL>::parity)
        {
            if (p.scene.bvh4 && p.scene.dynamic_fetch && p.filmp.is_default_box)
            {
                if (p.scene.prims_class == PRIMS_TRI) k_shadow<MCRT_REAL, false, PRIMS_TRI, 2><<<grid, 256, fastStackSharedBytes(256), s>>>(p);
                else if (p.scene.prims_class == PRIMS_TRI_SPHERE) k_shadow<MCRT_REAL, false, PRIMS_TRI_SPHERE, 2><<<grid, 256, fastStackSharedBytes(256), s>>>(p);
                else k_shadow<MCRT_REAL, false, PRIMS_ALL, 2><<<grid, 256, fastStackSharedBytes(256), s>>>(p);
                return;
            }
            if (p.scene.bvh4)
            {
                if (!p.filmp.is_default_box) k_shadow<MCRT_REAL, true, PRIMS_ALL, 1><<<grid, 256, fastStackSharedBytes(256), s>>>(p);
                else if (p.scene.prims_class == PRIMS_TRI) k_shadow<MCRT_REAL, false, PRIMS_TRI, 1><<<grid, 256, fastStackSharedBytes(256), s>>>(p);
                else if (p.scene.prims_class == PRIMS_TRI_SPHERE) k_shadow<MCRT_REAL, false, PRIMS_TRI_SPHERE, 1><<<grid, 256, fastStackSharedBytes(256), s>>>(p);
                else k_shadow<MCRT_REAL, false, PRIMS_ALL, 1><<<grid, 256, fastStackSharedBytes(256), s>>>(p);
                return;
            }
        }
        if (!p.filmp.is_default_box) k_shadow<MCRT_REAL, true, PRIMS_ALL, 0><<<grid, 256, 0, s>>>(p);
        else if (p.scene.prims_class == PRIMS_TRI) k_shadow<MCRT_REAL, false, PRIMS_TRI, 0><<<grid, 256, 0, s>>>(p);
        else if (p.scene.prims_class == PRIMS_TRI_SPHERE) k_shadow<MCRT_REAL, false, PRIMS_TRI_SPHERE, 0><<<grid, 256, 0, s>>>(p);
        else k_shadow<MCRT_REAL, false, PRIMS_ALL, 0><<<grid, 256, 0, s>>>(p);
    }
    template <> void Launch<MCRT_REAL>::shadeKey(const WaveParams<MCRT_REAL>& p, int grid, cudaStream_t s)
    {
        k_shade_key<MCRT_REAL><<<grid, 256, 0, s>>>(p);
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
        if constexpr (Mode<MCRT_REAL>::parity)
        {
            if (sc.bvh4) { k_trace_user<MCRT_REAL, true><<<grid, 256, fastStackSharedBytes(256), s>>>(sc, rays6, n, out_tuv, out_prim, c); return; }
        }
        k_trace_user<MCRT_REAL, false><<<grid, 256, 0, s>>>(sc, rays6, n, out_tuv, out_prim, c);
    }
}
