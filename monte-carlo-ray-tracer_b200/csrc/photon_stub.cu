// Temporary: photon-mapped path not yet implemented.
#include "photon.h"

namespace mcrt
{
    void photonFree(PhotonMaps& pm)
    {
        for (void* p : pm.allocs) cudaFree(p);
        for (void* p : pm.queue_allocs) cudaFree(p);
        pm.allocs.clear(); pm.queue_allocs.clear(); pm.valid = false;
    }
    int photonUpload(PhotonMaps&, const mcrt_photon_map_desc&, const mcrt_photon_map_desc&, uint32_t, uint32_t, cudaStream_t,
                     uint64_t&, std::string& err) { err = "photon mapping not implemented yet"; return MCRT_ERR_UNSUPPORTED; }
    template <class R> int photonEnsureQueue(PhotonMaps&, uint32_t, std::string& err) { err = "photon mapping not implemented yet"; return MCRT_ERR_UNSUPPORTED; }
    template <class R> PhotonLaunchArgs<R> photonLaunchArgs(const PhotonMaps&) { return PhotonLaunchArgs<R>(); }
    template <class R> void photonShade(const WaveParams<R>&, const PhotonLaunchArgs<R>&, int, int, cudaStream_t) { }
    int photonKnnUser(PhotonMaps&, int, const double*, size_t, uint32_t*, double*, uint32_t*, int, cudaStream_t, std::string& err)
    { err = "photon mapping not implemented yet"; return MCRT_ERR_UNSUPPORTED; }

    template int photonEnsureQueue<double>(PhotonMaps&, uint32_t, std::string&);
    template int photonEnsureQueue<float>(PhotonMaps&, uint32_t, std::string&);
    template PhotonLaunchArgs<double> photonLaunchArgs<double>(const PhotonMaps&);
    template PhotonLaunchArgs<float> photonLaunchArgs<float>(const PhotonMaps&);
    template void photonShade<double>(const WaveParams<double>&, const PhotonLaunchArgs<double>&, int, int, cudaStream_t);
    template void photonShade<float>(const WaveParams<float>&, const PhotonLaunchArgs<float>&, int, int, cudaStream_t);
}
