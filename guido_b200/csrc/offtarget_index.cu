// offtarget_index.cu -- guide-side seed index off-target search.
#include "scan_common.cuh"

namespace guido {

int launch_ot_index(DeviceState *, const PlanesView &, const OtParams &, const ScanRange &, const uint2 *,
                    int64_t, guido_ot_hit *, uint64_t, unsigned long long *, unsigned long long *,
                    cudaStream_t, int *) {
    set_error("seed-index search is not built yet");
    return GUIDO_ERR_ARG;
}

}  // namespace guido
