// mmej.cu -- placeholder until the warp kernel lands (next commit)
#include "device.cuh"
using namespace guido;
extern "C" int guido_mmej_batch(const char *, const int64_t *, const int32_t *, int64_t, int32_t, uint32_t *,
                                int64_t, int64_t *, int32_t *, guido_stats *) {
    set_error("guido_mmej_batch is not built yet");
    return GUIDO_ERR_ARG;
}
