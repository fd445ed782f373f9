// placeholder until the tcgen05 path lands (next commit)
#include "vq_common.cuh"
bool vq_scan_mma_supported(int64_t, int, int, int, int, int) { return false; }
size_t vq_scan_mma_workspace(int64_t, int, int, int, int) { return 0; }
int vq_scan_mma_run(const void*, int64_t, int, int, int, const float*, int, int, float*, int32_t*, void*, size_t,
                    cudaStream_t, int*) {
    vq_set_error("tcgen05 scan path not built");
    return VQ_EUNSUPPORTED;
}
