// placeholder: HNSW kernels land in the next commit
#include "vq_common.cuh"
extern "C" {
size_t vq_hnsw_workspace_bytes(int, int, int) { return 256; }
int vq_hnsw_search(const void*, int64_t, int, int, int, const int32_t*, const int32_t*, int, const int32_t*,
                   const int32_t*, int, int32_t, int, int, const float*, int, int, int, float*, int32_t*, uint32_t*,
                   void*, size_t, void*) {
    vq_set_error("hnsw_search not built yet");
    return VQ_EUNSUPPORTED;
}
size_t vq_hnsw_layer_workspace_bytes(int64_t, int, int, int, int, int) { return 256; }
int vq_hnsw_build_layer(const void*, int64_t, int, int, int, const int32_t*, int64_t, int, int, int, int32_t*, void*,
                        size_t, void*) {
    vq_set_error("hnsw_build_layer not built yet");
    return VQ_EUNSUPPORTED;
}
}
