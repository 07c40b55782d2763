// r1_wavefront.cuh -- wavefront variant (generate / intersect / shade kernels over compacted ray queues).
#pragma once
#include "r1_kernels.cuh"

namespace r1 {

struct WavefrontBuffers {
    void *pool = nullptr;
    size_t pool_bytes = 0;
};

inline void wavefront_free(WavefrontBuffers &b)
{
    if (b.pool) cudaFree(b.pool);
    b.pool = nullptr;
    b.pool_bytes = 0;
}

// returns a cudaError_t (0 = ok)
inline int wavefront_render(WavefrontBuffers &, const RenderArgs &, int, cudaStream_t, uint32_t *launches)
{
    *launches = 0;
    return (int)cudaErrorNotSupported;
}

}  // namespace r1
