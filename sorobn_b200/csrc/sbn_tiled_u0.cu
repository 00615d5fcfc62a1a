// sorobn_b200 -- step-kernel instantiations: every input carries a tile axis (NU = 0)
// (one of four translation units that share the ~290 instantiations of sbn_step_tiled; see sbn_launch.h)
#include "sbn_launch_impl.cuh"

cudaError_t sbn_tiled_u0_launch(int key, const SbnStep &q, int tile, bool preload, int64_t grid, cudaStream_t stream) {
    switch (key) {
        case 100: return launch_tiled_c<0, 1, 0, 0>(q, tile, preload, grid, stream);
        case 110: return launch_tiled_c<0, 1, 1, 0>(q, tile, preload, grid, stream);
        case 120: return launch_tiled_c<0, 1, 2, 0>(q, tile, preload, grid, stream);
        case 200: return launch_tiled_c<0, 2, 0, 0>(q, tile, preload, grid, stream);
        case 210: return launch_tiled_c<0, 2, 1, 0>(q, tile, preload, grid, stream);
        case 220: return launch_tiled_c<0, 2, 2, 0>(q, tile, preload, grid, stream);
    }
    return cudaErrorInvalidValue;
}

cudaError_t sbn_tiled_u0_set_attrs() {
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = set_tiled_attr_c<0, 1, 0, 0>();
    if (e == cudaSuccess) e = set_tiled_attr_c<0, 1, 1, 0>();
    if (e == cudaSuccess) e = set_tiled_attr_c<0, 1, 2, 0>();
    if (e == cudaSuccess) e = set_tiled_attr_c<0, 2, 0, 0>();
    if (e == cudaSuccess) e = set_tiled_attr_c<0, 2, 1, 0>();
    if (e == cudaSuccess) e = set_tiled_attr_c<0, 2, 2, 0>();
    return e;
}
