// sorobn_b200 -- step-kernel instantiations: two inputs without a tile axis (NU = 2), and the slab variants
// (one of four translation units that share the ~290 instantiations of sbn_step_tiled; see sbn_launch.h)
#include "sbn_launch_impl.cuh"

cudaError_t sbn_tiled_u2_launch(int key, const SbnStep &q, int tile, bool preload, int64_t grid, cudaStream_t stream) {
    switch (key) {
        case 2100: return launch_tiled_c<2, 1, 0, 0>(q, tile, preload, grid, stream);
        case 2110: return launch_tiled_c<2, 1, 1, 0>(q, tile, preload, grid, stream);
        case 2200: return launch_tiled_c<2, 2, 0, 0>(q, tile, preload, grid, stream);
    }
    return cudaErrorInvalidValue;
}

cudaError_t sbn_tiled_u2_set_attrs() {
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = set_tiled_attr_c<2, 1, 0, 0>();
    if (e == cudaSuccess) e = set_tiled_attr_c<2, 1, 1, 0>();
    if (e == cudaSuccess) e = set_tiled_attr_c<2, 2, 0, 0>();
    if (e == cudaSuccess) e = set_slab_attr<0>();
    if (e == cudaSuccess) e = set_slab_attr<1>();
    return e;
}

cudaError_t sbn_slab_launch(int nu, const SbnStep &q, int tile, int64_t grid, cudaStream_t stream) {
    return nu == 0 ? launch_slab<0>(q, tile, grid, stream) : launch_slab<1>(q, tile, grid, stream);
}
