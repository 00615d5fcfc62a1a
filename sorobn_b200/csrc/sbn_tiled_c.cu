// sorobn_b200 -- step-kernel instantiations: one input spans both tile axes (NC = 1), and the plain batched kernel
// (one of four translation units that share the ~290 instantiations of sbn_step_tiled; see sbn_launch.h)
#include "sbn_launch_impl.cuh"

cudaError_t sbn_tiled_c_launch(int key, const SbnStep &q, int tile, bool preload, int64_t grid, cudaStream_t stream) {
    switch (key) {
        case 1: return launch_tiled_c<0, 0, 0, 1>(q, tile, preload, grid, stream);
        case 11: return launch_tiled_c<0, 0, 1, 1>(q, tile, preload, grid, stream);
        case 101: return launch_tiled_c<0, 1, 0, 1>(q, tile, preload, grid, stream);
        case 111: return launch_tiled_c<0, 1, 1, 1>(q, tile, preload, grid, stream);
        case 1001: return launch_tiled_c<1, 0, 0, 1>(q, tile, preload, grid, stream);
        case 1011: return launch_tiled_c<1, 0, 1, 1>(q, tile, preload, grid, stream);
        case 1101: return launch_tiled_c<1, 1, 0, 1>(q, tile, preload, grid, stream);
        case 1111: return launch_tiled_c<1, 1, 1, 1>(q, tile, preload, grid, stream);
    }
    return cudaErrorInvalidValue;
}

cudaError_t sbn_tiled_c_set_attrs() {
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = set_tiled_attr_c<0, 0, 0, 1>();
    if (e == cudaSuccess) e = set_tiled_attr_c<0, 0, 1, 1>();
    if (e == cudaSuccess) e = set_tiled_attr_c<0, 1, 0, 1>();
    if (e == cudaSuccess) e = set_tiled_attr_c<0, 1, 1, 1>();
    if (e == cudaSuccess) e = set_tiled_attr_c<1, 0, 0, 1>();
    if (e == cudaSuccess) e = set_tiled_attr_c<1, 0, 1, 1>();
    if (e == cudaSuccess) e = set_tiled_attr_c<1, 1, 0, 1>();
    if (e == cudaSuccess) e = set_tiled_attr_c<1, 1, 1, 1>();
    return e;
}

cudaError_t sbn_batched_launch(const SbnStep &q, int64_t grid, cudaStream_t stream) {
    switch (q.n_in) {
        case 1: return launch_batched_n<1>(q, grid, stream);
        case 2: return launch_batched_n<2>(q, grid, stream);
        case 3: return launch_batched_n<3>(q, grid, stream);
        case 4: return launch_batched_n<4>(q, grid, stream);
        case 5: return launch_batched_n<5>(q, grid, stream);
        case 6: return launch_batched_n<6>(q, grid, stream);
        case 7: return launch_batched_n<7>(q, grid, stream);
        case 8: return launch_batched_n<8>(q, grid, stream);
    }
    return cudaErrorInvalidValue;
}

cudaError_t sbn_batched_set_attrs() {
    cudaError_t e = set_smem_attr_n<1>();
    if (e == cudaSuccess) e = set_smem_attr_n<2>();
    if (e == cudaSuccess) e = set_smem_attr_n<3>();
    if (e == cudaSuccess) e = set_smem_attr_n<4>();
    if (e == cudaSuccess) e = set_smem_attr_n<5>();
    if (e == cudaSuccess) e = set_smem_attr_n<6>();
    if (e == cudaSuccess) e = set_smem_attr_n<7>();
    if (e == cudaSuccess) e = set_smem_attr_n<8>();
    return e;
}
