// sorobn_b200 -- templates that instantiate and launch the step kernels; included ONLY by the
// sbn_tiled_*.cu translation units (each instantiates its share of the combinations).
#pragma once
#include "sbn_kernels.cuh"
#include "sbn_launch.h"

namespace {

constexpr int kV = kRowsPerThread;  // evidence rows per thread of the tiled kernel

template <int N_IN>
cudaError_t launch_batched_n(const SbnStep &q, int64_t grid, cudaStream_t stream) {
    const size_t smem = static_cast<size_t>(q.smem_floats) * 4;
    const dim3 g(static_cast<unsigned>(grid)), b(SBN_THREADS);
#define SBN_CASE(CXV)                                                         \
    case CXV:                                                                 \
        sbn_launch(sbn_step_batched<N_IN, CXV>, g, b, smem, stream, q);               \
        break;
    if constexpr (N_IN <= 4) {
        switch (q.cx) {
            SBN_CASE(1)
            SBN_CASE(2)
            SBN_CASE(3)
            SBN_CASE(4)
            SBN_CASE(5)
            SBN_CASE(6)
            SBN_CASE(8)
            default:
                sbn_launch(sbn_step_batched<N_IN, 0>, g, b, smem, stream, q);
        }
    } else {
        sbn_launch(sbn_step_batched<N_IN, 0>, g, b, smem, stream, q);
    }
#undef SBN_CASE
    return cudaGetLastError();
}

template <int N_IN, int CX>
cudaError_t set_smem_attr() {
    return cudaFuncSetAttribute(sbn_step_batched<N_IN, CX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                SBN_SMEM_BUDGET);
}
template <int N_IN>
cudaError_t set_smem_attr_n() {
    cudaError_t e = set_smem_attr<N_IN, 0>();
    if constexpr (N_IN <= 4) {
        if (e == cudaSuccess) e = set_smem_attr<N_IN, 1>();
        if (e == cudaSuccess) e = set_smem_attr<N_IN, 2>();
        if (e == cudaSuccess) e = set_smem_attr<N_IN, 3>();
        if (e == cudaSuccess) e = set_smem_attr<N_IN, 4>();
        if (e == cudaSuccess) e = set_smem_attr<N_IN, 5>();
        if (e == cudaSuccess) e = set_smem_attr<N_IN, 6>();
        if (e == cudaSuccess) e = set_smem_attr<N_IN, 8>();
    }
    return e;
}

// (NU, NA, NB, NC) combinations instantiated.  Without a C-side input: NU <= 2, 1 <= NA <= 2,
// NB <= 2, at most 4 inputs.  With one: NU, NA, NB <= 1.
#define SBN_TILED_COMBOS(X)                                                                          \
    X(0, 1, 0, 0) X(0, 1, 1, 0) X(0, 1, 2, 0) X(0, 2, 0, 0) X(0, 2, 1, 0) X(0, 2, 2, 0) X(1, 1, 0, 0)   \
    X(1, 1, 1, 0) X(1, 1, 2, 0) X(1, 2, 0, 0) X(1, 2, 1, 0) X(2, 1, 0, 0) X(2, 1, 1, 0) X(2, 2, 0, 0)
#define SBN_TILED_COMBOS_C(X)                                                                        \
    X(0, 0, 0, 1) X(0, 0, 1, 1) X(0, 1, 0, 1) X(0, 1, 1, 1) X(1, 0, 0, 1) X(1, 0, 1, 1) X(1, 1, 0, 1)   \
    X(1, 1, 1, 1)

// Preload variants (CX > 0) exist where the tile edge equals the eliminated cardinality
// (networks with one cardinality throughout: 2, 3, 4, 5 states) and for 8 states (T = 4);
// never with a C-side input or several eliminated variables.
template <int NU, int NA, int NB, int NC>
cudaError_t launch_tiled_c(const SbnStep &q, int tile, bool preload, int64_t grid, cudaStream_t stream) {
    const size_t smem = static_cast<size_t>(q.smem_floats) * 4;
    const dim3 g(static_cast<unsigned>(grid)), b(tiled_threads());
    if constexpr (NC > 0) {
        switch (tile) {
            case 2: sbn_launch(sbn_step_tiled<NU, NA, NB, NC, 2, kV, 0>, g, b, smem, stream, q); break;
            case 3: sbn_launch(sbn_step_tiled<NU, NA, NB, NC, 3, kV, 0>, g, b, smem, stream, q); break;
            case 4: sbn_launch(sbn_step_tiled<NU, NA, NB, NC, 4, kV, 0>, g, b, smem, stream, q); break;
            case 5: sbn_launch(sbn_step_tiled<NU, NA, NB, NC, 5, kV, 0>, g, b, smem, stream, q); break;
            default: return cudaErrorInvalidValue;
        }
    } else {
#define SBN_T(TV)                                                                            \
    case TV:                                                                                 \
        if (preload && q.cx == TV && !q.zoff) sbn_launch(sbn_step_tiled<NU, NA, NB, 0, TV, kV, TV>, g, b, smem, stream, q); \
        else if (preload && q.cx_inner == TV && q.zoff != nullptr)                                                  \
            sbn_launch(sbn_step_tiled<NU, NA, NB, 0, TV, kV, TV, false, true>, g, b, smem, stream, q);                  \
        else sbn_launch(sbn_step_tiled<NU, NA, NB, 0, TV, kV, 0>, g, b, smem, stream, q);            \
        break;
        switch (tile) {
            SBN_T(2)
            SBN_T(3)
            SBN_T(5)
            case 4:
                if (preload && q.cx == 4 && !q.zoff) sbn_launch(sbn_step_tiled<NU, NA, NB, 0, 4, kV, 4>, g, b, smem, stream, q);
                else if (preload && q.cx == 8 && !q.zoff) sbn_launch(sbn_step_tiled<NU, NA, NB, 0, 4, kV, 8>, g, b, smem, stream, q);
                else if (preload && q.cx_inner == 4 && q.zoff != nullptr)
                    sbn_launch(sbn_step_tiled<NU, NA, NB, 0, 4, kV, 4, false, true>, g, b, smem, stream, q);
                else if (preload && q.cx_inner == 8 && q.zoff != nullptr)
                    sbn_launch(sbn_step_tiled<NU, NA, NB, 0, 4, kV, 8, false, true>, g, b, smem, stream, q);
                else sbn_launch(sbn_step_tiled<NU, NA, NB, 0, 4, kV, 0>, g, b, smem, stream, q);
                break;
            default: return cudaErrorInvalidValue;
        }
#undef SBN_T
    }
    return cudaGetLastError();
}

template <int NU>
cudaError_t launch_slab(const SbnStep &q, int tile, int64_t grid, cudaStream_t stream) {
    const size_t smem = (static_cast<size_t>(q.slab_smem_off) + static_cast<size_t>(q.n_slab) * kSlabThreads * kV) * 4;
    const dim3 g(static_cast<unsigned>(grid)), b(kSlabThreads);
    switch (tile) {
        case 2: sbn_launch(sbn_step_tiled<NU, 1, 1, 0, 2, kV, 2, true>, g, b, smem, stream, q); break;
        case 3: sbn_launch(sbn_step_tiled<NU, 1, 1, 0, 3, kV, 3, true>, g, b, smem, stream, q); break;
        case 5: sbn_launch(sbn_step_tiled<NU, 1, 1, 0, 5, kV, 5, true>, g, b, smem, stream, q); break;
        case 4:
            if (q.cx == 8) sbn_launch(sbn_step_tiled<NU, 1, 1, 0, 4, kV, 8, true>, g, b, smem, stream, q);
            else sbn_launch(sbn_step_tiled<NU, 1, 1, 0, 4, kV, 4, true>, g, b, smem, stream, q);
            break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <int NU>
cudaError_t set_slab_attr() {
    const int bytes = static_cast<int>(kSlabSmemMax) + SBN_SMEM_BUDGET;
    cudaError_t e = cudaFuncSetAttribute(sbn_step_tiled<NU, 1, 1, 0, 2, kV, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sbn_step_tiled<NU, 1, 1, 0, 3, kV, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sbn_step_tiled<NU, 1, 1, 0, 4, kV, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sbn_step_tiled<NU, 1, 1, 0, 4, kV, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sbn_step_tiled<NU, 1, 1, 0, 5, kV, 5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e;
}

template <int NU, int NA, int NB, int NC>
cudaError_t set_tiled_attr_c() {
    cudaError_t e = cudaSuccess;
#define SBN_A(TV, CXV) \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sbn_step_tiled<NU, NA, NB, NC, TV, kV, CXV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SBN_SMEM_BIG);
    SBN_A(2, 0) SBN_A(3, 0) SBN_A(4, 0) SBN_A(5, 0)
    if constexpr (NC == 0) {
        SBN_A(2, 2) SBN_A(3, 3) SBN_A(4, 4) SBN_A(4, 8) SBN_A(5, 5)
#define SBN_AM(TV) \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sbn_step_tiled<NU, NA, NB, NC, TV, kV, TV, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SBN_SMEM_BIG);
        SBN_AM(2) SBN_AM(3) SBN_AM(4) SBN_AM(5)
#undef SBN_AM
        if (e == cudaSuccess) e = cudaFuncSetAttribute(sbn_step_tiled<NU, NA, NB, NC, 4, kV, 8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SBN_SMEM_BIG);
    }
#undef SBN_A
    return e;
}
}  // namespace
