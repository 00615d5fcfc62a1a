// sorobn_b200 -- engine internals shared by the translation units of libsorobn_b200.so
// (sbn_api.cu: parsing, classic step launches, C ABI; sbn_chain.cu: the on-chip segment kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/sorobn_b200.h"

struct EvAxis {
    int col, stride, card;
};
struct InDesc {
    bool is_slot;
    int id;
    bool batched;
    int sx;                      // stride of the eliminated axis when there is exactly one
    std::vector<EvAxis> ev;
    std::vector<int> estrides;   // stride per eliminated axis
    std::vector<int> strides;
};
struct StepDesc {
    int kind;
    int out_slot;
    int cx;                      // joint states of the eliminated variables (1 = product only)
    int64_t n_out;
    std::vector<int> cards;
    std::vector<int> ecards;     // cardinality per eliminated variable
    int64_t zoff_pos = -1;       // >= 0: int32 offset of the [n_in][cx] joint-state offset table
    int64_t zoff_tiled_pos = -1; // the same table with rows in the tiled kernel's input order
    std::vector<InDesc> in;
    // tiled fast path (sbn_step_tiled): tile edge, tile count, offset-table position
    int tile = 0;            // 0 = not eligible, use sbn_step_batched
    int nu = 0, na = 0, nb = 0, nc = 0;  // inputs without a tile axis / with axis 0 / axis 1 / both
    std::vector<int> order;      // kernel input slot -> index into `in` (U, then A, then B)
    int64_t n_tiles = 0;
    int64_t tile_off_pos = 0;  // int32 offset into sbn_program::d_tile_off
    // slab variant (expanding products): tiles grouped by the digits A and B share
    bool slab = false;
    bool big_tables = false;     // staged tables exceed SBN_SMEM_BUDGET: one CTA per SM (see smem_big)
    int slab_ma = 0, n_slab = 0;
    int64_t slab_off_pos = 0;
    int64_t slab_tile_off_pos = 0;  // tile table in slab order (rows of n_in + 5 words)
    int64_t tiles_per_super = 0, n_super = 0;
    // sliced staging (tables beyond SBN_SMEM_BUDGET): per chunk of `slice_tpc` tiles, per input,
    // (first float, floats, shared-memory offset) of the part of the table those tiles touch
    int64_t slice_pos = -1;
    int64_t slice_tpc = 0;
    int64_t slice_smem = 0;      // floats of shared memory of the largest chunk
};
struct Slot {
    bool batched;
    int64_t size;     // floats per row (batched) or in total
    int64_t padded;   // size rounded up to 4 floats (bulk-TMA granularity)
    float *ptr;
};

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

struct SbnSegment;  // sbn_chain.h
struct SbnPair;     // sbn_pair.h

struct sbn_program {
    int device = 0;
    bool f64 = false;  // single-event programs computed and returned in double
    int mode = 0, n_ev = 0, Q = 0, post_slot = 0, post_batched = 0;
    std::vector<std::pair<int64_t, int64_t>> tables;  // (offset, size) in floats
    std::vector<int64_t> table_padded;
    float *d_tables = nullptr;
    std::vector<Slot> slots;
    std::vector<StepDesc> steps;

    int64_t reserved_rows = 0;  // chunk capacity
    int64_t ld = 0;             // row pitch of batched scratch (floats)
    float *d_arena = nullptr;   // batched scratch
    float *d_shared = nullptr;  // unbatched scratch
    int32_t *d_tile_off = nullptr;  // per-step tile offset tables of the tiled kernel
    float *d_total = nullptr;   // per-row normaliser = P(event) of the last run [ld] (double when f64)
    uint8_t *d_ev = nullptr;    // staging for run_host  [n_ev][ld]
    float *d_out = nullptr;     //                         [Q][ld]   (double when f64)
    cudaStream_t stream = nullptr;
    // Branch streams for graph capture: the steps form a tree (every intermediate is consumed
    // once), so independent sub-trees are captured on different streams and become parallel
    // branches of the CUDA graph.
    static constexpr int kBranches = 4;
    cudaStream_t branch[kBranches] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> step_done;  // one event per step (+ normalise), capture-only
    std::vector<cudaEvent_t> pipe_events;  // run_host pipelining: fork, (upload done, kernels done) per column range, joins
    cudaGraphExec_t pipe_exec = nullptr;   // the pipelined host run as one graph (pinned host buffers)
    struct {
        const void *ev;
        int64_t ld_ev, n_rows;
        const void *out;
        int64_t ld_out;
    } pipe_key = {nullptr, 0, 0, nullptr, 0};
    int64_t pipe_launches = 0;
    bool use_branches = false;  // measured: no gain on the grid plan (one long chain); opt-in

    bool use_graph = true;
    bool use_tiled = true;
    bool use_slab = true;
    bool use_preload = true;  // tiled kernel: operand preload schedule where instantiated (else the x-loop)
    cudaGraphExec_t exec = nullptr;
    struct {
        const uint8_t *ev;
        int64_t ld_ev, n_rows;
        float *out;
        int64_t ld_out;
    } graph_key = {nullptr, 0, 0, nullptr, 0};
    int64_t graph_launches = 0;

    int64_t launches = 0;
    int64_t setup_launches = 0;  // evidence-independent launches issued once at creation

    // on-chip segments (sbn_chain.h): runs of batched steps executed by one persistent kernel
    std::vector<SbnSegment *> segments;
    std::vector<int> seg_first;      // per step: >= 0 = head of that segment, -2 = inside one, -1 = classic launch
    float *d_chain_scratch = nullptr;
    bool use_chain = true;
    bool use_tma = true;             // tensor-map TMA pipeline kernel for the steps it covers (sbn_tma.h)
    bool chain_fits = true;          // false when a slot-arena operand of a segment needs > 32-bit byte offsets
    std::vector<int32_t> h_tile_words;  // host copy of d_tile_off

    // paired steps (sbn_pair.h): a step and its consumer as one launch, the intermediate in registers
    std::vector<SbnPair *> pairs;
    std::vector<int> pair_first;     // per step: >= 0 = first step of that pair, -2 = its second step, -1 = on its own
    float *d_pair_canon = nullptr;   // canonical coefficient arrays of all pairs
    int32_t *d_pair_tiles = nullptr; // their tile tables
    bool use_pair = true;
    bool pairs_avoid_segments = false;  // planned with the on-chip segments switched on: no pair touches a segment's step
};

