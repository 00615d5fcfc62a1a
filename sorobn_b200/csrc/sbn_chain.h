// sorobn_b200 -- on-chip segments: runs of elimination steps executed by ONE persistent kernel.
//
// Every step of the elimination is row-local: evidence row b of an output depends on row b of
// the inputs only.  The classic path (sbn_kernels.cuh) still writes every intermediate factor
// to HBM and reads it back in the next launch -- 85 of the 158 KB per query on the benchmark
// grid are the 2.5 KB frontier of `625 <- sum_5 table x B625` going out and coming back.
// A *segment* is a contiguous run of batched steps that one CTA executes back to back for a
// block of 32 evidence rows: intermediates that are produced and consumed inside the segment
// live in shared memory (`[entry][32 rows]`, lanes = rows, conflict-free), or -- when shared
// memory is full -- in a per-CTA private scratch in global memory that stays L2-resident
// because the persistent CTA overwrites it row block after row block.  Only what crosses a
// segment boundary touches the slot arena in HBM.  When the segment ends in the posterior the
// normalisation (bayes_net.py:789-790) is fused as well.
//
// The reference's own loop for the same thing: bayes_net.py:778-786 (one pandas product +
// groupby-sum per eliminated variable).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/sorobn_b200.h"

#define SBN_CHAIN_ROWS 32        // evidence rows per CTA iteration: one per lane
#define SBN_CHAIN_SLOTS 7        // input slots of a step: U0 U1 | A0 A1 | B0 B1 | C0
#define SBN_CHAIN_SMEM (227 * 1024)

// where a factor lives
#define SBN_SP_TABLE 0    // CPT / evidence-independent table: staged in shared memory, gathered per row
#define SBN_SP_SMEM 1     // batched intermediate in the shared-memory arena   [entry][32]
#define SBN_SP_GLOBAL 2   // batched factor in the HBM slot arena              [entry][ld]
#define SBN_SP_SCRATCH 3  // batched intermediate in the CTA's private scratch [entry][32]

struct SbnChainIn {
    const float *ptr;      // TABLE: source of the bulk copy; GLOBAL: slot base (row 0)
    int32_t space;
    int32_t off;           // floats -- TABLE: inside the step's table buffer; SMEM: inside the arena; SCRATCH: inside the CTA's scratch
    int32_t sx, s0, s1;    // element strides: first eliminated variable, tile axis 0, tile axis 1
    int32_t zrow;          // row of this input in the step's joint-state offset table (zoff)
    int32_t stage_floats;  // TABLE: floats to stage (multiple of 4)
    int32_t n_ev;
    int32_t ev_col[SBN_MAX_EV];
    int32_t ev_stride[SBN_MAX_EV];
    int32_t ev_card[SBN_MAX_EV];
};

struct SbnChainStep {
    const int32_t *tile_off;   // [n_tiles][n_in + 2]: out entry, na | nb << 8, input element offsets (tiled-kernel order)
    const int32_t *zoff;       // [n_in][cx] joint-state element offsets, or nullptr (one eliminated variable: x * sx)
    float *out_ptr;            // GLOBAL: slot base
    int32_t out_space, out_off;
    int32_t present;           // bit k: slot k holds an input
    int32_t tile_col[SBN_CHAIN_SLOTS];  // column of slot k's offset inside a tile_off row
    int32_t T;                 // tile edge (2..5)
    int32_t cx;                // joint states of the eliminated variables (1 = product only)
    int32_t n_in, n_tiles, c0;
    int32_t table_bytes;       // bytes staged for this step (0 = none)
    SbnChainIn in[SBN_CHAIN_SLOTS];
};

struct SbnChainParams {
    const SbnChainStep *steps;
    int32_t n_steps;
    int32_t n_ev;
    const uint8_t *ev;
    int64_t ld_ev;
    int64_t ld;                // row pitch of the slot arena
    int32_t n_rows;
    int32_t n_rblocks;
    float *scratch;            // [gridDim.x][scratch_floats]
    int64_t scratch_floats;
    int32_t ev_bytes;          // bytes of the evidence-code block at the start of dynamic shared memory
    int32_t tab_floats;        // floats of ONE table buffer (two follow the codes)
    // fused normalisation (post_space < 0: the segment does not end in the posterior)
    int32_t post_space, post_off;
    int32_t Q;
    float min_total;
    float *out;
    int64_t ld_out;
    float *totals;
};

// ------------------------------------------------------------------ host side
struct sbn_program;

struct SbnSegment {
    int first = 0, last = 0;               // step indices in sbn_program::steps, inclusive
    std::vector<int> steps;                // the batched steps of the run, in order
    std::vector<SbnChainStep> host;        // descriptors (pointers patched by sbn_chain_bind)
    SbnChainStep *d_steps = nullptr;
    int64_t arena_floats = 0;              // shared-memory arena
    int64_t tab_floats = 0;                // one table buffer
    int64_t scratch_floats = 0;            // per CTA
    int threads = 0;
    size_t smem_bytes = 0;
    bool ends_in_posterior = false;
    int post_space = -1, post_off = 0;
    int64_t hbm_bytes_per_row = 0;         // what the segment still moves through the slot arena
};

// After plan_tiles: partition the batched steps into segments (fills sbn_program::segments and
// seg_first); a step outside every segment keeps its classic launch.  `tile_words` is the host
// copy of d_tile_off (offsets only; device pointers are patched in sbn_chain_bind).
void sbn_chain_plan(sbn_program *P);
// After the slot arena exists (sbn_program_reserve): patch pointers, upload, size the scratch.
cudaError_t sbn_chain_bind(sbn_program *P);
cudaError_t sbn_chain_launch(sbn_program *P, const SbnSegment &seg, const uint8_t *d_ev, int64_t ld_ev, int64_t n_rows,
                             float *d_out, int64_t ld_out, cudaStream_t stream);
void sbn_chain_free(sbn_program *P);
cudaError_t sbn_chain_set_attrs();
