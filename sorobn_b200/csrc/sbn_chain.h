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

// Everything the kernel adds to an address is a 32-bit BYTE offset, premultiplied on the host by
// the row pitch of the space the factor lives in (1 float for a table, 32 for shared memory and
// scratch, ld for the slot arena): the inner loop is one integer add and one load per operand.
struct SbnChainIn {
    const float *ptr;       // TABLE: source of the bulk copy; GLOBAL: slot base (row 0)
    int32_t space;
    uint32_t off;           // bytes -- TABLE: inside the step's table buffer; SMEM: inside the arena; SCRATCH: inside the CTA's scratch
    uint32_t sxb, sab, sbb; // byte strides: first eliminated variable, tile axis 0, tile axis 1
    int32_t col;            // column of this input in a tile row / row in the joint-state offset table
    int32_t stage_bytes;    // TABLE: bytes to stage (multiple of 16)
    int32_t n_ev;
    int32_t ev_col[SBN_MAX_EV];
    uint32_t ev_stride_b[SBN_MAX_EV];
    int32_t ev_card[SBN_MAX_EV];
};

// Host-side description of one step of a segment (sbn_chain_bind serialises it into the step
// records the kernel reads from shared memory and into the producer's staging list).
struct SbnChainStep {
    float *out_ptr;            // GLOBAL: slot base
    int32_t out_space;
    uint32_t out_off;          // bytes inside the arena / scratch
    uint32_t out_eb, out_c0b;  // byte stride of one output entry (= digit of axis 0) and of a digit of axis 1
    int32_t present;           // bit k: slot k holds an input
    int32_t T;                 // tile edge (2..5)
    int32_t cx;                // joint states of the eliminated variables (1 = product only)
    int32_t n_in, n_tiles;
    int32_t table_bytes;       // bytes staged for this step (0 = none)
    int32_t fast;              // 0 = generic step body; 1 + (A0 outside shared memory) + 2 * (B0 outside shared memory)
    SbnChainIn in[SBN_CHAIN_SLOTS];
};

// What the producer lane needs to stage one step's tables (bulk-TMA copies into the table ring).
struct SbnChainStage {
    uint32_t bytes;            // total bytes of the copies (0 = the step has no table)
    int32_t n;
    struct {
        const float *src;
        uint32_t dst_off;      // bytes inside the ring
        uint32_t bytes;
    } copy[SBN_CHAIN_SLOTS];
};

#define SBN_CHAIN_STAGES 4     // steps whose tables are in flight / resident at any time

struct SbnChainParams {
    const uint32_t *recs;      // step records (word layout: sbn_chain.cu), copied to shared memory at kernel start
    const SbnChainStage *stages;
    int32_t rec_words;
    int32_t n_steps;
    int32_t n_ev;
    int32_t n_rows;
    const uint8_t *ev;
    int64_t ld_ev;
    int64_t ld;                // row pitch of the slot arena
    int32_t n_rblocks;
    int32_t ev_bytes;          // dynamic shared memory: [evidence codes][step records][table ring][arena]
    int32_t ring_bytes;
    float *scratch;            // [gridDim.x][scratch_floats]
    int64_t scratch_floats;
    // fused normalisation (post_space < 0: the segment does not end in the posterior)
    int32_t post_space;
    uint32_t post_off;         // bytes
    int32_t Q;
    float min_total;
    float *out;
    int64_t ld_out;
    float *totals;
    unsigned long long *prof;  // developer aid (SOROBN_B200_CHAIN_PROF=1): per-step cycle counters, else nullptr
};

// ------------------------------------------------------------------ host side
struct sbn_program;

struct SbnChainHome {                      // where one step's operands live (decided by sbn_chain_plan)
    int step = 0;
    int out_space = SBN_SP_GLOBAL;
    int64_t out_off = 0;                   // floats
    int slot[4] = {0, 0, 0, 0};            // class slot of input i (tiled-kernel order)
    int space[4] = {0, 0, 0, 0};
    int64_t off[4] = {0, 0, 0, 0};         // floats
};

struct SbnSegment {
    int first = 0, last = 0;               // step indices in sbn_program::steps, inclusive
    std::vector<int> steps;                // the batched steps of the run, in order
    std::vector<SbnChainHome> homes;       // one per batched step
    std::vector<SbnChainStep> host;        // full descriptors (host only), built by sbn_chain_bind
    std::vector<int64_t> ring_off;         // per step: byte offset of its tables inside the table ring
    uint32_t *d_recs = nullptr;            // step records (what the compute warps read, from shared memory)
    SbnChainStage *d_stages = nullptr;     // what the producer lane reads
    uint32_t *d_words = nullptr;           // tile / joint-state offset tables in bytes
    int64_t rec_words = 0;
    int64_t ring_bytes = 0;                // table ring
    int64_t arena_floats = 0;              // shared-memory arena
    int64_t scratch_floats = 0;            // per CTA
    int threads = 0;
    size_t smem_bytes = 0;
    bool ends_in_posterior = false;
    int post_space = -1;
    int64_t post_off = 0;                  // floats
    int64_t hbm_bytes_per_row = 0;         // what the segment still moves through the slot arena
};

// After plan_tiles: partition the batched steps into segments (fills sbn_program::segments and
// seg_first); a step outside every segment keeps its classic launch.
void sbn_chain_plan(sbn_program *P);
// After the slot arena exists (sbn_program_reserve): patch pointers, upload, size the scratch.
cudaError_t sbn_chain_bind(sbn_program *P);
cudaError_t sbn_chain_launch(sbn_program *P, const SbnSegment &seg, const uint8_t *d_ev, int64_t ld_ev, int64_t n_rows,
                             float *d_out, int64_t ld_out, cudaStream_t stream);
void sbn_chain_free(sbn_program *P);
cudaError_t sbn_chain_set_attrs();
