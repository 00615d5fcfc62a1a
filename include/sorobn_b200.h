/*
 * sorobn_b200 -- C ABI of the B200 exact-inference engine.
 *
 * This is the drop-in boundary for the exact-inference path of MaxHalford/sorobn.
 * The reference has no native layer: the whole path is Python over pandas
 * (/root/reference/sorobn/bayes_net.py).  The entry points below are what a ctypes
 * binding inside the reference's `BayesNet` would call instead of
 *
 *   - `BayesNet._variable_elimination`      bayes_net.py:739-794  (the loop)
 *   - `pointwise_mul` / `pointwise_mul_two` bayes_net.py:106-256  (factor product)
 *   - `CDTAccessor.sum_out`                 bayes_net.py:54-103   (marginalisation)
 *   - the normalisation at                  bayes_net.py:789-790
 *
 * INTEGRATION.md shows that binding.  Plain pointers and sizes only: no torch,
 * numpy or pandas types cross this line.
 *
 * A *program* is the frozen form of one `(query variables, evidence variables)`
 * pair: the CPTs involved (dense fp32 tables) and the list of fused
 * "product of k factors -> sum out one variable" steps, one kernel launch each
 * (word layout: sorobn_b200/planner.py).  Running a program on B evidence rows
 * computes B posteriors, i.e. B calls of `BayesNet.query(..., algorithm="exact")`.
 *
 * Layouts (both "column-major" over rows so that device accesses coalesce):
 *   evidence : uint8 state codes, ev[c * ld_ev + b]   c < n_ev, b < n_rows
 *   posterior: float,             out[q * ld_out + b] q < Q,    b < n_rows
 * where q enumerates the joint states of the query variables, variables sorted by
 * name and the last one varying fastest -- the row order of the reference's answer
 * (`reorder_levels(sorted(...))`, `sort_index()`; bayes_net.py:872-875).
 * A row whose evidence has probability zero yields NaN (the reference returns an
 * empty Series there).  In float32 programs a row whose normaliser is below 1e-30 is also
 * NaN: float32 underflow may have dropped addends; re-run it with a float64 program.
 *
 * Every function returns 0 on success or a negative SBN_E_* code;
 * sbn_last_error() then describes the failure (thread-local string).
 */
#ifndef SOROBN_B200_H
#define SOROBN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBN_ABI_VERSION 9

#define SBN_OK 0
#define SBN_E_INVALID (-1)   /* malformed program / bad argument            */
#define SBN_E_CUDA (-2)      /* CUDA runtime error (see sbn_last_error)     */
#define SBN_E_NOMEM (-3)     /* scratch does not fit the device             */
#define SBN_E_NODEVICE (-4)  /* no usable sm_100 GPU                        */

typedef struct sbn_program sbn_program;

/* Fixed limits of the step kernel (also enforced by the planner). */
#define SBN_MAX_IN 8     /* factors multiplied in one launch   */
#define SBN_MAX_AXES 20  /* variables in one output factor     */
#define SBN_MAX_EV 8     /* evidence axes gathered per factor  */

int sbn_abi_version(void);
const char *sbn_last_error(void);

/* Number of CUDA devices visible; SBN_E_NODEVICE if none. */
int sbn_device_count(int *count);

/* Compile a program for `device`: validates `words` (planner.py layout), uploads the
 * CPT tables.  Replaces the per-query factor preparation of bayes_net.py:768-776. */
int sbn_program_create(int device, const int32_t *words, int64_t n_words, const float *tables,
                       int64_t n_table_floats, sbn_program **out);
void sbn_program_destroy(sbn_program *prog);

/* Same for a program evaluated in float64: tables, scratch and the posterior are doubles.
 * Single-event ("flat", mode 0) programs are what `BayesNet.query` uses; batched ones re-run
 * the rows a float32 program flagged.  One query is launch-latency bound, so it gets the
 * reference's own precision and range (float64, bayes_net.py throughout) for free; it is
 * also the fallback for evidence rows too unlikely for float32 (see run_host below). */
int sbn_program_create_f64(int device, const int32_t *words, int64_t n_words, const double *tables,
                           int64_t n_table_doubles, sbn_program **out);

/* Allocate scratch for chunks of up to `max_rows` evidence rows (larger batches are
 * processed in chunks).  Called implicitly by the run functions when needed. */
int sbn_program_reserve(sbn_program *prog, int64_t max_rows);

/* Answer `n_rows` queries with HOST buffers: copies the evidence codes to the device,
 * runs every step, copies the posteriors back and synchronises.  This is the call that
 * replaces `BayesNet._variable_elimination` (bayes_net.py:739) for a batch of events. */
int sbn_program_run_host(sbn_program *prog, const uint8_t *ev, int64_t ld_ev, int64_t n_rows, float *out,
                         int64_t ld_out);

/* float64 programs: flat (n_rows must be 1) or batched (the robust fallback for rows that
 * the float32 program flagged; plain kernel in double, several times slower). */
int sbn_program_run_host_f64(sbn_program *prog, const uint8_t *ev, int64_t ld_ev, int64_t n_rows, double *out,
                             int64_t ld_out);

/* P(event) of every evidence row: the normaliser the run divides by (bayes_net.py:790).  This is
 * what `BayesNet.predict_proba` returns (bayes_net.py:934-962: the full joint, marginalised over
 * the unobserved variables and looked up at the row) without ever building the joint.  The program
 * may have no query variable at all (planner: allow_empty_query).  prob[b], b < n_rows; NaN marks
 * a row below the float32 range (re-run it with a float64 program). */
int sbn_program_evidence_host(sbn_program *prog, const uint8_t *ev, int64_t ld_ev, int64_t n_rows, float *prob);
int sbn_program_evidence_host_f64(sbn_program *prog, const uint8_t *ev, int64_t ld_ev, int64_t n_rows, double *prob);

/* Same with DEVICE buffers, asynchronous on `stream` (a cudaStream_t; NULL = default
 * stream).  n_rows must not exceed the reserved chunk size. */
int sbn_program_run_device(sbn_program *prog, const uint8_t *d_ev, int64_t ld_ev, int64_t n_rows, float *d_out,
                           int64_t ld_out, void *stream);

/* Per-step device time of one run on device buffers (CUDA events around every launch;
 * diagnostic, not the fast path).  step_ms has n_steps + 1 entries (last = normalise). */
int sbn_program_profile(sbn_program *prog, const uint8_t *d_ev, int64_t ld_ev, int64_t n_rows, float *d_out,
                        int64_t ld_out, void *stream, float *step_ms, int64_t n_step_ms);

/* info[0]=Q  [1]=n_ev  [2]=n_steps  [3]=scratch floats per row  [4]=reserved rows
 * [5]=kernel launches issued by this program so far  [6]=mode (0 flat, 1 batched)
 * [7]=unbatched scratch floats
 * with n_info >= 12 also: [8]=on-chip segments in use  [9]=steps they cover  [10]=bytes per row the
 * segments still move through the HBM slot arena  [11]=per-CTA private scratch floats
 * with n_info >= 14 also: [12]=paired launches in use (a step and its consumer as one kernel, csrc/sbn_pair.h)
 * [13]=bytes per row those pairs do not move (their intermediates stay in registers) */
int sbn_program_info(const sbn_program *prog, int64_t *info, int64_t n_info);

/* How every program step is executed with the current switches: roles[i] = 0 evidence-independent (ran once, at
 * creation; or any step of a flat program), 1 its own launch, 2 / 3 first / second step of a paired launch
 * (csrc/sbn_pair.h: table x frontier twice), 4 / 5 first / second step of an expanding product fused with its
 * consumer, 6 inside an on-chip segment.  n_roles >= n_steps. */
int sbn_program_step_roles(const sbn_program *prog, int32_t *roles, int64_t n_roles);

/* 0 = plain launches; 1 = CUDA-graph replay of the step sequence (default); 3 = graph replay
 * with independent sub-trees of the elimination as parallel branches (experimental: measured
 * no gain on the benchmark plans, which are one long dependency chain). */
int sbn_program_set_graph(sbn_program *prog, int enabled);

/* Select the step kernel: 0 = the plain one-output-per-iteration kernel (general
 * fallback, cross-check in tests); 1 or 2 = register-tiled kernel with the operand
 * preload schedule where available (default); 4 = tiled, x-loop schedule only; 5 = tiled
 * without the shared-memory slab variant for expanding products; 7 = run the on-chip segments
 * (csrc/sbn_chain.h: runs of steps executed by one persistent kernel with the intermediates in
 * shared memory / an L2-resident scratch; opt-in, also SOROBN_B200_CHAIN=1), 6 = back to one launch
 * per step; 9 = run the steps it covers through the tensor-map TMA pipeline kernel (csrc/sbn_tma.h: 2-D / 4-D
 * `cp.async.bulk.tensor` boxes into a shared-memory ring fed by a producer warp; opt-in, also SOROBN_B200_TMA=1:
 * parity-green but not faster than the register-preload kernel), 8 = off again;
 * 10 = no paired steps (every step its own launch), 11 = paired steps where eligible (default; csrc/sbn_pair.h:
 * a step and its consumer run as ONE kernel that keeps the intermediate factor in registers; SOROBN_B200_PAIR=0
 * disables them at creation). */
int sbn_program_set_tiled(sbn_program *prog, int enabled);

/* ------------------------------------------------------------------ Gibbs sampling
 * `BayesNet._gibbs_sampling` (bayes_net.py:665-737) with one chain per evidence row.
 * Variables are numbered topologically (every parent id < its child's id); CPT `v` lives at
 * tables[cpt_off[v]] with axes [*parents(v), v], v fastest.  `query` lists the query
 * variables slowest first (the order of the posterior's rows); `cycle` is the resampling
 * order of the non-event variables (the reference: sorted by name).  out[q * ld_out + c] is
 * the fraction of chain c's iterations spent in joint query state q. */
typedef struct sbn_sampler sbn_sampler;
int sbn_gibbs_create(int device, int32_t n_vars, const int32_t *card, const int32_t *par_ptr, const int32_t *par_idx,
                     const int32_t *cpt_off, const float *tables, int64_t n_table_floats, int32_t n_query,
                     const int32_t *query, int32_t n_ev, const int32_t *ev_vars, int32_t n_cycle, const int32_t *cycle,
                     sbn_sampler **out);
int sbn_gibbs_run_host(sbn_sampler *sampler, const uint8_t *ev, int64_t ld_ev, int64_t n_chains, int64_t n_iterations,
                       uint64_t seed, float *out, int64_t ld_out);
void sbn_gibbs_destroy(sbn_sampler *sampler);

/* The conditional the chain resamples `var` from, P(var | Markov blanket) for ONE joint state
 * (joint[v] = state code of variable v, v < n_vars; only the blanket is read): out[x], x < card(var),
 * normalised.  It is the table `_gibbs_sampling` precomputes for every variable
 * (bayes_net.py:699-712), evaluated by the same device code the chains run -- deterministic, so the
 * tests pin it entry by entry to the reference's tables. */
int sbn_gibbs_conditional(sbn_sampler *sampler, int32_t var, const uint8_t *joint, float *out);

/* The other sampling algorithms of `BayesNet.query` on the same sampler object:
 * algo 0 = Gibbs (as above), 1 = likelihood weighting (bayes_net.py:621-663), 2 = rejection
 * sampling (bayes_net.py:577-619); both built on forward sampling (bayes_net.py:518-548),
 * n_iterations samples per evidence row.  A row of rejection sampling that keeps no sample
 * is NaN (the reference returns an empty Series). */
#define SBN_ALGO_GIBBS 0
#define SBN_ALGO_LIKELIHOOD 1
#define SBN_ALGO_REJECTION 2
#define SBN_ALGO_GIBBS_GENERIC 3  /* Gibbs through the generic kernel even when the straight-line one applies (tests) */
int sbn_sampler_run_host(sbn_sampler *sampler, int algo, const uint8_t *ev, int64_t ld_ev, int64_t n_rows,
                         int64_t n_iterations, uint64_t seed, float *out, int64_t ld_out);

/* Pinned host memory for evidence / posterior staging buffers. */
int sbn_host_alloc(void **ptr, int64_t bytes);
int sbn_host_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* SOROBN_B200_H */
