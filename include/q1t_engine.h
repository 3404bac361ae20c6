/*
 * q1t_engine.h -- inner C ABI of the B200 statevector engine.
 *
 * This is the drop-in seam behind q1tsim's `VectorState`: one entry point per
 * method of the reference's `trait QuState` (src/qustate.rs:5-89) plus the two
 * constructors of `VectorState` (src/vectorstate.rs:41-83) and a few accessors
 * the reference keeps private (test hooks).  A Rust `CudaVectorState` that
 * implements `QuState` by forwarding to these functions is shown in
 * INTEGRATION.md; the host side above this ABI in this repository is C++
 * (q1tsim_b200/csrc) because no Rust toolchain exists in the build image.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types.
 *  - complex numbers are interleaved (re, im) doubles; gate matrices are
 *    row-major 2^k x 2^k with `bits[0]` the most significant bit of the
 *    matrix index (gates.rs:53-80), exactly what `Gate::matrix()` returns.
 *  - qubit 0 is the most significant bit of the amplitude index
 *    (vectorstate.rs:249-250); classical bit i is bit i of the u64 word.
 *  - every function returns Q1T_OK or a negative error code; the message
 *    (same text as the reference's `Display for Error`, error.rs:192-255) is
 *    available from q1t_last_error().
 *  - random numbers come from the caller: `q1t_rng` is the C mirror of the
 *    reference's `R: rand::Rng` argument (a Rust shim passes a trampoline that
 *    calls `rng.next_u64()`), so the caller's generator is consumed in the
 *    same order as in the reference: one Binomial per column in column order
 *    (vectorstate.rs:263-275), one Uniform(0,total) draw per shot in column
 *    order (vectorstate.rs:120-133).
 *  - the engine is asynchronous internally (gates are queued and fused); every
 *    call that returns data is synchronous.  One caller thread per state.
 */
#ifndef Q1T_ENGINE_H
#define Q1T_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct q1t_state q1t_state;

typedef uint64_t (*q1t_next_u64_fn)(void *ctx);
typedef struct {
    q1t_next_u64_fn next_u64;
    void *ctx;
} q1t_rng;

/* error codes (error.rs:134-180 variants that VectorState can return) */
#define Q1T_OK 0
#define Q1T_ERR_INVALID_NR_BITS (-1)              /* Error::InvalidNrBits */
#define Q1T_ERR_INVALID_QBIT (-2)                 /* Error::InvalidQBit */
#define Q1T_ERR_NOT_ENOUGH_SPACE (-3)             /* Error::NotEnoughSpace */
#define Q1T_ERR_INVALID_NR_MEASUREMENT_BITS (-4)  /* Error::InvalidNrMeasurementBits */
#define Q1T_ERR_INVALID_NR_CONTROL_BITS (-5)      /* Error::InvalidNrControlBits */
#define Q1T_ERR_RNG (-6)                          /* injected generator ran dry */
#define Q1T_ERR_CUDA (-7)                         /* no device / CUDA failure / out of device memory */
#define Q1T_ERR_INVALID_ARGUMENT (-8)             /* NULL pointer, duplicate qubit in `bits`, ... */
#define Q1T_ERR_UNSUPPORTED (-9)
/* circuit-level variants (used by the outer ABI, include/q1tsim_ffi.h) */
#define Q1T_ERR_INVALID_CBIT (-10)                /* Error::InvalidCBit */
#define Q1T_ERR_NOT_EXECUTED (-11)                /* Error::NotExecuted */
#define Q1T_ERR_EXPORT (-13)                      /* Error::ExportError */
#define Q1T_ERR_PARSE (-12)                       /* Error::ParseError(UnknownGate | InvalidNrArguments) */

/* ---- construction: VectorState::new / from_qubit_coefs (vectorstate.rs:41-83) ---- */
int  q1t_state_new(size_t nr_bits, size_t nr_shots, int device, q1t_state **out);
int  q1t_state_from_qubit_coefs(const double *bit_coefs_re_im /* 2*nr_bits complex */, size_t nr_bits,
                                size_t nr_shots, int device, q1t_state **out);
void q1t_state_free(q1t_state *st);

/* ---- trait QuState (qustate.rs:5-89) ---- */
/* apply_gate (vectorstate.rs:166-178).  `desc` = gate.description(), used in error text. */
int q1t_apply_gate(q1t_state *st, const double *matrix, size_t matrix_dim, const size_t *bits, size_t nr_bits,
                   const char *desc);
/* apply_unary_gate_all (vectorstate.rs:180-189) */
/* a batch of q1t_apply_gate calls in one crossing of the boundary: matrices concatenated (2 * dim^2 doubles each,
 * row-major, re/im interleaved), bit lists concatenated; stops at the first error */
int q1t_apply_gates(q1t_state *st, size_t nr_gates, const double *matrices, const size_t *matrix_dims, const size_t *bits,
                    const size_t *nr_bits_per_gate);
int q1t_apply_unary_gate_all(q1t_state *st, const double *matrix, size_t matrix_dim, const char *desc);
/* apply_conditional_gate (vectorstate.rs:193-227); control = one byte per shot */
int q1t_apply_conditional_gate(q1t_state *st, const uint8_t *control, size_t nr_control, const double *matrix,
                               size_t matrix_dim, const size_t *bits, size_t nr_bits, const char *desc);
/* measure (vectorstate.rs:229-235): res must hold nr_shots words, is zeroed first */
int q1t_measure(q1t_state *st, size_t qbit, uint64_t *res, size_t res_len, q1t_rng rng);
/* measure_into (vectorstate.rs:237-329) */
int q1t_measure_into(q1t_state *st, size_t qbit, size_t cbit, uint64_t *res, size_t res_len, q1t_rng rng);
/* measure_all (vectorstate.rs:331-338) */
int q1t_measure_all(q1t_state *st, uint64_t *res, size_t res_len, q1t_rng rng);
/* measure_all_into (vectorstate.rs:340-344) */
int q1t_measure_all_into(q1t_state *st, const size_t *cbits, size_t nr_cbits, uint64_t *res, size_t res_len,
                         q1t_rng rng);
/* peek_into (vectorstate.rs:346-393) */
int q1t_peek_into(q1t_state *st, size_t qbit, size_t cbit, uint64_t *res, size_t res_len, q1t_rng rng);
/* peek_all_into (vectorstate.rs:395-400) */
int q1t_peek_all_into(q1t_state *st, const size_t *cbits, size_t nr_cbits, uint64_t *res, size_t res_len,
                      q1t_rng rng);
/* reset (vectorstate.rs:402-408) */
int q1t_reset(q1t_state *st, size_t bit, q1t_rng rng);
/* reset_all (vectorstate.rs:410-415) */
int q1t_reset_all(q1t_state *st);

/* ---- accessors (the reference keeps these fields private, vectorstate.rs:27-34) ---- */
size_t q1t_nr_bits(const q1t_state *st);
size_t q1t_nr_shots(const q1t_state *st);
size_t q1t_nr_columns(q1t_state *st);                 /* states.cols() */
int    q1t_counts(q1t_state *st, size_t *counts_out /* nr_columns */);
/* states[(offset..offset+len, col)] in amplitude-index order, qubit 0 = MSB */
int    q1t_read_amplitudes(q1t_state *st, size_t col, size_t offset, size_t len, double *out_re_im);
int    q1t_write_amplitudes(q1t_state *st, size_t col, size_t offset, size_t len, const double *in_re_im);
/* canonical-order reductions: w0 of vectorstate.rs:252-261 and column norms */
int    q1t_marginal0(q1t_state *st, size_t qbit, double *w0_out /* nr_columns */);
int    q1t_column_totals(q1t_state *st, double *totals_out /* nr_columns */);
/* run all queued (fused) gate work and wait for the device */
int    q1t_flush(q1t_state *st);
const char *q1t_last_error(const q1t_state *st);     /* st may be NULL: last constructor error */

/* ---- shard primitives: building blocks of the multi-GPU composition (one process per GPU,
 * rank r holds the amplitudes whose top log2(P) index bits equal r).  They expose
 * the pieces of measure_into / measure_all_into (vectorstate.rs:106-161, 237-329) so that the
 * host can chain the per-rank canonical totals in rank order and keep one generator stream. ---- */
int    q1t_state_new_empty(size_t nr_bits, size_t nr_shots, int device, q1t_state **out);   /* one all-zero column */
size_t q1t_nr_leaves(const q1t_state *st);            /* canonical leaves per column: 2^nr_bits / min(2^nr_bits, 1024) */
/* canonical leaf totals of |amp|^2 per column (qbit < nr_bits: only amplitudes with that qubit == 0;
 * SIZE_MAX: all).  out holds nr_columns * nr_leaves doubles. */
int    q1t_leaf_totals(q1t_state *st, size_t qbit, double *out);
/* resolve sorted draws of column col against caller-supplied inclusive leaf prefixes P[nr_leaves];
 * `base` = total weight before this shard (0 on one GPU) */
int    q1t_resolve_draws(q1t_state *st, size_t col, const double *P, double base, const double *chosen,
                         size_t nr_draws, uint64_t *idx_out);
/* collapse on a qubit that is a rank bit: every column is scaled by f0[c] (n0 == count), f1[c]
 * (n0 == 0) or split into (f0-scaled, f1-scaled) copies; a factor 0 zeroes this rank's shard */
int    q1t_scale_split_columns(q1t_state *st, const double *f0, const double *f1, const size_t *n0);
/* collapse/renormalise/branch with (w0, n0) decided by the caller (vectorstate.rs:277-326) */
int    q1t_collapse_columns(q1t_state *st, size_t qbit, const double *w0, const size_t *n0);
/* replace all columns by basis states |idx[k]> (UINT64_MAX: an all-zero column) with counts[k] shots */
int    q1t_replace_columns(q1t_state *st, size_t nr_columns, const uint64_t *idx, const size_t *counts);
/* device pointer of a (flushed, materialised) column: 2^nr_bits complex128, for peer exchange */
int    q1t_column_device_ptr(q1t_state *st, size_t col, void **ptr);
/* Qubit remap over NVLink peer memory (one process per GPU): q1t_ipc_export gives the 64-byte CUDA IPC
 * handle of a column; q1t_peer_swap maps the partner's handle and trades rank bit <-> local qubit
 * `local_qubit` in place: my amplitudes whose local bit differs from my rank bit `my_bit` are exchanged
 * with the partner's, each rank moving half of the pairs with one remote read + one remote write per
 * amplitude, no staging buffer.  Both ranks must call it between two barriers. */
int    q1t_ipc_export(q1t_state *st, size_t col, unsigned char *handle64);
int    q1t_peer_swap(q1t_state *st, size_t col, const unsigned char *peer_handle64, size_t local_qubit, int my_bit);
/* Peer group: the stream-ordered form of the qubit remap (no host synchronisation, no allocation, no handle
 * exchange per remap).  q1t_group_export registers the two shard buffers a one-column state alternates between
 * and a mailbox, and returns their CUDA IPC handles (3 x 64 bytes: buffer 0, buffer 1 -- all zero when only one
 * fits --, mailbox) and device pointers.  q1t_group_open maps the peers' once: `all_handles` = nranks x 3 x 64
 * bytes in rank order (ranks in other processes), or `all_ptrs` = nranks x 3 device pointers (ranks that are
 * states of this process, one per device or all on one device).  q1t_group_barrier enqueues a device-side
 * barrier over the group; q1t_group_remap enqueues barrier + swap + barrier: rank bit rank_bits[j] and local
 * qubit local_qubits[j] (j < k <= 4) trade places in one in-place pass, every rank exchanging with its 2^k - 1
 * partners at once.  All ranks must make the same calls in the same order. */
/* Shards of >= 2^20 amplitudes: q1t_block_totals leaves the canonical leaf totals and their in-block inclusive
 * prefixes on the device and returns only the block totals (nr_columns x nr_leaves/1024 doubles; qbit as for
 * q1t_leaf_totals); the host continues the chain over blocks in rank order and hands every column's slice back to
 * q1t_resolve_draws_blocks: block_prefix[0] = weight in front of this shard, block_prefix[b + 1] = global inclusive
 * prefix through this shard's block b.  Same fixed geometry as the single-GPU scan. */
int    q1t_block_totals(q1t_state *st, size_t qbit, double *out);
/* the same in two halves (everything enqueued / copy to the host and wait): a host layer draws and sorts its random
 * numbers, or launches on other shards, in between */
int    q1t_block_totals_launch(q1t_state *st, size_t qbit);
int    q1t_block_totals_fetch(q1t_state *st, double *out);
int    q1t_resolve_draws_blocks(q1t_state *st, size_t col, const double *block_prefix, const double *chosen, size_t nd, uint64_t *idx);
/* every column times the scalar re + i*im (a rank's share of a one-qubit gate on a rank bit that is still
 * pinned to a basis value); real factors are deferred into the next fused sweep like the Hadamard normalisations */
int    q1t_scale(q1t_state *st, double re, double im);
/* VectorState::from_qubit_coefs (vectorstate.rs:62-83) into an existing state: 2 * nr_bits complex coefficients */
int    q1t_set_product_state(q1t_state *st, const double *coefs);
int    q1t_group_export(q1t_state *st, unsigned char *handles3x64, void **ptrs3);
int    q1t_group_open(q1t_state *st, size_t nranks, size_t rank, const unsigned char *all_handles, void *const *all_ptrs);
int    q1t_group_barrier(q1t_state *st);
int    q1t_group_remap(q1t_state *st, size_t k, const int *rank_bits, const size_t *local_qubits);
int    q1t_group_close(q1t_state *st);
/* ---- a state sharded over the devices of ONE process (csrc/sharded.h): the QuState calls a circuit beyond one GPU needs ----
 * `devices`: nr_devices = 2^g entries, they may repeat (several shards on one GPU).  Shard r holds the amplitudes whose
 * top g index bits equal r.  Starts as |0..0> (VectorState::new, vectorstate.rs:41-53).  q1t_sharded_set_initial_layout
 * (only on the fresh state): dest[q] = the qubit the data labelled q ends as after the Swap relabels of the run to come;
 * the run then ends in the canonical layout (|0..0> is symmetric, any layout is a legal start). */
typedef struct q1t_sharded q1t_sharded;
int    q1t_sharded_new(size_t nr_bits, size_t nr_shots, size_t nr_devices, const int *devices, q1t_sharded **out);
void   q1t_sharded_free(q1t_sharded *h);
int    q1t_sharded_apply_gate(q1t_sharded *h, const double *matrix, size_t matrix_dim, const size_t *bits, size_t nr_bits, const char *desc);
int    q1t_sharded_set_initial_layout(q1t_sharded *h, const int *dest, size_t n);
int    q1t_sharded_measure_all_into(q1t_sharded *h, const size_t *cbits, size_t n, uint64_t *res, size_t res_len, q1t_rng rng);
int    q1t_sharded_peek_all_into(q1t_sharded *h, const size_t *cbits, size_t n, uint64_t *res, size_t res_len, q1t_rng rng);
int    q1t_sharded_reset_all(q1t_sharded *h);
int    q1t_sharded_read_amplitudes(q1t_sharded *h, size_t offset, size_t len, double *out);   /* canonical index order */
int    q1t_sharded_column_total(q1t_sharded *h, double *out);
int    q1t_sharded_counters(q1t_sharded *h, uint64_t *out3);       /* remaps, exchanged qubits, local relabels */
const char *q1t_sharded_last_error(q1t_sharded *h);               /* h = NULL: error of the last failed q1t_sharded_new on this thread */
/* rand 0.7 Uniform(0,total) draws as WeightedIndex::sample makes them (vectorstate.rs:126) */
double q1t_uniform_draw(q1t_rng rng, double total);
void   q1t_uniform_draws(q1t_rng rng, double total, size_t n, double *out);
/* q1t_uniform_draws in two halves: unit-interval values (one generator word each), and later the map of Uniform(0, total) */
void   q1t_uniform_units(q1t_rng rng, size_t n, double *out);
void   q1t_uniform_scale(double total, size_t n, double *inout);

/* execution statistics since creation / last reset of the counters */
typedef struct {
    uint64_t gates_queued;        /* logical gates received */
    uint64_t sweeps;              /* fused state sweeps (each amplitude of each column read+written once) */
    uint64_t sweep_column_passes; /* sum over sweeps of columns touched */
    uint64_t read_passes;         /* read-only reduction passes (marginals, scans), per column */
    uint64_t kernel_launches;     /* kernels launched by this state */
    uint64_t permute_sweeps;      /* sweeps spent only on qubit relabelling */
    uint64_t fallback_sweeps;     /* gates executed by the unfused generic kernel */
    uint64_t fused_relabels;      /* relabellings absorbed into the last gate sweep (no extra pass) */
    uint64_t sweep_bytes;         /* bytes the sweep launches had to move: 32 B per amplitude, less when the input is a basis state (support tracking) */
    double   sweep_ms;            /* device time of sweep kernels (CUDA events), if timing enabled */
    double   read_ms;             /* device time of read passes */
    double   peer_swap_ms;        /* device time of q1t_peer_swap kernels (always measured) */
    uint64_t peer_swap_bytes;     /* bytes this rank moved over NVLink in q1t_peer_swap (remote reads + remote writes) */
    uint64_t plan_cache_hits;     /* gate batches whose sweep plan came from the process-wide plan cache */
    uint64_t tma_sweeps;          /* sweeps whose tiles were loaded by TMA (cp.async.bulk.tensor), dense ladder sweeps */
    uint64_t graph_captures;      /* launch-bound sweep batches captured into a CUDA graph */
    uint64_t graph_replays;       /* ... and batches executed by replaying a cached graph (one launch for the whole batch) */
    uint64_t h2d_bytes;           /* bytes copied host -> device by this state (programs, tables, draws, amplitudes written) */
    uint64_t d2h_bytes;           /* bytes copied device -> host (totals, sampled indices, amplitudes read) */
    uint64_t fused_remaps;        /* qubit remaps run as the tile loads of the sweep that followed (option "fused_remap"): no swap pass */
} q1t_stats;
int q1t_get_stats(q1t_state *st, q1t_stats *out);
int q1t_reset_stats(q1t_state *st);
/* enable per-kernel CUDA-event timing (bench only; serialises the stream) */
int q1t_set_timing(q1t_state *st, int enabled);
/* engine knobs: "tile_bits" (8..13), "fuse" (0/1), "coalesce_bits" (2/3), "balance" (-1/0/1), "track_support" (0/1), "tma" (0/1), "graphs" (0/1),
 * "inplace_relabel" (-1 never, 0 only when no second column buffer fits into device memory, 1 always),
 * "mid_relabel" (0 off, 1 default: dense batches on >= 2^24 amplitudes may store relabelled in the middle of a plan -- contiguous
 * tiles and the next targets in the coalescing positions, DESIGN.md 4.5c -- when a second column buffer fits, 2: at every size),
 * "fused_remap" (0/1: q1t_group_remap only records the trade; the next dense ladder sweep reads its tiles from the peers' shards over
 * NVLink and writes into this rank's other registered buffer.  For host layers with one thread per shard: the deferred barriers are
 * launched from inside the next flush).
 * Returns Q1T_ERR_INVALID_ARGUMENT for unknown keys. */
int q1t_set_option(q1t_state *st, const char *key, long value);

/* ---- host-only helpers (no device needed) ---- */
/* built-in generators usable as q1t_rng: SplitMix64(seed) and an injected array of raw words */
typedef struct q1t_rng_state q1t_rng_state;
q1t_rng_state *q1t_rng_splitmix64(uint64_t seed);
q1t_rng_state *q1t_rng_from_words(const uint64_t *words, size_t n);   /* copies the words */
q1t_rng_state *q1t_rng_entropy(void);                                /* like rand::thread_rng() */
size_t         q1t_rng_consumed(const q1t_rng_state *r);
void           q1t_rng_free(q1t_rng_state *r);
q1t_rng        q1t_rng_handle(q1t_rng_state *r);
/* rand_distr 0.2 Binomial restated (vectorstate.rs:271-272) */
uint64_t q1t_binomial(q1t_rng rng, uint64_t n, double p);
/* `matrix()` of a built-in gate by name (composite.rs:287-445 table); returns #qubits or <0 */
int q1t_gate_matrix(const char *name, const double *params, size_t nr_params, double *out_re_im /* 2*64 */);
/* Composite::from_string(..).matrix() (composite.rs:273-450, :480-485): parses the description and
 * writes the row-major 2^k x 2^k matrix (re,im).  Returns k (<= 10), or Q1T_ERR_PARSE with the
 * reference's ParseError text in err_out, or Q1T_ERR_NOT_ENOUGH_SPACE if out_capacity_doubles < 2*4^k. */
int q1t_composite_matrix(const char *description, double *out_re_im, size_t out_capacity_doubles, char *err_out, size_t err_capacity);
/* Expression::parse(text).eval() (expression.rs:86-392): arithmetic with pi, + - * / ^, sin cos tan exp ln sqrt.
 * Returns 0 and the value, or Q1T_ERR_PARSE; *consumed = characters parsed. */
int q1t_eval_expression(const char *text, double *value_out, size_t *consumed, char *err_out, size_t err_capacity);
/* fusion planner dry run (no device): how many sweeps / rounds a gate list takes on nr_bits qubits.
 * gates: concatenated matrices, bits; returns Q1T_OK and fills out[0]=sweeps out[1]=rounds out[2]=ops
 * out[3]=fallback gates out[4]=permute sweeps needed to restore canonical order */
int q1t_plan_dry_run(size_t nr_bits, size_t nr_gates, const double *matrices, const size_t *matrix_dims,
                     const size_t *bits, const size_t *nr_gate_bits, long tile_bits, uint64_t *out);
/* in-place relabelling dry run (no device): the passes that restore canonical order when no second column
 * buffer fits.  dstpos[p] = destination position of index bit p.  Writes, per pass, tile_bits
 * tile positions into out_tiles and nr_bits destination positions into out_dstpos; returns the number of
 * passes (0 for the identity), Q1T_ERR_NOT_ENOUGH_SPACE if more than max_passes are needed. */
int q1t_plan_inplace_relabel(size_t nr_bits, long tile_bits, long coalesce_bits, const int *dstpos,
                             int *out_tiles, int *out_dstpos, size_t max_passes);
/* test hooks (no device): the sweep programs the planner produces for a fusable gate list, as the raw structs of
 * csrc/program.h (SweepProgram x max_sweeps, PhaseTab x max_ptabs, ptab_counts[i] = tables of sweep i) and the final
 * logical->physical bit map; returns the number of sweeps.  tests/plan_interpreter.py executes them on the CPU.
 * balance: bit 0 = the planner's balanced packing; bits 4.. = how the Swap relabelling is undone at the end (0 not at
 * all, 1 as the engine does out of place: fused into the last sweep or one relabel sweep, 2 the in-place passes).
 * q1t_plan_layout: struct sizes and constants for the reader. */
int q1t_plan_dump(size_t nr_bits, size_t nr_gates, const double *matrices, const size_t *matrix_dims, const size_t *bits,
                  const size_t *nr_gate_bits, long tile_bits, long coalesce_bits, int balance,
                  void *progs_out, size_t max_sweeps, void *ptabs_out, size_t max_ptabs, int *ptab_counts, int *perm_out);
int q1t_plan_layout(size_t *out, size_t n);
int q1t_device_count(void);
const char *q1t_version(void);

#ifdef __cplusplus
}
#endif
#endif
