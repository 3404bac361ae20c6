/*
 * q1t_oracle.h -- CPU ORACLE for the q1tsim statevector hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * the timed CPU baseline.  The product (q1tsim_b200/) never links or calls it.
 *
 * This is a plain-C restatement of the reference algorithm (Q1tBV/q1tsim
 * 0.5.0, Rust).  The reference cannot be compiled here (no rustc/cargo, no
 * vendored crates), so every function cites the reference file:line it
 * follows.
 *
 * Parity status:
 *  - gate application, conditional column splitting, collapse, reset,
 *    bit routing of measure_all, histogram:  PINNED against the reference's
 *    own deterministic unit tests (tests/test_oracle_reference_kats.py
 *    transcribes vectorstate.rs:425-509,640-707,773-830, circuit.rs:1456-1468,
 *    1619-1673,1688-1722,1924-1985, support.rs:99-118, stats.rs:39-59).
 *  - random sampling (rand 0.7 WeightedIndex / Uniform, rand_distr 0.2
 *    Binomial: un-vendored third-party crates, restated from their published
 *    algorithms): PARITY UNPINNED -- the reference holds no test that pins a
 *    sampled sequence (all use thread_rng + statistical bounds).
 */
#ifndef Q1T_ORACLE_H
#define Q1T_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } orc_cplx;

/* ---- random numbers ------------------------------------------------- */
/* kind 0: SplitMix64(seed)   kind 1: injected array of raw u64 words     */
typedef struct {
    int kind;
    uint64_t s;
    const uint64_t *arr;
    size_t n, pos;
    int exhausted;
} orc_rng;

void     orc_rng_seed(orc_rng *r, uint64_t seed);
void     orc_rng_array(orc_rng *r, const uint64_t *arr, size_t n);
uint64_t orc_rng_next(orc_rng *r);
uint64_t orc_binomial(orc_rng *r, uint64_t n, double p);

/* ---- state ------------------------------------------------------------ */
typedef struct {
    size_t nr_bits, nr_shots;
    size_t ncols;
    size_t *counts;          /* ncols */
    orc_cplx *states;        /* (2^nr_bits, ncols) row-major, as vectorstate.rs:34 */
} orc_state;

orc_state *orc_state_new(size_t nr_bits, size_t nr_shots);
orc_state *orc_state_from_qubit_coefs(const double *coefs_re_im, size_t nr_bits, size_t nr_shots);
void       orc_state_free(orc_state *s);
size_t     orc_state_ncols(const orc_state *s);
void       orc_state_counts(const orc_state *s, size_t *out);
void       orc_state_read_column(const orc_state *s, size_t col, double *out_re_im);
void       orc_state_write_column(orc_state *s, size_t col, const double *in_re_im);

/* mode: 0 = faithful (reference loop structure, temporaries, materialised
 *           permutation -- the timed CPU baseline),
 *       1 = fast (same arithmetic, direct strided loops, optional OpenMP) */
void orc_set_threads(int nthreads);
int  orc_apply_gate(orc_state *s, const double *mat, const size_t *bits, size_t k, int mode);
int  orc_apply_unary_gate_all(orc_state *s, const double *mat, int mode);
int  orc_apply_conditional_gate(orc_state *s, const uint8_t *control, size_t ncontrol,
                                const double *mat, const size_t *bits, size_t k, int mode);

/* order: 0 = reference summation order (sequential), 1 = canonical blocked
 * order shared with the GPU engine (DESIGN.md "canonical reduction order") */
int  orc_marginal0(const orc_state *s, size_t qbit, int order, double *w0_out);
int  orc_measure_into(orc_state *s, size_t qbit, size_t cbit, uint64_t *res, size_t res_len,
                      orc_rng *rng, int order);
int  orc_peek_into(const orc_state *s, size_t qbit, size_t cbit, uint64_t *res, size_t res_len,
                   orc_rng *rng, int order);
int  orc_measure_all_into(orc_state *s, const size_t *cbits, size_t ncbits, uint64_t *res,
                          size_t res_len, int collapse, orc_rng *rng, int order);
int  orc_reset(orc_state *s, size_t bit, orc_rng *rng, int order, int mode);
void orc_reset_all(orc_state *s);
void orc_column_totals(const orc_state *s, int order, double *out);

/* helpers restated from support.rs / qustate.rs / gates.rs */
uint64_t orc_reverse_bits(uint64_t idx, size_t nr_bits);
uint64_t orc_shuffle_bits(uint64_t idx, const size_t *bits, size_t n);
size_t   orc_collect_conditional_ranges(const size_t *counts, size_t ncols, const uint8_t *control,
                                        size_t *out_icol, size_t *out_len, uint8_t *out_apply);
int      orc_bit_permutation(size_t nr_bits, const size_t *bits, size_t k, size_t *perm_out);

/* gate matrices by (lower-case) name, composite.rs:287-445 name table.
 * returns number of qubits, or -1 unknown name, -2 wrong number of params */
int orc_gate_matrix(const char *name, const double *params, size_t nparams, double *out_re_im);

/* error codes */
#define ORC_OK 0
#define ORC_ERR_INVALID_NR_BITS (-1)
#define ORC_ERR_INVALID_QBIT (-2)
#define ORC_ERR_NOT_ENOUGH_SPACE (-3)
#define ORC_ERR_INVALID_NR_MEASUREMENT_BITS (-4)
#define ORC_ERR_INVALID_NR_CONTROL_BITS (-5)
#define ORC_ERR_RNG_EXHAUSTED (-6)
#define ORC_ERR_OUT_OF_MEMORY (-7)     /* the reference's dense collapse matrix (vectorstate.rs:150-158) does not fit into host memory */

#ifdef __cplusplus
}
#endif
#endif
