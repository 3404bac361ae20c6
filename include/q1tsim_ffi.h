/*
 * q1tsim_ffi.h -- the OUTER C ABI: the symbols of the reference's src/ffi.rs
 * (authoritative C declaration: python/q1tsimffi.py:35-82), re-implemented over
 * the B200 engine so that python/q1tsim.py runs unchanged against
 * q1tsim_b200/lib/libq1tsim.so.  Differences from the reference, all additive:
 *   - the gate-name table is the superset parsed by Composite::from_string
 *     (composite.rs:287-445: cs, ct, cu1, ccx, ... are reachable; ffi.rs:335-367
 *     accepts only 25 names, so the README QFT cannot be expressed there);
 *   - circuit_execute always runs the statevector backend on the GPU (the
 *     reference switches to its stabilizer backend for all-Clifford circuits,
 *     circuit.rs:576-583, which is out of scope here);
 *   - circuit_open_qasm / circuit_c_qasm / circuit_latex produce the reference's text (circuit.rs:877-1231);
 *   - extra entry points: circuit_add_matrix_gate, circuit_add_composite_gate, circuit_add_loop_gate,
 *     circuit_execute_with_rng,
 *     circuit_reexecute_with_rng, circuit_histogram_u64, circuit_engine_stats,
 *     circuit_set_device, circuit_set_devices (+ circuit_sharded_*), circuit_state.
 * Ownership (ffi.rs:139-169): every result_t is returned by value and owns
 * `data`; the caller passes it to result_free exactly once.
 */
#ifndef Q1TSIM_FFI_H
#define Q1TSIM_FFI_H

#include <stddef.h>
#include <stdint.h>

#include "q1t_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct circuit circuit_t;
typedef struct { const char *key; size_t count; } histelem_t;     /* ffi.rs:24-29 CHistElem */
typedef struct { double value; double *value_ptr; } parameter_t;   /* ffi.rs:40-46 CParameter */
typedef struct { void *data; size_t length; size_t size; uint32_t restype; } result_t;   /* ffi.rs:63-70 CResult */

#define RESULT_ERROR 0u      /* ffi.rs:5-9 */
#define RESULT_EMPTY 1u
#define RESULT_STRING 2u
#define RESULT_HISTOGRAM 3u
#define RESULT_CSTATE 5u

void       result_free(result_t res);                                        /* ffi.rs:165-169 */
circuit_t *circuit_new(size_t nr_qbits, size_t nr_cbits);                    /* ffi.rs:171-176 */
void       circuit_free(circuit_t *ptr);                                     /* ffi.rs:178-186 */
size_t     circuit_nr_qbits(const circuit_t *ptr);                           /* ffi.rs:188-194 */
size_t     circuit_nr_cbits(const circuit_t *ptr);                           /* ffi.rs:196-202 */
result_t   circuit_cstate(const circuit_t *ptr);                             /* ffi.rs:204-214 */
result_t   circuit_add_gate(circuit_t *ptr, const char *gate, const size_t *qbits, size_t nr_qbits,
                            const parameter_t *param_ptr, size_t nr_params);  /* ffi.rs:307-380 */
result_t   circuit_add_conditional_gate(circuit_t *ptr, const size_t *control_ptr, size_t nr_control,
                                        uint64_t target, const char *gate, const size_t *qbits_ptr, size_t nr_qbits,
                                        const parameter_t *param_ptr, size_t nr_params);   /* ffi.rs:382-459 */
result_t   circuit_measure(circuit_t *ptr, size_t qbit, size_t cbit, char dir, uint8_t collapse);        /* ffi.rs:494-522 */
result_t   circuit_measure_all(circuit_t *ptr, const size_t *cbits, size_t nr_cbits, char dir, uint8_t collapse); /* ffi.rs:524-557 */
result_t   circuit_reset(circuit_t *ptr, size_t qbit);                       /* ffi.rs:461-477 */
result_t   circuit_reset_all(circuit_t *ptr);                                /* ffi.rs:479-492 */
result_t   circuit_execute(circuit_t *ptr, size_t nr_shots);                 /* ffi.rs:559-575 */
result_t   circuit_reexecute(circuit_t *ptr);                                /* ffi.rs:577-593 */
result_t   circuit_histogram(const circuit_t *ptr);                          /* ffi.rs:595-611 */
result_t   circuit_latex(const circuit_t *ptr);                              /* ffi.rs:613-629; circuit.rs:1148-1231 */
result_t   circuit_open_qasm(const circuit_t *ptr);                          /* ffi.rs:632-648; circuit.rs:877-1017 */
result_t   circuit_c_qasm(const circuit_t *ptr);                             /* ffi.rs:650-666; circuit.rs:1019-1146 */

/* ---- additive entry points ---- */
/* arbitrary user gate given by its matrix() (gates.rs:174), row-major (re,im) */
result_t   circuit_add_matrix_gate(circuit_t *ptr, const char *description, const double *matrix_re_im,
                                   size_t matrix_dim, const size_t *qbits, size_t nr_qbits);
result_t   circuit_add_conditional_matrix_gate(circuit_t *ptr, const size_t *control_ptr, size_t nr_control,
                                               uint64_t target, const char *description, const double *matrix_re_im,
                                               size_t matrix_dim, const size_t *qbits, size_t nr_qbits);
result_t   circuit_barrier(circuit_t *ptr, const size_t *qbits, size_t nr_qbits);           /* circuit.rs:541-552 */
/* Composite::from_string(name, desc) (composite.rs:273-450) added on `qbits`, its body `nr_iterations`
 * times (Loop, staticloop.rs:71-92; 1 for a plain composite).  The sub-gates are flattened into the
 * circuit rather than multiplied into one 2^k x 2^k matrix.  Errors: the reference's
 * ParseError texts (error.rs:93-127) and "Expected {} bits for \"{}\", got {}". */
result_t   circuit_add_composite_gate(circuit_t *ptr, const char *name, const char *description,
                                      const size_t *qbits, size_t nr_qbits, size_t nr_iterations);
/* Loop::new(label, nr_iterations, Composite::from_string(label, body)) (staticloop.rs:38-50): like
 * circuit_add_composite_gate, but exported as a loop (`.label(n) ... .end` in c-Qasm) even for one
 * iteration.  A loop of 0 iterations adds nothing (the reference keeps an empty instruction). */
result_t   circuit_add_loop_gate(circuit_t *ptr, const char *label, const char *body_description,
                                 const size_t *qbits, size_t nr_qbits, size_t nr_iterations);
size_t     circuit_nr_ops(const circuit_t *ptr);
/* execute_with_rng / reexecute_with_rng (circuit.rs:573-641) with a caller-owned generator */
result_t   circuit_execute_with_rng(circuit_t *ptr, size_t nr_shots, q1t_rng rng);
result_t   circuit_reexecute_with_rng(circuit_t *ptr, q1t_rng rng);
/* execute_with (circuit.rs:594-600): start from a product state given by 2*nr_qbits complex coefficients */
result_t   circuit_execute_with_qubit_coefs(circuit_t *ptr, size_t nr_shots, q1t_rng rng, const double *coefs_re_im);
/* Circuit::histogram (circuit.rs:773-791): keys as u64; data = {uint64 key; size_t count}[length], restype 6 */
#define RESULT_HISTOGRAM_U64 6u
typedef struct { uint64_t key; size_t count; } histelem_u64_t;
result_t   circuit_histogram_u64(const circuit_t *ptr);
/* copy the classical register into a caller buffer (no allocation); returns number of words */
size_t     circuit_cstate_into(const circuit_t *ptr, uint64_t *out, size_t out_len);
/* preset the classical register (tests of circuit.rs:1628-1650 set c_state directly) */
result_t   circuit_set_cstate(circuit_t *ptr, const uint64_t *words, size_t n);
int        circuit_set_device(circuit_t *ptr, int device);
/* execute() on a state sharded over n devices of this process (power of two >= 2, devices may repeat; n = 0: back to one
 * device): gates, measure_all / peek_all in any basis, barriers -- the circuits of BASELINE config 5 (34-36 qubits on
 * 8 B200).  The reference has no analogue (its VectorState is one host array, vectorstate.rs:25-35). */
int        circuit_set_devices(circuit_t *ptr, const int *devices, size_t n);
int        circuit_sharded_amplitudes(circuit_t *ptr, size_t offset, size_t len, double *out);   /* canonical order, re/im pairs */
int        circuit_sharded_counters(circuit_t *ptr, uint64_t *out3);   /* remaps, exchanged qubits, local relabels of the last run */
q1t_state *circuit_state(circuit_t *ptr);          /* borrowed: the live q_state, NULL before execute */
int        circuit_engine_stats(circuit_t *ptr, q1t_stats *out);

#ifdef __cplusplus
}
#endif
#endif
