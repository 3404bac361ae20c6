// ffi_compat.cpp -- the OUTER C ABI: same symbols, signatures, result ownership
// and error strings as the reference's src/ffi.rs, over q1t::Circuit.
// See include/q1tsim_ffi.h for the (additive) differences.
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/q1tsim_ffi.h"
#include "capi_internal.h"
#include "circuit.h"

using q1t::Basis;
using q1t::Circuit;
using q1t::CircuitError;
using q1t::GateSpec;
using q1t::Param;

struct circuit {
    Circuit impl;
    q1t_rng_state *thread_rng;       // circuit_execute uses an entropy-seeded generator like rand::thread_rng()
    q1t_state view;                  // borrowed handle on the live q_state for the inner ABI accessors
    circuit(size_t nq, size_t nc) : impl(nq, nc), thread_rng(nullptr) { view.impl = nullptr; view.owns = false; }
    ~circuit() { if (thread_rng) q1t_rng_free(thread_rng); }
};

static char *dup_cstring(const std::string &s)
{
    char *p = static_cast<char *>(std::malloc(s.size() + 1));
    std::memcpy(p, s.c_str(), s.size() + 1);
    return p;
}
static result_t res_empty() { result_t r = { nullptr, 0, 0, RESULT_EMPTY }; return r; }
static result_t res_error(const std::string &msg) { result_t r = { dup_cstring(msg), 0, 0, RESULT_ERROR }; return r; }
static result_t res_from(const CircuitError &e) { return e.code ? res_error(e.msg) : res_empty(); }

static bool parse_basis(char dir, Basis &b)
{
    switch (dir) {
    case 'x': case 'X': b = Basis::X; return true;
    case 'y': case 'Y': b = Basis::Y; return true;
    case 'z': case 'Z': b = Basis::Z; return true;
    default: return false;
    }
}

static bool make_gate(const char *gate, const parameter_t *params, size_t nparams, GateSpec &g, std::string &err)
{
    std::vector<Param> ps(params ? nparams : 0);
    for (size_t i = 0; i < ps.size(); ++i) {
        ps[i].value = params[i].value;
        ps[i].ptr = params[i].value_ptr;      // non-NULL pointer = by-reference parameter (ffi.rs:48-61)
    }
    return q1t::gate_spec_from_name(gate, ps.data(), ps.size(), g, err) == Q1T_OK;
}

extern "C" {

void result_free(result_t res)
{
    switch (res.restype) {
    case RESULT_ERROR:
    case RESULT_STRING:
        std::free(res.data);
        break;
    case RESULT_HISTOGRAM: {
        histelem_t *el = static_cast<histelem_t *>(res.data);
        for (size_t i = 0; i < res.length; ++i) std::free(const_cast<char *>(el[i].key));
        std::free(el);
        break;
    }
    case RESULT_CSTATE:
    case RESULT_HISTOGRAM_U64:
        std::free(res.data);
        break;
    default:
        break;
    }
}

circuit_t *circuit_new(size_t nr_qbits, size_t nr_cbits) { return new circuit(nr_qbits, nr_cbits); }
void circuit_free(circuit_t *ptr) { delete ptr; }
size_t circuit_nr_qbits(const circuit_t *ptr) { return ptr ? ptr->impl.nr_qbits() : 0; }
size_t circuit_nr_cbits(const circuit_t *ptr) { return ptr ? ptr->impl.nr_cbits() : 0; }

result_t circuit_cstate(const circuit_t *ptr)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    if (!ptr->impl.executed()) return res_error("Circuit has not been run yet");       // ffi.rs:212
    const std::vector<uint64_t> &cs = ptr->impl.cstate();
    uint64_t *d = static_cast<uint64_t *>(std::malloc(sizeof(uint64_t) * (cs.size() ? cs.size() : 1)));
    std::memcpy(d, cs.data(), sizeof(uint64_t) * cs.size());
    result_t r = { d, cs.size(), cs.size(), RESULT_CSTATE };
    return r;
}

size_t circuit_cstate_into(const circuit_t *ptr, uint64_t *out, size_t out_len)
{
    if (!ptr || !ptr->impl.executed() || !out) return 0;
    const std::vector<uint64_t> &cs = ptr->impl.cstate();
    const size_t n = cs.size() < out_len ? cs.size() : out_len;
    std::memcpy(out, cs.data(), sizeof(uint64_t) * n);
    return n;
}

result_t circuit_add_gate(circuit_t *ptr, const char *gate, const size_t *qbits, size_t nr_qbits,
                          const parameter_t *param_ptr, size_t nr_params)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    if (!qbits) return res_error("Pointer to bit indices is NULL");
    if (!gate) return res_error("Invalid gate name");
    GateSpec g;
    std::string err;
    if (!make_gate(gate, param_ptr, nr_params, g, err)) return res_error(err);
    return res_from(ptr->impl.add_gate(g, std::vector<size_t>(qbits, qbits + nr_qbits)));
}

result_t circuit_add_conditional_gate(circuit_t *ptr, const size_t *control_ptr, size_t nr_control, uint64_t target,
                                      const char *gate, const size_t *qbits_ptr, size_t nr_qbits,
                                      const parameter_t *param_ptr, size_t nr_params)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    if (!control_ptr) return res_error("Pointer to control bit indices is NULL");
    if (!qbits_ptr) return res_error("Pointer to bit indices is NULL");
    if (!gate) return res_error("Invalid gate name");
    GateSpec g;
    std::string err;
    if (!make_gate(gate, param_ptr, nr_params, g, err)) return res_error(err);
    return res_from(ptr->impl.add_conditional_gate(std::vector<size_t>(control_ptr, control_ptr + nr_control), target, g,
                                                   std::vector<size_t>(qbits_ptr, qbits_ptr + nr_qbits)));
}

static bool make_matrix_gate(const char *desc, const double *m, size_t dim, GateSpec &g, std::string &err)
{
    size_t k = 0;
    while (((size_t)1 << k) < dim) ++k;
    if (!m || dim < 2 || ((size_t)1 << k) != dim || k > 12) { err = "Invalid gate matrix dimension"; return false; }
    g = GateSpec();
    g.nr_bits = k;
    g.matrix.resize(dim * dim);
    std::memcpy(static_cast<void *>(g.matrix.data()), m, sizeof(double) * 2 * dim * dim);
    if (desc) g.user_desc = desc;
    return true;
}

result_t circuit_add_matrix_gate(circuit_t *ptr, const char *description, const double *matrix, size_t dim,
                                 const size_t *qbits, size_t nr_qbits)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    if (!qbits) return res_error("Pointer to bit indices is NULL");
    GateSpec g;
    std::string err;
    if (!make_matrix_gate(description, matrix, dim, g, err)) return res_error(err);
    return res_from(ptr->impl.add_gate(g, std::vector<size_t>(qbits, qbits + nr_qbits)));
}

result_t circuit_add_conditional_matrix_gate(circuit_t *ptr, const size_t *control_ptr, size_t nr_control, uint64_t target,
                                             const char *description, const double *matrix, size_t dim,
                                             const size_t *qbits, size_t nr_qbits)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    if (!control_ptr) return res_error("Pointer to control bit indices is NULL");
    if (!qbits) return res_error("Pointer to bit indices is NULL");
    GateSpec g;
    std::string err;
    if (!make_matrix_gate(description, matrix, dim, g, err)) return res_error(err);
    return res_from(ptr->impl.add_conditional_gate(std::vector<size_t>(control_ptr, control_ptr + nr_control), target, g,
                                                   std::vector<size_t>(qbits, qbits + nr_qbits)));
}

result_t circuit_add_composite_gate(circuit_t *ptr, const char *name, const char *description,
                                    const size_t *qbits, size_t nr_qbits, size_t nr_iterations)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    if (!qbits && nr_qbits) return res_error("Pointer to bit indices is NULL");
    if (!description) return res_error("Invalid gate description");
    return res_from(ptr->impl.add_composite(name ? name : "composite", description,
                                            std::vector<size_t>(qbits, qbits + nr_qbits), nr_iterations, nr_iterations != 1));
}

result_t circuit_add_loop_gate(circuit_t *ptr, const char *label, const char *body_description,
                               const size_t *qbits, size_t nr_qbits, size_t nr_iterations)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    if (!qbits && nr_qbits) return res_error("Pointer to bit indices is NULL");
    if (!body_description) return res_error("Invalid gate description");
    return res_from(ptr->impl.add_composite(label ? label : "loop", body_description,
                                            std::vector<size_t>(qbits, qbits + nr_qbits), nr_iterations, true));
}

size_t circuit_nr_ops(const circuit_t *ptr) { return ptr ? ptr->impl.nr_ops() : 0; }

result_t circuit_barrier(circuit_t *ptr, const size_t *qbits, size_t nr_qbits)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    return res_from(ptr->impl.barrier(std::vector<size_t>(qbits, qbits + (qbits ? nr_qbits : 0))));
}

result_t circuit_reset(circuit_t *ptr, size_t qbit)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    return res_from(ptr->impl.reset(qbit));
}

result_t circuit_reset_all(circuit_t *ptr)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    ptr->impl.reset_all();
    return res_empty();
}

result_t circuit_measure(circuit_t *ptr, size_t qbit, size_t cbit, char dir, uint8_t collapse)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    Basis b;
    if (!parse_basis(dir, b)) return res_error("Invalid measurement basis '" + std::to_string((int)dir) + "'");
    return res_from(collapse ? ptr->impl.measure_basis(qbit, cbit, b) : ptr->impl.peek_basis(qbit, cbit, b));
}

result_t circuit_measure_all(circuit_t *ptr, const size_t *cbits, size_t nr_cbits, char dir, uint8_t collapse)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    if (!cbits) return res_error("Pointer to measurement bit indices is NULL");
    Basis b;
    if (!parse_basis(dir, b)) return res_error("Invalid measurement basis '" + std::to_string((int)dir) + "'");
    std::vector<size_t> cb(cbits, cbits + nr_cbits);
    return res_from(collapse ? ptr->impl.measure_all_basis(cb, b) : ptr->impl.peek_all_basis(cb, b));
}

static q1t_rng thread_rng(circuit_t *ptr)
{
    if (!ptr->thread_rng) ptr->thread_rng = q1t_rng_entropy();
    return q1t_rng_handle(ptr->thread_rng);
}

result_t circuit_execute(circuit_t *ptr, size_t nr_shots)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    return res_from(ptr->impl.execute(nr_shots, thread_rng(ptr)));
}

result_t circuit_reexecute(circuit_t *ptr)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    return res_from(ptr->impl.reexecute(thread_rng(ptr)));
}

result_t circuit_execute_with_rng(circuit_t *ptr, size_t nr_shots, q1t_rng rng)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    if (!rng.next_u64) return res_error("Random generator callback is NULL");
    return res_from(ptr->impl.execute(nr_shots, rng));
}

result_t circuit_reexecute_with_rng(circuit_t *ptr, q1t_rng rng)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    if (!rng.next_u64) return res_error("Random generator callback is NULL");
    return res_from(ptr->impl.reexecute(rng));
}

result_t circuit_execute_with_qubit_coefs(circuit_t *ptr, size_t nr_shots, q1t_rng rng, const double *coefs)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    if (!rng.next_u64 || !coefs) return res_error("NULL argument");
    return res_from(ptr->impl.execute(nr_shots, rng, coefs));
}

result_t circuit_histogram(const circuit_t *ptr)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    if (!ptr->impl.executed()) return res_error("The circuit has not been executed yet");
    const std::map<std::string, size_t> h = ptr->impl.histogram_string();
    histelem_t *el = static_cast<histelem_t *>(std::malloc(sizeof(histelem_t) * (h.size() ? h.size() : 1)));
    size_t i = 0;
    for (const auto &kv : h) { el[i].key = dup_cstring(kv.first); el[i].count = kv.second; ++i; }
    result_t r = { el, h.size(), h.size(), RESULT_HISTOGRAM };
    return r;
}

result_t circuit_histogram_u64(const circuit_t *ptr)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    if (!ptr->impl.executed()) return res_error("The circuit has not been executed yet");
    const std::map<uint64_t, size_t> h = ptr->impl.histogram();
    histelem_u64_t *el = static_cast<histelem_u64_t *>(std::malloc(sizeof(histelem_u64_t) * (h.size() ? h.size() : 1)));
    size_t i = 0;
    for (const auto &kv : h) { el[i].key = kv.first; el[i].count = kv.second; ++i; }
    result_t r = { el, h.size(), h.size(), RESULT_HISTOGRAM_U64 };
    return r;
}

result_t circuit_set_cstate(circuit_t *ptr, const uint64_t *words, size_t n)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    if (!words && n) return res_error("NULL classical register");
    return res_from(ptr->impl.set_cstate(words, n));
}

static result_t res_string(const std::string &text) { result_t r = { dup_cstring(text), 0, 0, RESULT_STRING }; return r; }

result_t circuit_latex(const circuit_t *ptr)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    std::string text;
    const CircuitError e = ptr->impl.latex(text);
    return e.code ? res_error(e.msg) : res_string(text);
}
result_t circuit_open_qasm(const circuit_t *ptr)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    std::string text;
    const CircuitError e = ptr->impl.open_qasm(text);
    return e.code ? res_error(e.msg) : res_string(text);
}
result_t circuit_c_qasm(const circuit_t *ptr)
{
    if (!ptr) return res_error("Pointer to circuit is NULL");
    std::string text;
    const CircuitError e = ptr->impl.c_qasm(text);
    return e.code ? res_error(e.msg) : res_string(text);
}

int circuit_set_device(circuit_t *ptr, int device)
{
    if (!ptr) return Q1T_ERR_INVALID_ARGUMENT;
    ptr->impl.device = device;
    return Q1T_OK;
}

// extension: execute() on a state sharded over `n` devices of this process (a power of two >= 2; the list may repeat a
// device: several shards on one GPU); n = 0 goes back to one device.  The circuit may then hold gates, measure_all /
// peek_all and barriers, and reaches 34-36 qubits on 8 B200.
int circuit_set_devices(circuit_t *ptr, const int *devices, size_t n)
{
    if (!ptr || (n && !devices) || n == 1 || (n & (n - 1))) return Q1T_ERR_INVALID_ARGUMENT;
    ptr->impl.devices.assign(devices, devices + n);
    return Q1T_OK;
}
// amplitudes of the last sharded run in canonical index order; remaps / exchanged qubits / local relabels it cost
int circuit_sharded_amplitudes(circuit_t *ptr, size_t offset, size_t len, double *out)
{
    if (!ptr || !out || !ptr->impl.sharded_state()) return Q1T_ERR_INVALID_ARGUMENT;
    return ptr->impl.sharded_state()->read_amplitudes(offset, len, out);
}
int circuit_sharded_counters(circuit_t *ptr, uint64_t *out3)
{
    if (!ptr || !out3) return Q1T_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < 3; ++i) out3[i] = ptr->impl.sharded_counters[i];
    return Q1T_OK;
}

q1t_state *circuit_state(circuit_t *ptr)
{
    if (!ptr || !ptr->impl.state()) return nullptr;
    ptr->view.impl = ptr->impl.state();
    ptr->view.owns = false;
    return &ptr->view;
}

int circuit_engine_stats(circuit_t *ptr, q1t_stats *out)
{
    if (!ptr || !out || !ptr->impl.state()) return Q1T_ERR_INVALID_ARGUMENT;
    *out = ptr->impl.state()->stats;
    return Q1T_OK;
}

}  // extern "C"
