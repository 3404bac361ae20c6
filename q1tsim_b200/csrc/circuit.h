// circuit.h -- host-side mirror of the reference's `Circuit` (src/circuit.rs):
// builder methods, the op interpreter `do_execute_with` and histograms, on top
// of DeviceVectorState.  Same names, argument meaning and error behaviour.
#pragma once
#include <complex>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "engine.h"
#include "sharded.h"

namespace q1t {

enum class Basis { X, Y, Z };

// gates/parameter.rs:21-50: a value, or a pointer read at *execute* time
struct Param {
    double value = 0;
    const double *ptr = nullptr;
    double get() const { return ptr ? *ptr : value; }
};

struct GateSpec {
    std::string name;                  // lower-case table name, or empty for a matrix gate
    std::vector<Param> params;
    std::vector<std::complex<double>> matrix;   // user gate: matrix() given directly
    size_t nr_bits = 0;
    std::string user_desc;             // description() of a matrix gate, as given by the caller
    std::string description(bool with_values = true) const;
    // evaluate matrix() now (gates read their parameters at execution time)
    int evaluate(std::vector<std::complex<double>> &out) const;
};

struct CircuitOp {       // circuit.rs:27-51
    enum Kind { Gate, ConditionalGate, Reset, ResetAll, Measure, MeasureAll, Peek, PeekAll, Barrier } kind;
    GateSpec gate;
    std::vector<size_t> bits;          // qubits (Gate, ConditionalGate, Barrier) or cbits (MeasureAll, PeekAll)
    std::vector<size_t> control;
    uint64_t target = 0;
    size_t qbit = 0, cbit = 0;
    Basis basis = Basis::Z;
    // sub-gate of a flattened Composite / Loop (add_composite): the exporters print the group as the
    // single instruction it is in the reference's op list
    size_t group_id = 0, group_repeat = 1;
    bool group_loop = false;
    std::string group_name;
    std::vector<size_t> group_bits;    // the qubits the composite / loop was added on
};

struct CircuitError {
    int code = 0;
    std::string msg;
    explicit operator bool() const { return code != 0; }
};

class Circuit {
public:
    Circuit(size_t nr_qbits, size_t nr_cbits) : nr_qbits_(nr_qbits), nr_cbits_(nr_cbits) {}
    size_t nr_qbits() const { return nr_qbits_; }
    size_t nr_cbits() const { return nr_cbits_; }

    CircuitError add_gate(const GateSpec &g, const std::vector<size_t> &bits);
    CircuitError add_conditional_gate(const std::vector<size_t> &control, uint64_t target, const GateSpec &g,
                                      const std::vector<size_t> &bits);
    CircuitError measure_basis(size_t qbit, size_t cbit, Basis b);
    CircuitError measure_all_basis(const std::vector<size_t> &cbits, Basis b);
    CircuitError peek_basis(size_t qbit, size_t cbit, Basis b);
    CircuitError peek_all_basis(const std::vector<size_t> &cbits, Basis b);
    CircuitError reset(size_t qbit);
    void reset_all();
    CircuitError barrier(const std::vector<size_t> &qbits);
    // Composite::from_string (composite.rs:273-450) / Loop (staticloop.rs:71-92): the sub-gates of the
    // description are flattened into the op list on `bits`, the body `repeat` times
    CircuitError add_composite(const std::string &name, const std::string &desc, const std::vector<size_t> &bits, size_t repeat = 1,
                               bool is_loop = false);
    size_t nr_ops() const { return ops_.size(); }

    CircuitError execute(size_t nr_shots, q1t_rng rng, const double *qubit_coefs = nullptr);
    CircuitError reexecute(q1t_rng rng);

    bool executed() const { return has_cstate_; }
    const std::vector<uint64_t> &cstate() const { return c_state_; }
    ShardedVectorState *sharded_state() { return s_state_.get(); }
    CircuitError set_cstate(const uint64_t *w, size_t n);
    std::map<uint64_t, size_t> histogram() const;
    std::map<std::string, size_t> histogram_string() const;
    // export.cpp: circuit.rs:877-1146
    CircuitError open_qasm(std::string &out) const;
    CircuitError c_qasm(std::string &out) const;
    CircuitError latex(std::string &out) const;          // latex.cpp: circuit.rs:1148-1231
    DeviceVectorState *state() { return q_state_.get(); }
    int device = 0;
    // >= 2 entries (they may repeat): execute() runs on a state sharded over these devices of this process
    // (ShardedVectorState): gates, measure_all / peek_all in any basis, barriers
    std::vector<int> devices;
    uint64_t sharded_counters[3] = { 0, 0, 0 };      // remaps, exchanged qubits, local relabels of the last sharded run

private:
    size_t nr_qbits_, nr_cbits_;
    std::unique_ptr<DeviceVectorState> q_state_;
    bool has_cstate_ = false;
    std::vector<uint64_t> c_state_;
    std::vector<CircuitOp> ops_;
    std::unique_ptr<ShardedVectorState> s_state_;
    std::vector<int> s_devices_;
    CircuitError execute_sharded(size_t nr_shots, q1t_rng rng);
    // lowering of the leading run of constant-parameter gates, recorded at the first execute() (do_execute)
    std::vector<LoweredGate> lowered_;
    size_t lowered_len_ = 0;
    bool lowered_valid_ = false, fresh_state_ = false;
    uint64_t lowered_key_[2] = { 0, 0 };
    void publish_lowered();
    size_t next_group_ = 0;
    size_t export_group(size_t at, const std::vector<std::string> &names, bool cq, std::string &out, CircuitError &err) const;
    CircuitError state_err(int rc);
    CircuitError do_execute(q1t_rng rng);
};

int gate_spec_from_name(const char *name, const Param *params, size_t nparams, GateSpec &out, std::string &err);

// composite.cpp: text front-end (composite.rs:92-450, expression.rs:86-300)
struct SubGateDesc { std::string name; std::vector<double> args; std::vector<size_t> bits; };
bool parse_expression(const char *text, double &out, const char **rest, std::string &err);
bool parse_composite(const std::string &desc, std::vector<SubGateDesc> &out, size_t &nr_bits, std::string &err);
int composite_matrix(const std::string &desc, std::vector<std::complex<double>> &out, std::string &err);

}  // namespace q1t
