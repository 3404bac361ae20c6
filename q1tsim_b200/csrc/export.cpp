// export.cpp -- OpenQasm and c-Qasm export of a circuit: host-side text only.
//   Circuit::open_qasm      circuit.rs:877-1017 (+ is_full_register / check_open_qasm_condition_bits, :843-875)
//   Circuit::c_qasm         circuit.rs:1019-1146
//   per-gate instructions   `impl OpenQasm` / `impl CQasm` of src/gates/<gate>.rs, the
//                           declare_controlled_qasm! templates (controlled.rs:222-296, :402-552),
//                           Composite (composite.rs:525-609), Loop (staticloop.rs:111-188)
//   conditional forms       export/openqasm.rs:39-44, export/cqasm.rs:39-56
//   error texts             error.rs:35-63
// The texts are the reference's, quirks included (U2/U3 c-Qasm lines, `ry` exported as `u3`).
#include <charconv>
#include <cstring>
#include <string>
#include <vector>

#include "circuit.h"

namespace q1t {

namespace {

// Rust `{}` of an f64: shortest digits that round-trip, never an exponent
std::string fmt_f64(double v)
{
    if (v != v) return "NaN";
    if (v == 1.0 / 0.0) return "inf";
    if (v == -1.0 / 0.0) return "-inf";
    char buf[400];
    const std::to_chars_result r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

struct QasmEntry {
    const char *name;
    const char *args[3];      // parameter names of the templates
    const char *open_qasm;    // nullptr: declare_controlled_qasm! default form
    const char *c_qasm;
};

// `{i}` = name of gate bit i, `{arg}` = parameter, other `{...}` = arithmetic evaluated at export time
const QasmEntry kTable[] = {
    { "h", { 0, 0, 0 }, "h {0}", "h {0}" },                                                     // hadamard.rs:125-141
    { "i", { 0, 0, 0 }, "id {0}", "i {0}" },                                                    // identity.rs:73-89
    { "x", { 0, 0, 0 }, "x {0}", "x {0}" },                                                     // x.rs:108-124
    { "y", { 0, 0, 0 }, "y {0}", "y {0}" },                                                     // y.rs:100-116
    { "z", { 0, 0, 0 }, "z {0}", "z {0}" },                                                     // z.rs:88-104
    { "s", { 0, 0, 0 }, "s {0}", "s {0}" },                                                     // s.rs:106-122
    { "sdg", { 0, 0, 0 }, "sdg {0}", "sdag {0}" },                                              // s.rs:232-248
    { "t", { 0, 0, 0 }, "t {0}", "t {0}" },                                                     // t.rs:80-96
    { "tdg", { 0, 0, 0 }, "tdg {0}", "tdag {0}" },                                              // t.rs:181-197
    { "v", { 0, 0, 0 }, "u3(pi/2, -pi/2, pi/2) {0}", "x90 {0}" },                               // v.rs:85-101
    { "vdg", { 0, 0, 0 }, "u3(pi/2, pi/2, -pi/2) {0}", "mx90 {0}" },                            // v.rs:184-200
    { "rx", { "theta", 0, 0 }, "rx({theta}) {0}", "rx {0}, {theta}" },                          // rx.rs:115-131
    { "ry", { "theta", 0, 0 }, "u3({theta}, 0, 0) {0}", "ry {0}, {theta}" },                    // ry.rs:115-134
    { "rz", { "lambda", 0, 0 }, "rz({lambda}) {0}", "rz {0}, {lambda}" },                       // rz.rs:105-121
    { "u1", { "lambda", 0, 0 }, "u1({lambda}) {0}", "rz {0}, {lambda}" },                       // u1.rs:87-104
    { "u2", { "phi", "lambda", 0 }, "u2({phi}, {lambda}) {0}",
      "rz {0}, {{lambda} + pi}\nh {0}\nrz {0} {phi}" },                                         // u2.rs:82-100
    { "u3", { "theta", "phi", "lambda" }, "u3({theta}, {phi}, {lambda}) {0}",
      "rz {0}, {lambda}\nry {0}, {theta}\n; rz {0} {phi}" },                                    // u3.rs:87-106
    { "cx", { 0, 0, 0 }, "cx {0}, {1}", "cnot {0}, {1}" },                                      // cx.rs:87-104
    { "cy", { 0, 0, 0 }, "cy {0}, {1}", "sdag {1}\ncnot {0}, {1}\ns {1}" },                     // cy.rs:87-106
    { "cz", { 0, 0, 0 }, "cz {0}, {1}", "cz {0}, {1}" },                                        // cz.rs:87-104
    { "swap", { 0, 0, 0 }, "cx {0}, {1}; cx {1}, {0}; cx {0}, {1}", "swap {0}, {1}" },          // swap.rs:113-131
    // controlled.rs:402-552
    { "ch", { 0, 0, 0 }, nullptr, nullptr },
    { "crx", { "theta", 0, 0 },
      "s {1}; cx {0}, {1}; ry(-{theta}/2) {1}; cx {0}, {1}; ry({theta}/2) {1}; sdg {1}",
      "s {1}\ncnot {0}, {1}\nry {1}, {-0.5 * {theta}}\ncnot {0}, {1}\nry {1}, {0.5 * {theta}}\nsdag {1}" },
    { "cry", { "theta", 0, 0 },
      "cx {0}, {1}; u3(-{theta}/2, 0, 0) {1}; cx {0}, {1}; u3({theta}/2, 0, 0) {1}",
      "cnot {0}, {1}\nry {1}, -{0.5 * {theta}}\ncnot {0}, {1}\nry {1}, {0.5 * {theta}}" },
    { "crz", { "lambda", 0, 0 }, nullptr, nullptr },
    { "cs", { 0, 0, 0 }, "cu1(pi/2) {0}, {1}", "crk {0}, {1}, 1" },
    { "csdg", { 0, 0, 0 }, "cu1(-pi/2) {0}, {1}", "cr {0}, {1}, -1.570796326794897" },
    { "ct", { 0, 0, 0 }, "cu1(pi/4) {0}, {1}", "crk {0}, {1}, 2" },
    { "ctdg", { 0, 0, 0 }, "cu1(-pi/4) {0}, {1}", "cr {0}, {1}, -0.7853981633974483" },
    { "cu1", { "lambda", 0, 0 }, nullptr, "cr {0}, {1}, {lambda}" },
    { "cu2", { "phi", "lambda", 0 }, nullptr, nullptr },
    { "cu3", { "theta", "phi", "lambda" }, nullptr,
      "rz {1}, {0.5 * ({lambda}-{phi})}\ncnot {0}, {1}\nrz {1}, {-0.5 * ({phi}+{lambda})}\nry {1}, {-0.5 * {theta}}\n"
      "cnot {0}, {1}\nry {1}, {0.5 * {theta}}\nrz {1}, {phi}\nrz {0}, {0.5 * ({phi} + {lambda})}" },
    { "cv", { 0, 0, 0 }, nullptr, nullptr },
    { "cvdg", { 0, 0, 0 }, nullptr, nullptr },
    { "ccrx", { "theta", 0, 0 },
      "s {2}; cx {1}, {2}; ry(-{theta}/4) {2}; cx {1}, {2}; ry({theta}/4) {2}; cx {0}, {1}; cx {1}, {2}; "
      "ry({theta}/4) {2}; cx {1}, {2}; ry(-{theta}/4) {2}; cx {0}, {1}; cx {0}, {2}; ry(-{theta}/4) {2}; "
      "cx {0}, {2}; ry({theta}/4) {2}; sdg {2}",
      "s {2}\ncnot {1}, {2}\nry {2}, {-0.25 * {theta}}\ncnot {1}, {2}\nry {2}, {0.25 * {theta}}\ncnot {0}, {1}\n"
      "cnot {1}, {2}\nry {2}, {0.25 * {theta}}\ncnot {1}, {2}\nry {2}, {-0.25 * {theta}}\ncnot {0}, {1}\n"
      "cnot {0}, {2}\nry {2}, {-0.25 * {theta}}\ncnot {0}, {2}\nry {2}, {0.25 * {theta}}\nsdag {2}" },
    { "ccry", { "theta", 0, 0 },
      "cx {1}, {2}; u3(-{theta}/4, 0, 0) {2}; cx {1}, {2}; u3({theta}/4, 0, 0) {2}; cx {0}, {1}; cx {1}, {2}; "
      "u3({theta}/4, 0, 0) {2}; cx {1}, {2}; u3(-{theta}/4, 0, 0) {2}; cx {0}, {1}; cx {0}, {2}; "
      "u3(-{theta}/4, 0, 0) {2}; cx {0}, {2}; u3({theta}/4, 0, 0) {2}",
      "cnot {1}, {2}\nry {2}, {-0.25 * {theta}}\ncnot {1}, {2}\nry {2}, {0.25 * {theta}}\ncnot {0}, {1}\n"
      "cnot {1}, {2}\nry {2}, {0.25 * {theta}}\ncnot {1}, {2}\nry {2}, {-0.25 * {theta}}\ncnot {0}, {1}\n"
      "cnot {0}, {2}\nry {2}, {-0.25 * {theta}}\ncnot {0}, {2}\nry {2}, {0.25 * {theta}}" },
    { "ccrz", { "lambda", 0, 0 },
      "crz({lambda}/2) {1}, {2}; cx {0}, {1}; crz(-{lambda}/2) {1}, {2}; cx {0}, {1}; crz({lambda}/2) {0}, {2}",
      "cr {1}, {2}, {0.5 * {lambda}}\ncnot {0}, {1}\ncr {1}, {2}, {-0.5 * {lambda}}\ncnot {0}, {1}\n"
      "cr {0}, {2}, {0.5 * {lambda}}" },
    { "ccx", { 0, 0, 0 }, nullptr, "toffoli {0}, {1}, {2}" },
    { "ccz", { 0, 0, 0 }, "h {2}; ccx {0}, {1}, {2}; h {2}", "h {2}\ntoffoli {0}, {1}, {2}\nh {2}" },
    { nullptr, { 0, 0, 0 }, nullptr, nullptr }
};

void replace_all(std::string &s, const std::string &pat, const std::string &rep)
{
    for (size_t at = s.find(pat); at != std::string::npos; at = s.find(pat, at + rep.size())) s.replace(at, pat.size(), rep);
}

CircuitError export_error(const std::string &msg)
{
    CircuitError e;
    e.code = Q1T_ERR_EXPORT;
    e.msg = msg;
    return e;
}

// one gate -> instruction text; `cq` selects c-Qasm
CircuitError gate_qasm(const GateSpec &g, const std::vector<std::string> &names, const std::vector<size_t> &bits, bool cq,
                       std::string &out)
{
    const char *method = cq ? "c-Qasm" : "OpenQasm";
    const QasmEntry *e = nullptr;
    for (const QasmEntry *t = kTable; t->name; ++t)
        if (g.name == t->name) e = t;
    if (!e)                                              // user gates: default trait methods, openqasm.rs:23-30
        return export_error(std::string("Export to ") + method + " was not implemented for \"" + g.description(false) + "\"");
    if (bits.size() != g.nr_bits) {                      // check_nr_bits, gates.rs:176-186
        char buf[320];
        std::snprintf(buf, sizeof buf, "Expected %zu bits for \"%s\", got %zu", g.nr_bits, g.description(false).c_str(), bits.size());
        CircuitError err;
        err.code = Q1T_ERR_INVALID_NR_BITS;
        err.msg = buf;
        return err;
    }
    std::vector<std::string> args;
    for (const Param &p : g.params) args.push_back(fmt_f64(p.get()));
    const char *tmpl = cq ? e->c_qasm : e->open_qasm;
    if (!tmpl) {
        // declare_controlled_qasm! without template (controlled.rs:224-262): lower-case type name,
        // OpenQasm: name(args) bits ; c-Qasm: name bits, args
        std::string res = g.name, joined;
        for (size_t a = 0; a < args.size(); ++a) joined += (a ? ", " : "") + args[a];
        if (!cq && !args.empty()) res += "(" + joined + ")";
        for (size_t b = 0; b < bits.size(); ++b) res += (b ? ", " : " ") + names[bits[b]];
        if (cq && !args.empty()) res += ", " + joined;
        out = res;
        return CircuitError();
    }
    // controlled.rs:264-296 (the single-gate impls are the same substitution with no arithmetic left)
    std::string res = tmpl;
    for (size_t b = 0; b < bits.size(); ++b) replace_all(res, "{" + std::to_string(b) + "}", names[bits[b]]);
    for (size_t a = 0; a < args.size() && a < 3 && e->args[a]; ++a) replace_all(res, std::string("{") + e->args[a] + "}", args[a]);
    size_t off = 0;
    for (size_t i = res.find('{', off); i != std::string::npos; i = res.find('{', off)) {
        const size_t close = res.find('}', i);
        if (close != std::string::npos) {
            const std::string expr = res.substr(i + 1, close - i - 1);
            double val = 0;
            const char *rest = nullptr;
            std::string perr;
            if (parse_expression(expr.c_str(), val, &rest, perr) && rest && *rest == '\0') res.replace(i, close + 1 - i, fmt_f64(val));
        }
        off = i + 1;
    }
    out = res;
    return CircuitError();
}

// export/openqasm.rs:39-44, export/cqasm.rs:39-56
CircuitError conditional_gate_qasm(const GateSpec &g, const std::string &condition, const std::vector<std::string> &names,
                                   const std::vector<size_t> &bits, bool cq, std::string &out)
{
    std::string unc;
    CircuitError e = gate_qasm(g, names, bits, cq, unc);
    if (e) return e;
    if (!cq) { out = "if (" + condition + ") " + unc; return e; }
    const size_t sp = unc.find(' ');
    if (sp == std::string::npos) return export_error("Unable to find gate name or argument in \"" + unc + "\"");
    out = "c-" + unc.substr(0, sp) + " " + condition + ", " + unc.substr(sp + 1);
    return e;
}

bool is_iota(const std::vector<size_t> &v, size_t n)
{
    if (v.size() != n) return false;
    for (size_t i = 0; i < n; ++i) if (v[i] != i) return false;
    return true;
}

}  // namespace

// A flattened Composite / Loop (Circuit::add_composite) is one instruction of the reference's op list:
// its sub-gates are joined the way Composite::open_qasm / Loop::open_qasm join them.  Returns the index
// one past the group that starts at `at`.
size_t Circuit::export_group(size_t at, const std::vector<std::string> &names, bool cq, std::string &out, CircuitError &err) const
{
    const CircuitOp &first = ops_[at];
    const size_t id = first.group_id, repeat = first.group_repeat;
    size_t end = at;
    while (end < ops_.size() && ops_[end].group_id == id) ++end;
    const size_t body = repeat ? (end - at) / repeat : 0;
    std::string text;
    for (size_t k = 0; k < body; ++k) {
        std::string one;
        err = gate_qasm(ops_[at + k].gate, names, ops_[at + k].bits, cq, one);
        if (err) return end;
        if (k) text += cq ? "\n" : "; ";
        text += one;
    }
    if (!first.group_loop) { out = text; return end; }                   // composite.rs:527-543, :569-586
    if (cq) {                                                            // staticloop.rs:157-163
        out = "." + first.group_name + "(" + std::to_string(repeat) + ")\n" + text + "\n.end";
        return end;
    }
    out.clear();                                                         // staticloop.rs:113-131
    for (size_t it = 0; it < repeat; ++it) out += (it ? ";\n" : "") + text;
    return end;
}

CircuitError Circuit::open_qasm(std::string &res) const
{
    res = "OPENQASM 2.0;\ninclude \"qelib1.inc\";\n";
    std::vector<std::string> qn, cn;
    if (nr_qbits_ > 0) res += "qreg q[" + std::to_string(nr_qbits_) + "];\n";
    for (size_t i = 0; i < nr_qbits_; ++i) qn.push_back("q[" + std::to_string(i) + "]");
    if (nr_cbits_ > 0) res += "creg b[" + std::to_string(nr_cbits_) + "];\n";
    for (size_t i = 0; i < nr_cbits_; ++i) cn.push_back("b[" + std::to_string(i) + "]");
    const std::vector<std::string> whole(1, "q");
    GateSpec h, sdg;
    std::string perr, line;
    gate_spec_from_name("h", nullptr, 0, h, perr);
    gate_spec_from_name("sdg", nullptr, 0, sdg, perr);
    const std::vector<size_t> bit0(1, 0);
    CircuitError e;
    for (size_t i = 0; i < ops_.size();) {
        const CircuitOp &op = ops_[i];
        if (op.kind == CircuitOp::Gate && op.group_id) {
            i = export_group(i, qn, false, line, e);
            if (e) return e;
            res += line + ";\n";
            continue;
        }
        ++i;
        switch (op.kind) {
        case CircuitOp::Gate:
            if ((e = gate_qasm(op.gate, qn, op.bits, false, line))) return e;
            res += line + ";\n";
            break;
        case CircuitOp::ConditionalGate: {
            if (op.control.empty()) {                                    // circuit.rs:914-917
                if ((e = gate_qasm(op.gate, qn, op.bits, false, line))) return e;
                res += line + ";\n";
                break;
            }
            // the control bits must span the whole classical register, in any order (circuit.rs:843-875)
            std::vector<bool> seen(nr_cbits_, false);
            bool full = op.control.size() == nr_cbits_;
            for (size_t c : op.control) {
                if (c >= nr_cbits_ || seen[c]) { full = false; break; }
                seen[c] = true;
            }
            if (!full) return export_error("OpenQasm can only perform conditional operations based on a complete classical register");
            uint64_t starget = 0;
            for (size_t t = 0; t < op.control.size(); ++t) starget |= ((op.target >> t) & 1ull) << op.control[t];
            if ((e = conditional_gate_qasm(op.gate, "b == " + std::to_string(starget), qn, op.bits, false, line))) return e;
            res += line + ";\n";
            break;
        }
        case CircuitOp::Measure:
            if (op.basis == Basis::Y) { gate_qasm(sdg, qn, std::vector<size_t>(1, op.qbit), false, line); res += line + ";\n"; }
            if (op.basis != Basis::Z) { gate_qasm(h, qn, std::vector<size_t>(1, op.qbit), false, line); res += line + ";\n"; }
            res += "measure " + qn[op.qbit] + " -> " + cn[op.cbit] + ";\n";
            break;
        case CircuitOp::MeasureAll:
            if (op.basis == Basis::Y) { gate_qasm(sdg, whole, bit0, false, line); res += line + ";\n"; }
            if (op.basis != Basis::Z) { gate_qasm(h, whole, bit0, false, line); res += line + ";\n"; }
            if (is_iota(op.bits, nr_cbits_)) res += "measure q -> b;\n";
            else
                for (size_t q = 0; q < op.bits.size(); ++q) {
                    if (q >= qn.size()) return export_error("measure_all lists more classical bits than there are qubits");
                    res += "measure " + qn[q] + " -> " + cn[op.bits[q]] + ";\n";
                }
            break;
        case CircuitOp::Peek:
        case CircuitOp::PeekAll:
            return export_error("Peeking into the quantum state is not a physical operation, and is not supported in OpenQasm");
        case CircuitOp::Reset:
            res += "reset " + qn[op.qbit] + ";\n";
            break;
        case CircuitOp::ResetAll:
            res += "reset q;\n";
            break;
        case CircuitOp::Barrier:
            if (is_iota(op.bits, nr_qbits_)) res += "barrier q;\n";
            else {
                res += "barrier ";
                for (size_t b = 0; b < op.bits.size(); ++b) res += (b ? ", " : "") + qn[op.bits[b]];
                res += ";\n";
            }
            break;
        }
    }
    return CircuitError();
}

CircuitError Circuit::c_qasm(std::string &res) const
{
    static const char *const no_creg = "In cQasm, no classical registers can be specified. Measurements must be made to a "
                                       "classical bit with the same index as the qubit";
    res = "version 1.0\n";
    std::vector<std::string> qn, cn;
    if (nr_qbits_ > 0) res += "qubits " + std::to_string(nr_qbits_) + "\n";
    for (size_t i = 0; i < nr_qbits_; ++i) { qn.push_back("q[" + std::to_string(i) + "]"); cn.push_back("b[" + std::to_string(i) + "]"); }
    GateSpec h, sdg;
    std::string perr, line;
    gate_spec_from_name("h", nullptr, 0, h, perr);
    gate_spec_from_name("sdg", nullptr, 0, sdg, perr);
    CircuitError e;
    for (size_t i = 0; i < ops_.size();) {
        const CircuitOp &op = ops_[i];
        if (op.kind == CircuitOp::Gate && op.group_id) {
            i = export_group(i, qn, true, line, e);
            if (e) return e;
            res += line + "\n";
            continue;
        }
        ++i;
        switch (op.kind) {
        case CircuitOp::Gate:
            if ((e = gate_qasm(op.gate, qn, op.bits, true, line))) return e;
            res += line + "\n";
            break;
        case CircuitOp::ConditionalGate: {
            if (op.control.empty()) {
                if ((e = gate_qasm(op.gate, qn, op.bits, true, line))) return e;
                res += line + "\n";
                break;
            }
            // c-Qasm conditions hold when all listed bits are 1: flip the bits whose target value is 0
            // around the gate (circuit.rs:1054-1078); classical bit i is named after qubit i
            std::string flips, condition;
            for (size_t t = 0; t < op.control.size(); ++t) {
                const size_t idx = op.control[t];
                if (idx >= cn.size()) return export_error(no_creg);
                if (t >= 64 || ((op.target >> t) & 1ull) == 0) flips += "not " + cn[idx] + "\n";
                condition += (t ? ", " : "") + cn[idx];
            }
            if ((e = conditional_gate_qasm(op.gate, condition, qn, op.bits, true, line))) return e;
            res += flips + line + "\n" + flips;
            break;
        }
        case CircuitOp::Measure:
            if (op.qbit != op.cbit) return export_error(no_creg);
            res += std::string(op.basis == Basis::X ? "measure_x" : op.basis == Basis::Y ? "measure_y" : "measure") + " q["
                   + std::to_string(op.qbit) + "]\n";
            break;
        case CircuitOp::MeasureAll:
            for (size_t q = 0; q < op.bits.size(); ++q)
                if (op.bits[q] != q) return export_error(no_creg);
            for (size_t q = 0; q < nr_qbits_ && op.basis != Basis::Z; ++q) {
                if (op.basis == Basis::Y) { gate_qasm(sdg, qn, std::vector<size_t>(1, q), true, line); res += line + "\n"; }
                gate_qasm(h, qn, std::vector<size_t>(1, q), true, line);
                res += line + "\n";
            }
            res += "measure_all\n";
            break;
        case CircuitOp::Peek:
        case CircuitOp::PeekAll:
            return export_error("Peeking into the quantum state is not a physical operation, and is not supported in c-Qasm");
        case CircuitOp::Reset:
            res += "prep_z " + qn[op.qbit] + "\n";
            break;
        case CircuitOp::ResetAll:
            for (size_t q = 0; q < nr_qbits_; ++q) res += "prep_z " + qn[q] + "\n";
            break;
        case CircuitOp::Barrier:
            break;                                                       // not available in c-Qasm (circuit.rs:1138-1140)
        }
    }
    return CircuitError();
}

}  // namespace q1t
