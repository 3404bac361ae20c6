// latex.cpp -- export of a circuit to LaTeX / Qcircuit (completes the exporters of the ffi.rs surface,
// ffi.rs:624-639): host-side text only.
//   LatexState             export/latex.rs:19-543 (LatexExportState)
//   gate_latex             `impl Latex` of src/gates/<gate>.rs, C<G> controlled.rs:83-121, Swap swap.rs:134-150,
//                          default trait method export/latex.rs:547-554
//   Circuit::latex         circuit.rs:1148-1231; Loop staticloop.rs:190-221 (composites are drawn expanded,
//                          the reference's default, composite.rs:611-633 -- which is what the flattened op list is)
// Same output text and the same error variants as the reference.
#include <algorithm>
#include <cstdio>
#include <string>
#include <tuple>
#include <vector>

#include "circuit.h"

namespace q1t {

namespace {

CircuitError make_error(int code, const std::string &msg)
{
    CircuitError e;
    e.code = code;
    e.msg = msg;
    return e;
}
CircuitError bad_qbit(size_t b) { return make_error(Q1T_ERR_INVALID_QBIT, "Invalid index " + std::to_string(b) + " for a quantum bit"); }
CircuitError bad_cbit(size_t b) { return make_error(Q1T_ERR_INVALID_CBIT, "Invalid index " + std::to_string(b) + " for a classical bit"); }
std::string itoa(long v) { return std::to_string(v); }

// support.rs:20-44
std::vector<std::pair<size_t, size_t>> get_ranges(std::vector<size_t> nrs)
{
    std::vector<std::pair<size_t, size_t>> ranges;
    if (nrs.empty()) return ranges;
    std::sort(nrs.begin(), nrs.end());
    size_t first = nrs[0], last = nrs[0];
    for (size_t i = 1; i < nrs.size(); ++i) {
        if (nrs[i] == last + 1) { ++last; continue; }
        ranges.push_back(std::make_pair(first, last));
        first = last = nrs[i];
    }
    ranges.push_back(std::make_pair(first, last));
    return ranges;
}

class LatexState {
public:
    LatexState(size_t nq, size_t nc) : nq_(nq), nc_(nc), in_use_(nq + nc, true) {}

    // latex.rs:156-206
    CircuitError start_range_op(const std::vector<size_t> &qbits, const std::vector<size_t> *cbits)
    {
        std::vector<size_t> bits;
        CircuitError e = bit_indices(qbits, cbits, bits);
        if (e) return e;
        if (bits.empty()) return e;
        const size_t first = *std::min_element(bits.begin(), bits.end()), last = *std::max_element(bits.begin(), bits.end());
        if (reserved_.empty()) {
            bool used = false;
            for (size_t b = first; b <= last; ++b) used = used || in_use_[b];
            if (used) add_column();
            reserved_.push_back(std::make_pair(first, last));
        } else {
            if (reserved_.back().first <= first && reserved_.back().second >= last) reserved_.push_back(std::make_pair(first, last));
            else return make_error(Q1T_ERR_EXPORT, "Trying to reserve range of bits, but a previous reservation is still open");
        }
        return e;
    }
    void end_range_op()
    {
        if (reserved_.empty()) return;
        for (size_t b = reserved_.back().first; b <= reserved_.back().second; ++b) in_use_[b] = true;
        reserved_.pop_back();
    }
    // latex.rs:209-220
    CircuitError set_field(size_t bit, const std::string &contents)
    {
        if (reserved_.empty()) {
            if (bit >= nq_) return bad_qbit(bit);                     // reserve(&[bit], None), latex.rs:127-138
            if (in_use_[bit]) add_column();
        }
        if (matrix_.empty()) add_column();
        matrix_.back()[bit] = contents;
        present_.back()[bit] = 1;
        in_use_[bit] = true;
        return CircuitError();
    }
    // latex.rs:227-247
    CircuitError set_measurement(size_t qbit, size_t cbit, const char *basis)
    {
        const size_t cidx = nq_ + cbit;
        const std::vector<size_t> q(1, qbit), c(1, cbit);
        CircuitError e = start_range_op(q, &c);
        if (e) return e;
        if ((e = set_field(qbit, basis ? std::string("\\meterB{") + basis + "}" : std::string("\\meter")))) return e;
        if ((e = set_field(cidx, "\\cw \\cwx[" + itoa((long)qbit - (long)cidx) + "]"))) return e;
        end_range_op();
        return e;
    }
    CircuitError set_reset(size_t qbit) { return set_field(qbit, "\\push{~\\ket{0}~} \\ar @{|-{}} [0,-1]"); }       // latex.rs:250-253
    // latex.rs:263-292
    CircuitError set_condition(const std::vector<size_t> &control, uint64_t target, const std::vector<size_t> &qbits)
    {
        for (size_t b : qbits) if (b >= nq_) return bad_qbit(b);
        for (size_t b : control) if (b >= nc_) return bad_cbit(b);
        if (qbits.empty()) return CircuitError();
        size_t pbit = *std::max_element(qbits.begin(), qbits.end());
        std::vector<std::pair<size_t, size_t>> bp;
        for (size_t pos = 0; pos < control.size(); ++pos) bp.push_back(std::make_pair(nq_ + control[pos], pos));
        std::sort(bp.begin(), bp.end());
        for (const auto &x : bp) {
            const bool set = x.second < 64 && ((target >> x.second) & 1ull);
            CircuitError e = set_field(x.first, std::string(set ? "\\cctrl" : "\\cctrlo") + "{" + itoa((long)pbit - (long)x.first) + "}");
            if (e) return e;
            pbit = x.first;
        }
        return CircuitError();
    }
    // latex.rs:301-346
    CircuitError add_block_gate(const std::vector<size_t> &qbits, const std::string &desc)
    {
        const std::vector<std::pair<size_t, size_t>> ranges = get_ranges(qbits);
        if (ranges.empty()) return CircuitError();
        CircuitError e = start_range_op(qbits, nullptr);
        if (e) return e;
        size_t prev_last = 0;
        for (size_t r = 0; r < ranges.size(); ++r) {
            const size_t first = ranges[r].first, last = ranges[r].second;
            const std::string link = r ? " \\qwx[" + itoa((long)prev_last - (long)first) + "]" : "";
            if (first == last) e = set_field(first, "\\gate{" + desc + "}" + link);
            else {
                e = set_field(first, "\\multigate{" + itoa((long)(last - first)) + "}{" + desc + "}" + link);
                for (size_t b = first + 1; b <= last && !e; ++b) e = set_field(b, "\\ghost{" + desc + "}");
            }
            if (e) return e;
            prev_last = last;
        }
        end_range_op();
        return e;
    }
    // latex.rs:356-395
    void start_loop(size_t count)
    {
        reserve_all();
        open_loops_.push_back(std::make_pair(matrix_.size() - 1, count));
    }
    CircuitError end_loop()
    {
        if (open_loops_.empty()) return make_error(Q1T_ERR_EXPORT, "Unable to close loop, because no loop is currently open");
        loops_.push_back(std::make_tuple(open_loops_.back().first, matrix_.size() - 1, open_loops_.back().second));
        open_loops_.pop_back();
        reserve_all();
        return CircuitError();
    }
    CircuitError add_cds(size_t bit, size_t count, const char *label)
    {
        reserve_all();
        CircuitError e = set_field(bit, "\\cds{" + itoa((long)count) + "}{" + label + "}");
        reserve_all();
        return e;
    }
    // latex.rs:398-412
    CircuitError set_barrier(const std::vector<size_t> &qbits)
    {
        for (size_t b : qbits) if (b >= nq_) return bad_qbit(b);
        const std::vector<std::pair<size_t, size_t>> ranges = get_ranges(qbits);
        add_column();
        for (const auto &r : ranges) {
            CircuitError e = set_field(r.first, "\\qw \\barrier{" + itoa((long)(r.second - r.first)) + "}");
            if (e) return e;
        }
        return CircuitError();
    }
    // latex.rs:420-490
    std::string code() const
    {
        std::string res = "\\Qcircuit @C=1em @R=.7em {\n";
        if (!loops_.empty()) {
            size_t prev = 0;
            res += "    & ";
            for (const auto &l : loops_) {
                const size_t start = std::get<0>(l), end = std::get<1>(l), count = std::get<2>(l);
                for (size_t i = prev; i < start; ++i) res += "& ";
                const std::string s = std::to_string(start + 2), t = std::to_string(end + 2);
                res += "\\mbox{} \\POS\"2," + s + "\".\"2," + s + "\".\"2," + t + "\".\"2," + t + "\"!C*+<.7em>\\frm{^\\}},+U*++!D{" +
                       std::to_string(count) + "\\times}";
                prev = start;
            }
            res += "\\\\\n    ";
            for (size_t i = 0; i < matrix_.size(); ++i) res += "& ";
            res += "\\\\\n";
        }
        bool last_used = false;
        for (bool u : in_use_) last_used = last_used || u;
        for (size_t i = 0; i < nq_ + nc_; ++i) {
            res += i < nq_ ? "    \\lstick{\\ket{0}}" : "    \\lstick{0}";
            for (size_t c = 0; c < matrix_.size(); ++c) {
                res += " & ";
                if (present_[c][i]) res += matrix_[c][i];
                else res += i < nq_ ? "\\qw" : "\\cw";
            }
            if (last_used) res += i < nq_ ? " & \\qw" : " & \\cw";
            res += " \\\\\n";
        }
        res += "}\n";
        return res;
    }
    bool set_controlled(bool c) { const bool old = controlled_; controlled_ = c; return old; }
    bool is_controlled() const { return controlled_; }

private:
    size_t nq_, nc_;
    std::vector<std::vector<std::string>> matrix_;       // one entry per drawn column
    std::vector<std::vector<char>> present_;
    std::vector<bool> in_use_;
    bool controlled_ = false;
    std::vector<std::pair<size_t, size_t>> reserved_;
    std::vector<std::tuple<size_t, size_t, size_t>> loops_;
    std::vector<std::pair<size_t, size_t>> open_loops_;

    void add_column()
    {
        matrix_.push_back(std::vector<std::string>(nq_ + nc_));
        present_.push_back(std::vector<char>(nq_ + nc_, 0));
        in_use_.assign(nq_ + nc_, false);
    }
    void reserve_all()
    {
        bool used = false;
        for (bool u : in_use_) used = used || u;
        if (used) add_column();
    }
    CircuitError bit_indices(const std::vector<size_t> &qbits, const std::vector<size_t> *cbits, std::vector<size_t> &out) const
    {
        for (size_t b : qbits) if (b >= nq_) return bad_qbit(b);
        out = qbits;
        if (cbits)
            for (size_t b : *cbits) {
                if (b >= nc_) return bad_cbit(b);
                out.push_back(nq_ + b);
            }
        return CircuitError();
    }
};

std::string fmt4(double v)
{
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.4f", v);
    return buf;
}

CircuitError nr_bits_error(const GateSpec &g, size_t have)
{
    char buf[320];
    std::snprintf(buf, sizeof buf, "Expected %zu bits for \"%s\", got %zu", g.nr_bits, g.description(false).c_str(), have);
    return make_error(Q1T_ERR_INVALID_NR_BITS, buf);
}

// `name` = table name with `nctl` leading controls already stripped
CircuitError named_gate_latex(const std::string &name, const std::vector<double> &p, const std::vector<size_t> &bits, LatexState &st)
{
    if (name.size() > 1 && name[0] == 'c') {
        // C<G>, controlled.rs:83-121: control = bits[0], the rest is drawn as controlled
        const size_t control = bits[0];
        const std::vector<size_t> rest(bits.begin() + 1, bits.end());
        const size_t mn = *std::min_element(rest.begin(), rest.end()), mx = *std::max_element(rest.begin(), rest.end());
        CircuitError e = st.start_range_op(bits, nullptr);
        if (e) return e;
        if (mn > control && mx > control) e = st.set_field(control, "\\ctrl{" + itoa((long)(mn - control)) + "}");
        else if (mn < control && mx < control) e = st.set_field(control, "\\ctrl{" + itoa((long)mx - (long)control) + "}");
        else return make_error(Q1T_ERR_EXPORT, "Unable to draw controlled gate with control in the middle");     // the reference panics here
        if (e) return e;
        const bool was = st.set_controlled(true);
        e = named_gate_latex(name.substr(1), p, rest, st);
        st.set_controlled(was);
        if (e) return e;
        st.end_range_op();
        return e;
    }
    if (name == "i") return st.set_field(bits[0], "\\qw");                                                       // identity.rs
    if (name == "x") return st.set_field(bits[0], st.is_controlled() ? "\\targ" : "\\gate{X}");                  // x.rs
    if (name == "z") return st.set_field(bits[0], st.is_controlled() ? "\\control \\qw" : "\\gate{Z}");          // z.rs
    if (name == "swap") {                                                                                         // swap.rs:134-150
        const size_t b0 = std::min(bits[0], bits[1]), b1 = std::max(bits[0], bits[1]);
        CircuitError e = st.start_range_op(bits, nullptr);
        if (e) return e;
        if ((e = st.set_field(b0, "\\qswap \\qwx[" + itoa((long)(b1 - b0)) + "]"))) return e;
        if ((e = st.set_field(b1, "\\qswap"))) return e;
        st.end_range_op();
        return e;
    }
    std::string d;
    if (name == "h") d = "H";
    else if (name == "y") d = "Y";
    else if (name == "s") d = "S";
    else if (name == "sdg") d = "S^\\dagger";
    else if (name == "t") d = "T";
    else if (name == "tdg") d = "T^\\dagger";
    else if (name == "v") d = "V";
    else if (name == "vdg") d = "V^\\dagger";
    else if (name == "rx" && p.size() == 1) d = "R_x(" + fmt4(p[0]) + ")";
    else if (name == "ry" && p.size() == 1) d = "R_y(" + fmt4(p[0]) + ")";
    else if (name == "rz" && p.size() == 1) d = "R_z(" + fmt4(p[0]) + ")";
    else if (name == "u1" && p.size() == 1) d = "U_1(" + fmt4(p[0]) + ")";
    else if (name == "u2" && p.size() == 2) d = "U_2(" + fmt4(p[0]) + ", " + fmt4(p[1]) + ")";
    else if (name == "u3" && p.size() == 3) d = "U_3(" + fmt4(p[0]) + ", " + fmt4(p[1]) + ", " + fmt4(p[2]) + ")";
    else return make_error(Q1T_ERR_EXPORT, "Export to LaTeX was not implemented for \"" + name + "\"");
    return st.add_block_gate(bits, d);
}

CircuitError gate_latex(const GateSpec &g, const std::vector<size_t> &bits, LatexState &st)
{
    if (bits.size() != g.nr_bits) return nr_bits_error(g, bits.size());           // check_nr_bits, gates.rs:176-186
    if (g.name.empty()) return st.add_block_gate(bits, g.description(false));     // user gate: default trait method, latex.rs:547-554
    std::vector<double> p;
    for (const Param &x : g.params) p.push_back(x.get());
    return named_gate_latex(g.name, p, bits, st);
}

}  // namespace

// circuit.rs:1148-1231
CircuitError Circuit::latex(std::string &out) const
{
    LatexState st(nr_qbits_, nr_cbits_);
    CircuitError e;
    for (size_t i = 0; i < ops_.size(); ++i) {
        const CircuitOp &op = ops_[i];
        switch (op.kind) {
        case CircuitOp::Gate:
            if (op.group_id && op.group_loop) {
                // Loop, staticloop.rs:190-221: two iterations or fewer are drawn out; more as body, dots, body under a brace
                size_t end = i;
                while (end < ops_.size() && ops_[end].group_id == op.group_id) ++end;
                const size_t repeat = op.group_repeat, body = repeat ? (end - i) / repeat : 0;
                auto draw_body = [&]() -> CircuitError {
                    for (size_t k = 0; k < body; ++k) {
                        CircuitError be = gate_latex(ops_[i + k].gate, ops_[i + k].bits, st);
                        if (be) return be;
                    }
                    return CircuitError();
                };
                if (repeat <= 2) {
                    for (size_t it = 0; it < repeat && !e; ++it) e = draw_body();
                } else {
                    const std::vector<size_t> &lb = op.group_bits;
                    if (lb.empty()) return make_error(Q1T_ERR_EXPORT, "loop without bits");
                    const size_t mn = *std::min_element(lb.begin(), lb.end()), mx = *std::max_element(lb.begin(), lb.end());
                    st.start_loop(repeat);
                    if ((e = draw_body())) return e;
                    if ((e = st.add_cds(mn, mx - mn, "\\cdots"))) return e;
                    if ((e = draw_body())) return e;
                    e = st.end_loop();
                }
                if (e) return e;
                i = end - 1;
                break;
            }
            if ((e = gate_latex(op.gate, op.bits, st))) return e;
            break;
        case CircuitOp::ConditionalGate: {
            if ((e = st.start_range_op(op.bits, &op.control))) return e;
            const bool was = st.set_controlled(true);
            e = gate_latex(op.gate, op.bits, st);
            st.set_controlled(was);
            if (e) return e;
            if ((e = st.set_condition(op.control, op.target, op.bits))) return e;
            st.end_range_op();
            break;
        }
        case CircuitOp::Measure:
            if ((e = st.set_measurement(op.qbit, op.cbit, op.basis == Basis::X ? "X" : op.basis == Basis::Y ? "Y" : nullptr))) return e;
            break;
        case CircuitOp::MeasureAll:
            for (size_t q = 0; q < op.bits.size(); ++q)
                if ((e = st.set_measurement(q, op.bits[q], op.basis == Basis::X ? "X" : op.basis == Basis::Y ? "Y" : nullptr))) return e;
            break;
        case CircuitOp::Peek:
            return make_error(Q1T_ERR_EXPORT, "Export to LaTeX was not implemented for \"peek\"");
        case CircuitOp::PeekAll:
            return make_error(Q1T_ERR_EXPORT, "Export to LaTeX was not implemented for \"peek all\"");
        case CircuitOp::Reset:
            if ((e = st.set_reset(op.qbit))) return e;
            break;
        case CircuitOp::ResetAll: {
            if (nr_qbits_ == 0) break;
            std::vector<size_t> span;
            span.push_back(0);
            span.push_back(nr_qbits_ - 1);
            if ((e = st.start_range_op(span, nullptr))) return e;
            for (size_t q = 0; q < nr_qbits_; ++q)
                if ((e = st.set_reset(q))) return e;
            st.end_range_op();
            break;
        }
        case CircuitOp::Barrier:
            if ((e = st.set_barrier(op.bits))) return e;
            break;
        }
    }
    out = st.code();
    return CircuitError();
}

}  // namespace q1t
