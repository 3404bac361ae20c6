// composite.cpp -- text front-end for composite gates (host only).
//
//   parse_expression   expression.rs:86-300  (Expression::parse + eval): real / integer literals,
//                      `pi`, + - * / ^, unary minus, parentheses, sin cos tan exp ln sqrt
//   parse_composite    composite.rs:92-450   (Composite::from_string): "NAME(args) bits; ..." with
//                      the reference's gate table, arity checks and ParseError texts (error.rs:93-127)
//   Circuit::add_composite   SURVEY 8(f)2: the sub-gates are FLATTENED into the circuit's op list
//                      (composite bit i -> bits[i]), optionally repeated (Loop, staticloop.rs:78-92),
//                      instead of going through the composite's 2^k x 2^k matrix() as
//                      composite.rs:480-485 does -- so a C^9X ladder costs its sub-gates, not a
//                      1024 x 1024 dense block, and its phases fuse with their neighbours.
//   composite_matrix   the matrix() of a description (for small k; test hook and user-gate export)
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "circuit.h"

namespace q1t {

namespace {

void skip_ws(const char *&p)
{
    while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') ++p;
}

bool parse_sum(const char *&p, double &out, std::string &err);

// expression.rs:91-133 -- real literal, integer literal, or `pi`
bool parse_literal(const char *&p, double &out, std::string &err)
{
    const char *q = p;
    skip_ws(q);
    const char *s = q;
    // real: [0-9]+\.[0-9]* | \.[0-9]+, optional exponent
    const char *r = s;
    size_t nint = 0, nfrac = 0;
    while (std::isdigit((unsigned char)*r)) { ++r; ++nint; }
    if (*r == '.') {
        ++r;
        while (std::isdigit((unsigned char)*r)) { ++r; ++nfrac; }
        if (nint > 0 || nfrac > 0) {
            if (*r == 'e' || *r == 'E') {
                const char *e = r + 1;
                if (*e == '+' || *e == '-') ++e;
                if (std::isdigit((unsigned char)*e)) {
                    while (std::isdigit((unsigned char)*e)) ++e;
                    r = e;
                }
            }
            out = std::strtod(std::string(s, r).c_str(), nullptr);
            p = r;
            return true;
        }
    }
    // integer: [1-9][0-9]* | 0
    if (std::isdigit((unsigned char)*s)) {
        const char *e = s;
        if (*e == '0') ++e;
        else while (std::isdigit((unsigned char)*e)) ++e;
        errno = 0;
        const unsigned long long v = std::strtoull(std::string(s, e).c_str(), nullptr, 10);
        if (errno == ERANGE) {                        // parse::<u64>() fails, expression.rs:116-123
            err = std::string("Failed to parse argument \"") + p + "\"";
            return false;
        }
        out = (double)v;
        p = e;
        return true;
    }
    if (s[0] == 'p' && s[1] == 'i') {
        out = 3.14159265358979323846264338327950288;      // std::f64::consts::PI
        p = s + 2;
        return true;
    }
    err = std::string("Failed to parse argument \"") + p + "\"";
    return false;
}

// expression.rs:141-161
bool parse_parenthesized(const char *&p, double &out, std::string &err)
{
    const char *q = p;
    skip_ws(q);
    if (*q == '(') {
        const char *start = p;
        ++q;
        if (!parse_sum(q, out, err)) return false;
        skip_ws(q);
        if (*q != ')') { err = std::string("Unclosed parentheses in expression: \"") + start + "\""; return false; }
        p = q + 1;
        return true;
    }
    return parse_literal(p, out, err);
}

// expression.rs:171-192
bool parse_function(const char *&p, double &out, std::string &err)
{
    static const char *const names[] = { "sin", "cos", "tan", "exp", "ln", "sqrt", nullptr };
    const char *q = p;
    skip_ws(q);
    for (int f = 0; names[f]; ++f) {
        const size_t len = std::strlen(names[f]);
        if (std::strncmp(q, names[f], len) != 0) continue;
        const char *r = q + len;
        skip_ws(r);
        if (*r != '(') continue;
        const char *start = p;
        ++r;
        double x;
        if (!parse_sum(r, x, err)) return false;
        skip_ws(r);
        if (*r != ')') { err = std::string("Unclosed parentheses in expression: \"") + start + "\""; return false; }
        switch (f) {
        case 0: out = std::sin(x); break;
        case 1: out = std::cos(x); break;
        case 2: out = std::tan(x); break;
        case 3: out = std::exp(x); break;
        case 4: out = std::log(x); break;
        default: out = std::sqrt(x); break;
        }
        p = r + 1;
        return true;
    }
    return parse_parenthesized(p, out, err);
}

// expression.rs:200-213 (right associative)
bool parse_power(const char *&p, double &out, std::string &err)
{
    double left;
    if (!parse_function(p, left, err)) return false;
    const char *q = p;
    skip_ws(q);
    if (*q == '^') {
        ++q;
        double right;
        if (!parse_power(q, right, err)) return false;
        out = std::pow(left, right);
        p = q;
        return true;
    }
    out = left;
    return true;
}

// expression.rs:221-241
bool parse_negative(const char *&p, double &out, std::string &err)
{
    bool flip = false;
    const char *q = p;
    for (;;) {
        const char *r = q;
        skip_ws(r);
        if (*r != '-') break;
        q = r + 1;
        flip = !flip;
    }
    double v;
    if (!parse_power(q, v, err)) return false;
    out = flip ? -v : v;
    p = q;
    return true;
}

// expression.rs:249-270
bool parse_product(const char *&p, double &out, std::string &err)
{
    double left;
    if (!parse_negative(p, left, err)) return false;
    for (;;) {
        const char *q = p;
        skip_ws(q);
        if (*q != '*' && *q != '/') break;
        const char op = *q++;
        double right;
        if (!parse_negative(q, right, err)) return false;
        left = op == '*' ? left * right : left / right;
        p = q;
    }
    out = left;
    return true;
}

// expression.rs:278-299
bool parse_sum(const char *&p, double &out, std::string &err)
{
    double left;
    if (!parse_product(p, left, err)) return false;
    for (;;) {
        const char *q = p;
        skip_ws(q);
        if (*q != '+' && *q != '-') break;
        const char op = *q++;
        double right;
        if (!parse_product(q, right, err)) return false;
        left = op == '+' ? left + right : left - right;
        p = q;
    }
    out = left;
    return true;
}

struct Arity { const char *name; int nr_args, nr_bits; };
// composite.rs:287-445
const Arity kTable[] = {
    { "ccrx", 1, 3 }, { "ccry", 1, 3 }, { "ccrz", 1, 3 }, { "ccx", 0, 3 }, { "ccz", 0, 3 }, { "ch", 0, 2 },
    { "crx", 1, 2 }, { "cry", 1, 2 }, { "crz", 1, 2 }, { "cs", 0, 2 }, { "csdg", 0, 2 }, { "ct", 0, 2 },
    { "ctdg", 0, 2 }, { "cu1", 1, 2 }, { "cu2", 2, 2 }, { "cu3", 3, 2 }, { "cv", 0, 2 }, { "cvdg", 0, 2 },
    { "cx", 0, 2 }, { "cy", 0, 2 }, { "cz", 0, 2 }, { "h", 0, 1 }, { "i", 0, 1 }, { "rx", 1, 1 }, { "ry", 1, 1 },
    { "rz", 1, 1 }, { "s", 0, 1 }, { "sdg", 0, 1 }, { "t", 0, 1 }, { "tdg", 0, 1 }, { "swap", 0, 2 },
    { "u1", 1, 1 }, { "u2", 2, 1 }, { "u3", 3, 1 }, { "v", 0, 1 }, { "vdg", 0, 1 }, { "x", 0, 1 }, { "y", 0, 1 },
    { "z", 0, 1 }, { nullptr, 0, 0 } };

}  // namespace

bool parse_expression(const char *text, double &out, const char **rest, std::string &err)
{
    const char *p = text;
    if (!parse_sum(p, out, err)) return false;
    if (rest) *rest = p;
    return true;
}

// composite.rs:216-233 (parse_gate_desc) for one part
static bool parse_gate_desc(const std::string &part, SubGateDesc &g, std::string &err)
{
    const char *p = part.c_str();
    // name: (?i)^\s*([a-z][a-z0-9]*)
    const char *q = p;
    skip_ws(q);
    if (!std::isalpha((unsigned char)*q)) { err = "Failed to find gate name in \"" + part + "\""; return false; }
    const char *e = q;
    while (std::isalnum((unsigned char)*e)) ++e;
    g.name.assign(q, e);
    p = e;
    // args
    g.args.clear();
    q = p;
    skip_ws(q);
    if (*q == '(') {
        const char *args_start = p;
        ++q;
        for (;;) {
            double x;
            if (!parse_sum(q, x, err)) return false;
            g.args.push_back(x);
            const char *r = q;
            skip_ws(r);
            if (*r == ',') { q = r + 1; continue; }
            break;
        }
        skip_ws(q);
        if (*q != ')') { err = std::string("Unclosed parentheses in expression: \"") + args_start + "\""; return false; }
        p = q + 1;
    }
    // bits: (^\s*(\d+))+
    g.bits.clear();
    for (;;) {
        q = p;
        skip_ws(q);
        if (!std::isdigit((unsigned char)*q)) break;
        const char *d = q;
        while (std::isdigit((unsigned char)*d)) ++d;
        if (d - q > 18) { err = "Failed to parse bit number in \"" + std::string(q, d) + "\""; return false; }
        g.bits.push_back((size_t)std::strtoull(std::string(q, d).c_str(), nullptr, 10));
        p = d;
    }
    if (g.bits.empty()) { err = "Unable to find the bits gate " + g.name + " operates on"; return false; }
    q = p;
    skip_ws(q);
    if (*q) {
        std::string rest(q);
        while (!rest.empty() && std::isspace((unsigned char)rest.back())) rest.pop_back();
        err = "Trailing text after gate description: \"" + rest + "\"";
        return false;
    }
    return true;
}

// Composite::from_string, composite.rs:273-450
bool parse_composite(const std::string &desc, std::vector<SubGateDesc> &out, size_t &nr_bits, std::string &err)
{
    out.clear();
    nr_bits = 0;
    size_t pos = 0;
    for (;;) {
        const size_t semi = desc.find(';', pos);
        const std::string part = desc.substr(pos, semi == std::string::npos ? std::string::npos : semi - pos);
        SubGateDesc g;
        if (!parse_gate_desc(part, g, err)) return false;
        for (size_t b : g.bits) nr_bits = std::max(nr_bits, b + 1);
        out.push_back(g);
        if (semi == std::string::npos) break;
        pos = semi + 1;
    }
    for (SubGateDesc &g : out) {
        std::string lower = g.name;
        for (char &ch : lower) ch = (char)std::tolower((unsigned char)ch);
        const Arity *a = nullptr;
        for (const Arity *t = kTable; t->name; ++t)
            if (lower == t->name) { a = t; break; }
        if (!a) { err = "Unknown gate \"" + g.name + "\""; return false; }
        char buf[256];
        if ((size_t)a->nr_args != g.args.size()) {
            std::snprintf(buf, sizeof buf, "Expected %d arguments to \"%s\" gate, got %zu", a->nr_args, g.name.c_str(), g.args.size());
            err = buf;
            return false;
        }
        if ((size_t)a->nr_bits != g.bits.size()) {
            std::snprintf(buf, sizeof buf, "Expected %d bits for \"%s\" gate, got %zu", a->nr_bits, g.name.c_str(), g.bits.size());
            err = buf;
            return false;
        }
        g.name = lower;
    }
    return true;
}

// Flatten: sub-gate bit i of the composite -> bits[i]; the body is appended `repeat` times
// (Loop::apply_slice, staticloop.rs:78-84).  Every sub-gate is validated like a gate added directly.
CircuitError Circuit::add_composite(const std::string &name, const std::string &desc, const std::vector<size_t> &bits, size_t repeat,
                                    bool is_loop)
{
    std::vector<SubGateDesc> subs;
    size_t k = 0;
    std::string err;
    CircuitError e;
    if (!parse_composite(desc, subs, k, err)) { e.code = Q1T_ERR_PARSE; e.msg = err; return e; }
    if (bits.size() != k) {                      // Gate::check_nr_bits, gates.rs:176-186
        char buf[320];
        std::snprintf(buf, sizeof buf, "Expected %zu bits for \"%s\", got %zu", k, name.c_str(), bits.size());
        e.code = Q1T_ERR_INVALID_NR_BITS; e.msg = buf;
        return e;
    }
    const size_t mark = ops_.size();
    for (size_t it = 0; it < repeat; ++it)
        for (const SubGateDesc &g : subs) {
            GateSpec spec;
            std::vector<Param> ps(g.args.size());
            for (size_t a = 0; a < g.args.size(); ++a) ps[a].value = g.args[a];
            if (gate_spec_from_name(g.name.c_str(), ps.data(), ps.size(), spec, err)) { e.code = Q1T_ERR_PARSE; e.msg = err; break; }
            std::vector<size_t> mapped(g.bits.size());
            for (size_t b = 0; b < g.bits.size(); ++b) mapped[b] = bits[g.bits[b]];
            e = add_gate(spec, mapped);
            if (e) break;
        }
    if (e) { ops_.erase(ops_.begin() + (long)mark, ops_.end()); return e; }      // all or nothing
    ++next_group_;
    for (size_t i = mark; i < ops_.size(); ++i) {
        ops_[i].group_id = next_group_;
        ops_[i].group_repeat = repeat;
        ops_[i].group_loop = is_loop;
        ops_[i].group_name = name;
        ops_[i].group_bits = bits;
    }
    return e;
}

// matrix() of a description: apply the sub-gates to the identity, composite.rs:480-485.
// Row-major 2^k x 2^k, gate index MSB = composite bit 0.  Returns k, or -1 (error text in err).
int composite_matrix(const std::string &desc, std::vector<std::complex<double>> &out, std::string &err)
{
    typedef std::complex<double> C;
    std::vector<SubGateDesc> subs;
    size_t k = 0;
    if (!parse_composite(desc, subs, k, err)) return -1;
    if (k > 10) { err = "composite_matrix: more than 10 bits"; return -1; }
    const size_t D = (size_t)1 << k;
    out.assign(D * D, C(0, 0));
    for (size_t a = 0; a < D; ++a) out[a * D + a] = C(1, 0);
    std::vector<C> m(64 * 64), col(8);
    for (const SubGateDesc &g : subs) {
        const int nb = builtin_gate_matrix(g.name.c_str(), g.args.data(), g.args.size(), m.data());
        if (nb < 0) { err = "Unknown gate \"" + g.name + "\""; return -1; }
        const size_t G = (size_t)1 << nb;
        // positions (from the LSB of the row index) of the gate's bits: composite bit b is index bit k-1-b
        std::vector<int> pos(nb);
        for (int j = 0; j < nb; ++j) pos[j] = (int)(k - 1 - g.bits[j]);
        for (int a = 0; a < nb; ++a)
            for (int b = a + 1; b < nb; ++b)
                if (pos[a] == pos[b]) { err = "duplicate bit in sub-gate " + g.name; return -1; }
        // out <- (gate on those bits) * out, column by column
        for (size_t c = 0; c < D; ++c)
            for (size_t base = 0; base < D; ++base) {
                bool lowest = true;
                for (int j = 0; j < nb; ++j) if ((base >> pos[j]) & 1) lowest = false;
                if (!lowest) continue;
                for (size_t gi = 0; gi < G; ++gi) {
                    size_t r = base;
                    for (int j = 0; j < nb; ++j) if ((gi >> (nb - 1 - j)) & 1) r |= (size_t)1 << pos[j];
                    col[gi] = out[r * D + c];
                }
                for (size_t gi = 0; gi < G; ++gi) {
                    C acc(0, 0);
                    for (size_t gj = 0; gj < G; ++gj) acc += m[gi * G + gj] * col[gj];
                    size_t r = base;
                    for (int j = 0; j < nb; ++j) if ((gi >> (nb - 1 - j)) & 1) r |= (size_t)1 << pos[j];
                    out[r * D + c] = acc;
                }
            }
    }
    return (int)k;
}

}  // namespace q1t
