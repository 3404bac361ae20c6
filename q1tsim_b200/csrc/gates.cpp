// gates.cpp -- `matrix()` of the reference's built-in gates, by name.
// Name table: Composite::from_string, composite.rs:287-445 (a superset of what
// ffi.rs:335-367 accepts).  Definitions: src/gates/<gate>.rs `fn matrix`.
#include <cmath>
#include <complex>
#include <cstring>
#include <string>

#include "engine.h"

namespace q1t {

typedef std::complex<double> C;

static C polar(double r, double th) { return C(r * std::cos(th), r * std::sin(th)); }

// one-qubit base gates and swap; returns #qubits, -1 unknown, -2 bad arity
static int base_gate(const std::string &nm, const double *p, size_t np, C *m)
{
    const double x = 0.70710678118654752440;      // std::f64::consts::FRAC_1_SQRT_2
    const C z(0, 0), o(1, 0), i(0, 1);
    auto arity = [&](size_t want) { return np == want; };
    auto set2 = [&](C a, C b, C c, C d) { m[0] = a; m[1] = b; m[2] = c; m[3] = d; return 1; };
    if (nm == "h") return arity(0) ? set2(C(x, 0), C(x, 0), C(x, 0), C(-x, -0.0)) : -2;        // hadamard.rs:89-93
    if (nm == "i") return arity(0) ? set2(o, z, z, o) : -2;                                     // identity.rs
    if (nm == "x") return arity(0) ? set2(z, o, o, z) : -2;                                     // x.rs
    if (nm == "y") return arity(0) ? set2(z, -i, i, z) : -2;                                    // y.rs:53-58
    if (nm == "z") return arity(0) ? set2(o, z, z, -o) : -2;                                    // z.rs
    if (nm == "s") return arity(0) ? set2(o, z, z, i) : -2;                                     // s.rs:60-66
    if (nm == "sdg") return arity(0) ? set2(o, z, z, -i) : -2;
    if (nm == "t") return arity(0) ? set2(o, z, z, C(x, x)) : -2;                               // t.rs:52-59
    if (nm == "tdg") return arity(0) ? set2(o, z, z, C(x, -x)) : -2;
    if (nm == "v") return arity(0) ? set2(C(.5, .5), C(.5, -.5), C(.5, -.5), C(.5, .5)) : -2;   // v.rs:58-63
    if (nm == "vdg") return arity(0) ? set2(C(.5, -.5), C(.5, .5), C(.5, .5), C(.5, -.5)) : -2; // v.rs:157-162
    if (nm == "rx") {                                                                           // rx.rs:64-70
        if (!arity(1)) return -2;
        const double h = 0.5 * p[0];
        const C c(std::cos(h), 0), si(0, std::sin(h));
        return set2(c, -si, -si, c);
    }
    if (nm == "ry") {                                                                           // ry.rs:64-70
        if (!arity(1)) return -2;
        const double h = 0.5 * p[0];
        const C c(std::cos(h), 0), s(std::sin(h), 0);
        return set2(c, -s, s, c);
    }
    if (nm == "rz") {                                                                           // rz.rs:65-70
        if (!arity(1)) return -2;
        const C q = polar(1.0, 0.5 * p[0]);
        return set2(std::conj(q), z, z, q);
    }
    if (nm == "u1") return arity(1) ? set2(o, z, z, polar(1.0, p[0])) : -2;                     // u1.rs:69-75
    if (nm == "u2") {                                                                           // u2.rs:70-79
        if (!arity(2)) return -2;
        return set2(C(x, 0), -polar(x, p[1]), polar(x, p[0]), polar(x, p[0] + p[1]));
    }
    if (nm == "u3") {                                                                           // u3.rs:74-84
        if (!arity(3)) return -2;
        const double h = 0.5 * p[0], c = std::cos(h), s = std::sin(h);
        return set2(C(c, 0), -polar(s, p[2]), polar(s, p[1]), polar(c, p[1] + p[2]));
    }
    if (nm == "swap") {                                                                         // swap.rs:78-88
        if (!arity(0)) return -2;
        for (int a = 0; a < 16; ++a) m[a] = z;
        m[0] = o; m[6] = o; m[9] = o; m[15] = o;
        return 2;
    }
    return -1;
}

int builtin_gate_matrix(const char *name, const double *params, size_t nparams, C *out)
{
    static const char *const table[] = { "h", "i", "s", "sdg", "t", "tdg", "v", "vdg", "x", "y", "z", "rx", "ry", "rz",
        "u1", "u2", "u3", "cx", "cy", "cz", "ch", "cs", "csdg", "ct", "ctdg", "cv", "cvdg", "swap", "crx", "cry",
        "crz", "cu1", "cu2", "cu3", "ccx", "ccz", "ccrx", "ccry", "ccrz", nullptr };
    std::string nm(name ? name : "");
    for (char &ch : nm) if (ch >= 'A' && ch <= 'Z') ch = (char)(ch + 32);
    bool known = false;
    for (int a = 0; table[a]; ++a) if (nm == table[a]) known = true;
    if (!known) return -1;
    C base[16];
    int nctl = 0;
    std::string rest = nm;
    int nb = base_gate(rest, params, nparams, base);
    while (nb == -1 && rest.size() > 1 && rest[0] == 'c') {     // C<G> = I (+) G, controlled.rs:60-69
        rest = rest.substr(1);
        ++nctl;
        nb = base_gate(rest, params, nparams, base);
    }
    if (nb < 0) return nb;
    const size_t G = (size_t)1 << (nb + nctl), g0 = (size_t)1 << nb;
    for (size_t a = 0; a < G * G; ++a) out[a] = C(0, 0);
    for (size_t a = 0; a < G; ++a) out[a * G + a] = C(1, 0);
    for (size_t r = 0; r < g0; ++r)
        for (size_t c = 0; c < g0; ++c) out[(G - g0 + r) * G + (G - g0 + c)] = base[r * g0 + c];
    return nb + nctl;
}

}  // namespace q1t
