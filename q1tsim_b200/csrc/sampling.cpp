// sampling.cpp -- host-side random sampling of the measurement path.
//
// The reference draws from third-party crates that are not vendored under
// /root/reference: rand 0.7 (`Uniform<f64>`, `Standard`, `WeightedIndex`) and
// rand_distr 0.2 (`Binomial`), call sites vectorstate.rs:123-126, 271-272,
// 379-380.  Their published algorithms are restated here so that the caller's
// generator (q1t_rng = the C mirror of `R: Rng`) is consumed word for word as
// the reference would consume it.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "engine.h"

struct q1t_rng_state {
    int kind;                    // 0 splitmix64, 1 injected words
    uint64_t s;
    std::vector<uint64_t> words;
    size_t pos;
    bool failed;
};

namespace q1t {

static uint64_t builtin_next(void *ctx)
{
    q1t_rng_state *r = static_cast<q1t_rng_state *>(ctx);
    if (r->kind == 1) {
        if (r->pos >= r->words.size()) { r->failed = true; return 0; }
        return r->words[r->pos++];
    }
    r->pos++;
    r->s += 0x9E3779B97F4A7C15ull;
    uint64_t z = r->s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

bool rng_failed(q1t_rng rng)
{
    return rng.next_u64 == builtin_next && static_cast<q1t_rng_state *>(rng.ctx)->failed;
}

// rand 0.7 UniformFloat<f64>::new: scale is lowered ulp by ulp until the largest
// sample stays below `high`
UniformF64 uniform_new(double low, double high)
{
    const double max_rand = 1.0 - 2.220446049250313e-16;
    UniformF64 u{ low, high - low };
    while (u.scale * max_rand + low >= high) u.scale = std::nextafter(u.scale, -INFINITY);
    return u;
}

// rand 0.7 UniformFloat<f64>::sample: 52 mantissa bits into [1,2), minus 1
double uniform_sample(const UniformF64 &u, q1t_rng rng)
{
    const uint64_t bits = (rng.next_u64(rng.ctx) >> 12) | 0x3FF0000000000000ull;
    double v12;
    std::memcpy(&v12, &bits, sizeof v12);
    return (v12 - 1.0) * u.scale + u.low;
}

// the same draw in two halves: value0_1 now, the affine map when low / scale are known (monotone in v)
double uniform_unit(q1t_rng rng)
{
    const uint64_t bits = (rng.next_u64(rng.ctx) >> 12) | 0x3FF0000000000000ull;
    double v12;
    std::memcpy(&v12, &bits, sizeof v12);
    return v12 - 1.0;
}
double uniform_scale(const UniformF64 &u, double v01) { return v01 * u.scale + u.low; }

static double standard_f64(q1t_rng rng)
{
    return (double)(rng.next_u64(rng.ctx) >> 11) * (1.0 / 9007199254740992.0);
}

// f64::powi as compiled by rustc (compiler-rt __powidf2)
static double powi(double a, int b)
{
    const bool recip = b < 0;
    double r = 1.0;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}

// rand_distr 0.2 Binomial: inversion (BINV) below n*min(p,1-p) = 10, otherwise
// BTPE (Kachitvichyanukul & Schmeiser, CACM 31(2), 1988)
uint64_t binomial_sample(q1t_rng rng, uint64_t n_u, double p_in)
{
    if (p_in == 0.0) return 0;
    if (p_in == 1.0) return n_u;
    const double p = p_in <= 0.5 ? p_in : 1.0 - p_in;
    const double q = 1.0 - p;
    uint64_t result;
    if ((double)n_u * p < 10.0 && n_u <= (uint64_t)INT32_MAX) {
        const double s = p / q;
        const double a = (double)(n_u + 1) * s;
        double r = powi(q, (int)n_u);
        double u = standard_f64(rng);
        uint64_t x = 0;
        while (u > r) {
            u -= r;
            x += 1;
            r *= a / (double)x - s;
        }
        result = x;
    } else {
        const double n = (double)n_u;
        const double np = n * p, npq = np * q, f_m = np + p;
        const int64_t m = (int64_t)f_m;
        const double p1 = std::floor(2.195 * std::sqrt(npq) - 4.6 * q) + 0.5;
        const double x_m = (double)m + 0.5, x_l = x_m - p1, x_r = x_m + p1;
        const double c = 0.134 + 20.5 / (15.3 + (double)m);
        const double p2 = p1 * (1. + 2. * c);
        auto lambda = [](double a) { return a * (1. + 0.5 * a); };
        const double lambda_l = lambda((f_m - x_l) / (f_m - x_l * p));
        const double lambda_r = lambda((x_r - f_m) / (x_r * q));
        const double p3 = p2 + c / lambda_l, p4 = p3 + c / lambda_r;
        const UniformF64 gen_u = uniform_new(0., p4), gen_v = uniform_new(0., 1.);
        auto stirling = [](double a) {
            const double a2 = a * a;
            return (13860. - (462. - (132. - (99. - 140. / a2) / a2) / a2) / a2) / a / 166320.;
        };
        int64_t y;
        for (;;) {
            if (rng_failed(rng)) return 0;
            const double u = uniform_sample(gen_u, rng);
            double v = uniform_sample(gen_v, rng);
            if (!(u > p1)) { y = (int64_t)(x_m - p1 * v + u); break; }
            if (!(u > p2)) {
                const double x = x_l + (u - p1) / c;
                v = v * c + 1.0 - std::fabs(x - x_m) / p1;
                if (v > 1.) continue;
                y = (int64_t)x;
            } else if (!(u > p3)) {
                y = (int64_t)(x_l + std::log(v) / lambda_l);
                if (y < 0) continue;
                v *= (u - p2) * lambda_l;
            } else {
                y = (int64_t)(x_r - std::log(v) / lambda_r);
                if (y > 0 && (uint64_t)y > n_u) continue;
                v *= (u - p3) * lambda_r;
            }
            const int64_t k = std::llabs(y - m);
            if (!(k > 20 && (double)k < 0.5 * npq - 1.)) {
                const double s = p / q, a = s * (n + 1.);
                double f = 1.0;
                if (m < y) { for (int64_t i = m + 1;; ++i) { f *= a / (double)i - s; if (i == y) break; } }
                else if (m > y) { for (int64_t i = y + 1;; ++i) { f /= a / (double)i - s; if (i == m) break; } }
                if (v > f) continue;
                break;
            }
            const double kf = (double)k;
            const double rho = (kf / npq) * ((kf * (kf / 3. + 0.625) + 1. / 6.) / npq + 0.5);
            const double t = -0.5 * kf * kf / npq;
            const double alpha = std::log(v);
            if (alpha < t - rho) break;
            if (alpha > t + rho) continue;
            const double x1 = (double)(y + 1), f1 = (double)(m + 1);
            const double z = (double)((int64_t)n + 1 - m), w = (double)((int64_t)n - y + 1);
            if (alpha > x_m * std::log(f1 / x1) + (n - (double)m + 0.5) * std::log(z / w)
                            + (double)(y - m) * std::log(w * p / (x1 * q))
                            + stirling(f1) + stirling(z) - stirling(x1) - stirling(w))
                continue;
            break;
        }
        result = (uint64_t)y;
    }
    return p != p_in ? n_u - result : result;
}

}  // namespace q1t

extern "C" {

q1t_rng_state *q1t_rng_splitmix64(uint64_t seed)
{
    q1t_rng_state *r = new q1t_rng_state();
    r->kind = 0; r->s = seed; r->pos = 0; r->failed = false;
    return r;
}
q1t_rng_state *q1t_rng_from_words(const uint64_t *words, size_t n)
{
    q1t_rng_state *r = new q1t_rng_state();
    r->kind = 1; r->s = 0; r->pos = 0; r->failed = false;
    if (words && n) r->words.assign(words, words + n);
    return r;
}
q1t_rng_state *q1t_rng_entropy(void)
{
    std::random_device rd;
    const uint64_t seed = ((uint64_t)rd() << 32) ^ (uint64_t)rd();
    return q1t_rng_splitmix64(seed);
}
size_t q1t_rng_consumed(const q1t_rng_state *r) { return r ? r->pos : 0; }
void q1t_rng_free(q1t_rng_state *r) { delete r; }
q1t_rng q1t_rng_handle(q1t_rng_state *r)
{
    q1t_rng h;
    h.next_u64 = q1t::builtin_next;
    h.ctx = r;
    return h;
}
uint64_t q1t_binomial(q1t_rng rng, uint64_t n, double p) { return q1t::binomial_sample(rng, n, p); }

}  // extern "C"
