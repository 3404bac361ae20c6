/*
 * q1t_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See q1t_oracle.h for the status header.  Each function cites the reference
 * (Q1tBV/q1tsim 0.5.0) file:line it restates.  Compile with
 * -ffp-contract=off so that complex arithmetic is the textbook
 * (ac-bd, ad+bc) / re*re+im*im of num-complex 0.2 (no FMA).
 */
#include "q1t_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;
void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

/* ------------------------------------------------------------------ */
/* complex helpers: num-complex 0.2 Complex<f64> Mul/Add/norm_sqr       */
/* ------------------------------------------------------------------ */
static inline orc_cplx cmul(orc_cplx a, orc_cplx b)
{
    orc_cplx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
    return r;
}
static inline orc_cplx cadd(orc_cplx a, orc_cplx b)
{
    orc_cplx r = { a.re + b.re, a.im + b.im };
    return r;
}
static inline double cnorm_sqr(orc_cplx a) { return a.re * a.re + a.im * a.im; }

/* ------------------------------------------------------------------ */
/* RNG: SplitMix64 or injected raw words            */
/* ------------------------------------------------------------------ */
void orc_rng_seed(orc_rng *r, uint64_t seed)
{
    memset(r, 0, sizeof *r);
    r->kind = 0;
    r->s = seed;
}
void orc_rng_array(orc_rng *r, const uint64_t *arr, size_t n)
{
    memset(r, 0, sizeof *r);
    r->kind = 1;
    r->arr = arr;
    r->n = n;
}
uint64_t orc_rng_next(orc_rng *r)
{
    if (r->kind == 1) {
        if (r->pos >= r->n) { r->exhausted = 1; return 0; }
        return r->arr[r->pos++];
    }
    r->s += 0x9E3779B97F4A7C15ull;
    uint64_t z = r->s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* rand 0.7 `Standard` for f64: 53 random bits, [0,1) */
static double rng_f64_standard(orc_rng *r)
{
    return (double)(orc_rng_next(r) >> 11) * (1.0 / 9007199254740992.0);
}

/* rand 0.7 UniformFloat<f64>::new(low, high) / sample():
 * value1_2 = float with exponent 0 and 52 random mantissa bits; result =
 * (value1_2 - 1.0) * scale + low, where scale is decreased by ulps until
 * scale * max_rand + low < high. */
typedef struct { double low, scale; } uni_f64;
static uni_f64 uniform_new(double low, double high)
{
    const double max_rand = 1.0 - 2.220446049250313e-16; /* (2^52-1)/2^52 */
    uni_f64 u;
    u.low = low;
    u.scale = high - low;
    while (u.scale * max_rand + low >= high) {
        u.scale = nextafter(u.scale, -INFINITY);
    }
    return u;
}
static double uniform_sample(const uni_f64 *u, orc_rng *r)
{
    uint64_t x = orc_rng_next(r) >> 12;
    uint64_t bits = x | 0x3FF0000000000000ull;
    double v12;
    memcpy(&v12, &bits, sizeof v12);
    return (v12 - 1.0) * u->scale + u->low;
}

/* compiler-rt __powidf2, which is what Rust's f64::powi lowers to */
static double powi_rt(double a, int b)
{
    const int recip = b < 0;
    double r = 1.0;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}

static double btpe_lambda(double a) { return a * (1.0 + 0.5 * a); }
static double btpe_stirling(double a)
{
    double a2 = a * a;
    return (13860. - (462. - (132. - (99. - 140. / a2) / a2) / a2) / a2) / a / 166320.;
}

/* rand_distr 0.2 Binomial::sample (BINV for n*min(p,1-p) < 10, else BTPE,
 * Kachitvichyanukul & Schmeiser 1988).  Call sites: vectorstate.rs:271-272,
 * 379-380.  Third-party algorithm restated from its published form:
 * PARITY UNPINNED. */
uint64_t orc_binomial(orc_rng *rng, uint64_t n_u, double p_in)
{
    if (p_in == 0.0) return 0;
    if (p_in == 1.0) return n_u;
    double p = p_in <= 0.5 ? p_in : 1.0 - p_in;
    double q = 1.0 - p;
    uint64_t result;
    if ((double)n_u * p < 10.0 && n_u <= (uint64_t)INT32_MAX) {
        double s = p / q;
        double a = (double)(n_u + 1) * s;
        double r = powi_rt(q, (int)n_u);
        double u = rng_f64_standard(rng);
        uint64_t x = 0;
        while (u > r) {
            u -= r;
            x += 1;
            r *= a / (double)x - s;
        }
        result = x;
    } else {
        const int64_t SQUEEZE_THRESHOLD = 20;
        double n = (double)n_u;
        double np = n * p;
        double npq = np * q;
        double f_m = np + p;
        int64_t m = (int64_t)f_m;
        double p1 = floor(2.195 * sqrt(npq) - 4.6 * q) + 0.5;
        double x_m = (double)m + 0.5;
        double x_l = x_m - p1;
        double x_r = x_m + p1;
        double c = 0.134 + 20.5 / (15.3 + (double)m);
        double p2 = p1 * (1. + 2. * c);
        double lambda_l = btpe_lambda((f_m - x_l) / (f_m - x_l * p));
        double lambda_r = btpe_lambda((x_r - f_m) / (x_r * q));
        double p3 = p2 + c / lambda_l;
        double p4 = p3 + c / lambda_r;
        int64_t y;
        uni_f64 gen_u = uniform_new(0., p4);
        uni_f64 gen_v = uniform_new(0., 1.);
        for (;;) {
            if (rng->exhausted) return 0;
            double u = uniform_sample(&gen_u, rng);
            double v = uniform_sample(&gen_v, rng);
            if (!(u > p1)) {
                y = (int64_t)(x_m - p1 * v + u);
                break;
            }
            if (!(u > p2)) {
                double x = x_l + (u - p1) / c;
                v = v * c + 1.0 - fabs(x - x_m) / p1;
                if (v > 1.) continue;
                y = (int64_t)x;
            } else if (!(u > p3)) {
                y = (int64_t)(x_l + log(v) / lambda_l);
                if (y < 0) continue;
                v *= (u - p2) * lambda_l;
            } else {
                y = (int64_t)(x_r - log(v) / lambda_r);
                if (y > 0 && (uint64_t)y > n_u) continue;
                v *= (u - p3) * lambda_r;
            }
            int64_t k = llabs(y - m);
            if (!(k > SQUEEZE_THRESHOLD && (double)k < 0.5 * npq - 1.)) {
                double s = p / q;
                double a = s * (n + 1.);
                double f = 1.0;
                if (m < y) {
                    int64_t i = m;
                    for (;;) { i += 1; f *= a / (double)i - s; if (i == y) break; }
                } else if (m > y) {
                    int64_t i = y;
                    for (;;) { i += 1; f /= a / (double)i - s; if (i == m) break; }
                }
                if (v > f) continue;
                break;
            }
            double kf = (double)k;
            double rho = (kf / npq) * ((kf * (kf / 3. + 0.625) + 1. / 6.) / npq + 0.5);
            double t = -0.5 * kf * kf / npq;
            double alpha = log(v);
            if (alpha < t - rho) break;
            if (alpha > t + rho) continue;
            double x1 = (double)(y + 1);
            double f1 = (double)(m + 1);
            double z = (double)((int64_t)n + 1 - m);
            double w = (double)((int64_t)n - y + 1);
            if (alpha > x_m * log(f1 / x1) + (n - (double)m + 0.5) * log(z / w)
                          + (double)(y - m) * log(w * p / (x1 * q))
                          + btpe_stirling(f1) + btpe_stirling(z) - btpe_stirling(x1) - btpe_stirling(w))
                continue;
            break;
        }
        result = (uint64_t)y;
    }
    return (p != p_in) ? n_u - result : result;
}

/* ------------------------------------------------------------------ */
/* support.rs:50-75                                                     */
/* ------------------------------------------------------------------ */
uint64_t orc_reverse_bits(uint64_t idx, size_t nr_bits)
{
    uint64_t res = 0, sidx = idx;
    for (size_t i = 0; i < nr_bits; i++) {
        res |= (sidx & 1) << (nr_bits - 1 - i);
        sidx >>= 1;
    }
    return res;
}
uint64_t orc_shuffle_bits(uint64_t idx, const size_t *bits, size_t n)
{
    uint64_t res = 0, sidx = idx;
    for (size_t i = 0; i < n; i++) {
        res |= (sidx & 1) << bits[i];
        sidx >>= 1;
    }
    return res;
}

/* qustate.rs:100-127: run-length encode the per-shot control flags inside
 * each column; returns number of ranges. */
size_t orc_collect_conditional_ranges(const size_t *counts, size_t ncols, const uint8_t *control,
                                      size_t *out_icol, size_t *out_len, uint8_t *out_apply)
{
    size_t nr = 0, off = 0;
    for (size_t icol = 0; icol < ncols; icol++) {
        size_t count = counts[icol];
        if (count == 0) continue; /* reference would index control[off] out of run; no shots -> no range */
        size_t begin = off;
        uint8_t prev = control[off] != 0;
        for (size_t ibit = off + 1; ibit < off + count; ibit++) {
            if ((control[ibit] != 0) != prev) {
                out_icol[nr] = icol; out_len[nr] = ibit - begin; out_apply[nr] = prev; nr++;
                begin = ibit;
                prev = !prev;
            }
        }
        if (begin < off + count) {
            out_icol[nr] = icol; out_len[nr] = off + count - begin; out_apply[nr] = prev; nr++;
        }
        off += count;
    }
    return nr;
}

/* gates.rs:53-80 + permutation.rs:38-89: perm_out[new] = old */
int orc_bit_permutation(size_t nr_bits, const size_t *bits, size_t k, size_t *perm_out)
{
    size_t N = (size_t)1 << nr_bits;
    size_t *perm1 = malloc(N * sizeof(size_t));
    size_t *ab = malloc((k ? k : 1) * sizeof(size_t));
    if (!perm1 || !ab) { free(perm1); free(ab); return -1; }
    for (size_t i = 0; i < N; i++) perm1[i] = i;
    memcpy(ab, bits, k * sizeof(size_t));
    size_t nab = k;
    while (nab > 0) {
        size_t s = ab[--nab];
        size_t idx = nr_bits - s - 1;
        size_t bit = (size_t)1 << idx;
        size_t lmask = bit - 1;
        size_t umask = ~(bit | lmask);
        for (size_t i = 0; i < N; i++) {
            size_t v = perm1[i];
            perm1[i] = ((v & umask) >> 1) | (v & lmask) | ((v & bit) << s);
        }
        for (size_t j = 0; j < nab; j++)
            if (ab[j] < s) ab[j] += 1;
    }
    /* Permutation::new validation (seen[]) then inverse() (validated again) */
    unsigned char *seen = calloc(N, 1);
    int ok = seen != NULL;
    for (size_t i = 0; ok && i < N; i++) {
        if (perm1[i] >= N || seen[perm1[i]]) ok = 0; else seen[perm1[i]] = 1;
    }
    if (ok) {
        for (size_t i = 0; i < N; i++) perm_out[perm1[i]] = i;
        memset(seen, 0, N);
        for (size_t i = 0; ok && i < N; i++) {
            if (seen[perm_out[i]]) ok = 0; else seen[perm_out[i]] = 1;
        }
    }
    free(seen); free(perm1); free(ab);
    return ok ? 0 : -1;
}

/* ------------------------------------------------------------------ */
/* state: vectorstate.rs:25-83, 410-415                                 */
/* ------------------------------------------------------------------ */
orc_state *orc_state_new(size_t nr_bits, size_t nr_shots)
{
    orc_state *s = calloc(1, sizeof *s);
    size_t N = (size_t)1 << nr_bits;
    s->nr_bits = nr_bits; s->nr_shots = nr_shots; s->ncols = 1;
    s->counts = malloc(sizeof(size_t));
    s->counts[0] = nr_shots;
    s->states = calloc(N, sizeof(orc_cplx));
    s->states[0].re = 1.0;
    return s;
}

/* vectorstate.rs:62-83: kron chain over per-qubit normalised coefficients;
 * kron_mat (cmatrix.rs:40-51) multiplies a1 * a0[[i,j]] (element * scalar). */
orc_state *orc_state_from_qubit_coefs(const double *coefs, size_t nr_bits, size_t nr_shots)
{
    orc_state *s = calloc(1, sizeof *s);
    size_t N = (size_t)1 << nr_bits;
    s->nr_bits = nr_bits; s->nr_shots = nr_shots; s->ncols = 1;
    s->counts = malloc(sizeof(size_t));
    s->counts[0] = nr_shots;
    orc_cplx *cur = malloc(N * sizeof(orc_cplx));
    orc_cplx *nxt = malloc(N * sizeof(orc_cplx));
    cur[0].re = 1.0; cur[0].im = 0.0;
    size_t len = 1;
    for (size_t b = 0; b < nr_bits; b++) {
        orc_cplx c0 = { coefs[4 * b], coefs[4 * b + 1] }, c1 = { coefs[4 * b + 2], coefs[4 * b + 3] };
        double norm = sqrt(cnorm_sqr(c0) + cnorm_sqr(c1));
        /* Complex / f64 divides both parts */
        orc_cplx b0 = { c0.re / norm, c0.im / norm }, b1 = { c1.re / norm, c1.im / norm };
        for (size_t i = 0; i < len; i++) {
            nxt[2 * i] = cmul(b0, cur[i]);
            nxt[2 * i + 1] = cmul(b1, cur[i]);
        }
        len *= 2;
        orc_cplx *t = cur; cur = nxt; nxt = t;
    }
    free(nxt);
    s->states = cur;
    return s;
}

void orc_state_free(orc_state *s)
{
    if (!s) return;
    free(s->counts); free(s->states); free(s);
}
size_t orc_state_ncols(const orc_state *s) { return s->ncols; }
void orc_state_counts(const orc_state *s, size_t *out) { memcpy(out, s->counts, s->ncols * sizeof(size_t)); }
void orc_state_read_column(const orc_state *s, size_t col, double *out)
{
    size_t N = (size_t)1 << s->nr_bits, C = s->ncols;
    for (size_t i = 0; i < N; i++) { out[2 * i] = s->states[i * C + col].re; out[2 * i + 1] = s->states[i * C + col].im; }
}
void orc_state_write_column(orc_state *s, size_t col, const double *in)
{
    size_t N = (size_t)1 << s->nr_bits, C = s->ncols;
    for (size_t i = 0; i < N; i++) { s->states[i * C + col].re = in[2 * i]; s->states[i * C + col].im = in[2 * i + 1]; }
}
void orc_reset_all(orc_state *s)
{
    size_t N = (size_t)1 << s->nr_bits;
    free(s->states); free(s->counts);
    s->states = calloc(N, sizeof(orc_cplx));
    s->states[0].re = 1.0;
    s->counts = malloc(sizeof(size_t));
    s->counts[0] = s->nr_shots;
    s->ncols = 1;
}

/* ------------------------------------------------------------------ */
/* gate application                                                     */
/* ------------------------------------------------------------------ */

/* Gate::apply_mat_slice default (gates.rs:273-326) on a block of `rows`
 * rows x C columns starting at `st` (row stride C), FAITHFUL structure:
 * owned copies of the sub-blocks and one fresh temporary per product, as
 * ndarray's `&s0*m + &s1*m` does. */
static void dense_block_faithful(orc_cplx *st, size_t rows, size_t C, const orc_cplx *mat, size_t k)
{
    size_t G = (size_t)1 << k;
    size_t n = rows >> k;          /* rows per sub-block */
    size_t bl = n * C;             /* elements per sub-block */
    if (k <= 2) {
        orc_cplx **s = malloc(G * sizeof *s);
        for (size_t g = 0; g < G; g++) {          /* .to_owned() copies */
            s[g] = malloc(bl * sizeof(orc_cplx));
            memcpy(s[g], st + g * bl, bl * sizeof(orc_cplx));
        }
        for (size_t i = 0; i < G; i++) {
            orc_cplx *acc = malloc(bl * sizeof(orc_cplx));   /* &s0 * m -> new array */
            for (size_t e = 0; e < bl; e++) acc[e] = cmul(s[0][e], mat[i * G + 0]);
            for (size_t j = 1; j < G; j++) {
                orc_cplx *t = malloc(bl * sizeof(orc_cplx)); /* &sj * m -> new array */
                for (size_t e = 0; e < bl; e++) t[e] = cmul(s[j][e], mat[i * G + j]);
                for (size_t e = 0; e < bl; e++) acc[e] = cadd(acc[e], t[e]);
                free(t);
            }
            memcpy(st + i * bl, acc, bl * sizeof(orc_cplx)); /* .assign() */
            free(acc);
        }
        for (size_t g = 0; g < G; g++) free(s[g]);
        free(s);
    } else {
        orc_cplx *res = calloc(rows * C, sizeof(orc_cplx));
        for (size_t i = 0; i < G; i++) {
            for (size_t j = 0; j < G; j++) {
                orc_cplx *x = malloc(bl * sizeof(orc_cplx));  /* slice.to_owned() * m */
                for (size_t e = 0; e < bl; e++) x[e] = cmul(st[j * bl + e], mat[i * G + j]);
                for (size_t e = 0; e < bl; e++) res[i * bl + e] = cadd(res[i * bl + e], x[e]);
                free(x);
            }
        }
        memcpy(st, res, rows * C * sizeof(orc_cplx));
        free(res);
    }
}

/* gates.rs:121-152 (apply_gate_mat_slice), faithful structure */
static int apply_faithful(orc_cplx *states, size_t nr_bits, size_t C, const orc_cplx *mat,
                          const size_t *bits, size_t k)
{
    size_t N = (size_t)1 << nr_bits;
    if (k == 1) {
        size_t bit = bits[0];
        size_t block_size = (size_t)1 << (nr_bits - bit);
        size_t nr_blocks = (size_t)1 << bit;
        for (size_t i = 0; i < nr_blocks; i++)
            dense_block_faithful(states + i * block_size * C, block_size, C, mat, 1);
        return 0;
    }
    size_t *perm = malloc(N * sizeof(size_t));
    if (!perm) return -1;
    if (orc_bit_permutation(nr_bits, bits, k, perm) != 0) { free(perm); return -1; }
    orc_cplx *work = calloc(N * C, sizeof(orc_cplx));
    for (size_t c = 0; c < C; c++)                 /* perm.apply_vec_into per column */
        for (size_t ni = 0; ni < N; ni++) work[ni * C + c] = states[perm[ni] * C + c];
    dense_block_faithful(work, N, C, mat, k);
    for (size_t c = 0; c < C; c++)                 /* perm.apply_inverse_vec_into */
        for (size_t ni = 0; ni < N; ni++) states[perm[ni] * C + c] = work[ni * C + c];
    free(work); free(perm);
    return 0;
}

/* Same arithmetic (per output element: ((s0*m0 + s1*m1) + s2*m2) ..., and for
 * k>2 starting from 0 as gates.rs:310-325 does), direct strided loops.
 * Gate index g: bits[0] is its MSB (SURVEY 3.2 closed form of bit_permutation). */
static void apply_fast(orc_cplx *states, size_t nr_bits, size_t C, const orc_cplx *mat,
                       const size_t *bits, size_t k)
{
    size_t N = (size_t)1 << nr_bits;
    size_t G = (size_t)1 << k;
    size_t pos[64];
    size_t sorted[64];
    for (size_t j = 0; j < k; j++) { pos[j] = nr_bits - 1 - bits[j]; sorted[j] = pos[j]; }
    for (size_t a = 0; a < k; a++)
        for (size_t b = a + 1; b < k; b++)
            if (sorted[b] < sorted[a]) { size_t t = sorted[a]; sorted[a] = sorted[b]; sorted[b] = t; }
    size_t *offs = malloc(G * sizeof(size_t));
    for (size_t g = 0; g < G; g++) {
        size_t o = 0;
        for (size_t j = 0; j < k; j++)
            if ((g >> (k - 1 - j)) & 1) o |= (size_t)1 << pos[j];
        offs[g] = o;
    }
    size_t ngroups = N >> k;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1 && ngroups >= 4096)
#endif
    for (size_t grp = 0; grp < ngroups; grp++) {
        size_t base = grp;
        for (size_t a = 0; a < k; a++) {            /* insert zero bits at sorted positions */
            size_t low = base & (((size_t)1 << sorted[a]) - 1);
            base = ((base >> sorted[a]) << (sorted[a] + 1)) | low;
        }
        orc_cplx in[64], out[64];
        for (size_t c = 0; c < C; c++) {
            if (G <= 64) {
                for (size_t g = 0; g < G; g++) in[g] = states[(base | offs[g]) * C + c];
                for (size_t i = 0; i < G; i++) {
                    orc_cplx acc;
                    if (k <= 2) {
                        acc = cmul(in[0], mat[i * G]);
                        for (size_t j = 1; j < G; j++) acc = cadd(acc, cmul(in[j], mat[i * G + j]));
                    } else {
                        acc.re = 0.0; acc.im = 0.0;
                        for (size_t j = 0; j < G; j++) acc = cadd(acc, cmul(in[j], mat[i * G + j]));
                    }
                    out[i] = acc;
                }
                for (size_t g = 0; g < G; g++) states[(base | offs[g]) * C + c] = out[g];
            } else {
                orc_cplx *bi = malloc(G * sizeof(orc_cplx)), *bo = malloc(G * sizeof(orc_cplx));
                for (size_t g = 0; g < G; g++) bi[g] = states[(base | offs[g]) * C + c];
                for (size_t i = 0; i < G; i++) {
                    orc_cplx acc = { 0.0, 0.0 };
                    for (size_t j = 0; j < G; j++) acc = cadd(acc, cmul(bi[j], mat[i * G + j]));
                    bo[i] = acc;
                }
                for (size_t g = 0; g < G; g++) states[(base | offs[g]) * C + c] = bo[g];
                free(bi); free(bo);
            }
        }
    }
    free(offs);
}

static int check_bits(const orc_state *s, const size_t *bits, size_t k)
{
    for (size_t j = 0; j < k; j++) if (bits[j] >= s->nr_bits) return ORC_ERR_INVALID_QBIT;
    return ORC_OK;
}

/* vectorstate.rs:166-178 */
int orc_apply_gate(orc_state *s, const double *mat, const size_t *bits, size_t k, int mode)
{
    int e = check_bits(s, bits, k);
    if (e) return e;
    if (mode == 0) return apply_faithful(s->states, s->nr_bits, s->ncols, (const orc_cplx *)mat, bits, k);
    apply_fast(s->states, s->nr_bits, s->ncols, (const orc_cplx *)mat, bits, k);
    return ORC_OK;
}

/* vectorstate.rs:180-189 */
int orc_apply_unary_gate_all(orc_state *s, const double *mat, int mode)
{
    for (size_t b = 0; b < s->nr_bits; b++) {
        int e = orc_apply_gate(s, mat, &b, 1, mode);
        if (e) return e;
    }
    return ORC_OK;
}

/* vectorstate.rs:193-227 */
int orc_apply_conditional_gate(orc_state *s, const uint8_t *control, size_t ncontrol,
                               const double *mat, const size_t *bits, size_t k, int mode)
{
    if (ncontrol != s->nr_shots) return ORC_ERR_INVALID_NR_CONTROL_BITS;
    int e = check_bits(s, bits, k);
    if (e) return e;
    size_t N = (size_t)1 << s->nr_bits, C = s->ncols;
    size_t maxr = s->nr_shots + C;
    size_t *ricol = malloc(maxr * sizeof(size_t)), *rlen = malloc(maxr * sizeof(size_t));
    uint8_t *rapply = malloc(maxr);
    size_t nr = orc_collect_conditional_ranges(s->counts, C, control, ricol, rlen, rapply);
    orc_cplx *ns = calloc(N * (nr ? nr : 1), sizeof(orc_cplx));
    orc_cplx *col = malloc(N * sizeof(orc_cplx));
    for (size_t r = 0; r < nr; r++) {
        for (size_t i = 0; i < N; i++) col[i] = s->states[i * C + ricol[r]];
        if (rapply[r]) {
            /* gates.rs:86-115 (vector path) -- same arithmetic as the matrix path with C=1 */
            if (mode == 0) apply_faithful(col, s->nr_bits, 1, (const orc_cplx *)mat, bits, k);
            else apply_fast(col, s->nr_bits, 1, (const orc_cplx *)mat, bits, k);
        }
        for (size_t i = 0; i < N; i++) ns[i * nr + r] = col[i];
    }
    free(col);
    free(s->states); s->states = ns;
    free(s->counts); s->counts = malloc((nr ? nr : 1) * sizeof(size_t));
    for (size_t r = 0; r < nr; r++) s->counts[r] = rlen[r];
    s->ncols = nr;
    free(ricol); free(rlen); free(rapply);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* reductions                                                           */
/* ------------------------------------------------------------------ */

/* CANONICAL blocked order: leaves of LEAF=min(N,1024) consecutive
 * amplitudes; inside a leaf, element e goes to lane e%32 and lanes accumulate
 * sequentially in e/32 order, then a 5-stage xor butterfly (16,8,4,2,1)
 * combines the 32 lanes.  Leaf totals are chained sequentially inside blocks
 * of 1024 leaves, block totals are chained sequentially.  `mask`/`want`
 * select amplitudes (others count as +0.0). */
#define CANON_LEAF 1024
#define CANON_BLOCK 1024

static double canon_leaf_total(const orc_cplx *states, size_t C, size_t col, size_t first, size_t leaf,
                               size_t mask, size_t want)
{
    double acc[32];
    for (int l = 0; l < 32; l++) acc[l] = 0.0;
    for (size_t e = 0; e < leaf; e++) {
        size_t idx = first + e;
        double p = ((idx & mask) == want) ? cnorm_sqr(states[idx * C + col]) : 0.0;
        acc[e & 31] += p;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        double nw[32];
        for (int l = 0; l < 32; l++) nw[l] = acc[l] + acc[l ^ off];
        memcpy(acc, nw, sizeof acc);
    }
    return acc[0];
}

/* inclusive canonical prefix over leaves; returns malloc'd array of nleaves */
static double *canon_leaf_prefix(const orc_state *s, size_t col, size_t mask, size_t want, size_t *nleaves_out)
{
    size_t N = (size_t)1 << s->nr_bits, C = s->ncols;
    size_t leaf = N < CANON_LEAF ? N : CANON_LEAF;
    size_t nleaves = N / leaf;
    double *P = malloc(nleaves * sizeof(double));
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1 && nleaves >= 64)
#endif
    for (size_t L = 0; L < nleaves; L++) P[L] = canon_leaf_total(s->states, C, col, L * leaf, leaf, mask, want);
    double bprefix = 0.0;
    for (size_t b0 = 0; b0 < nleaves; b0 += CANON_BLOCK) {
        size_t b1 = b0 + CANON_BLOCK < nleaves ? b0 + CANON_BLOCK : nleaves;
        double run = 0.0;
        for (size_t L = b0; L < b1; L++) { run += P[L]; P[L] = run; }   /* in-block inclusive */
        double btotal = run;
        for (size_t L = b0; L < b1; L++) P[L] = bprefix + P[L];
        bprefix = bprefix + btotal;     /* == P[b1-1] */
    }
    *nleaves_out = nleaves;
    return P;
}

static double canon_total(const orc_state *s, size_t col, size_t mask, size_t want)
{
    size_t nl;
    double *P = canon_leaf_prefix(s, col, mask, want, &nl);
    double t = P[nl - 1];
    free(P);
    return t;
}

/* vectorstate.rs:249-261: per block  w0s += block.mapv(norm_sqr).sum_axis(0)
 * (ndarray sum over rows = sequential accumulation per column). */
static double ref_marginal0(const orc_state *s, size_t col, size_t qbit)
{
    size_t C = s->ncols;
    size_t block_size = (size_t)1 << (s->nr_bits - qbit - 1);
    size_t nr_blocks = (size_t)1 << qbit;
    double w0 = 0.0;
    size_t off = 0;
    for (size_t b = 0; b < nr_blocks; b++) {
        double bs = 0.0;
        for (size_t i = off; i < off + block_size; i++) bs += cnorm_sqr(s->states[i * C + col]);
        w0 += bs;
        off += 2 * block_size;
    }
    return w0;
}

int orc_marginal0(const orc_state *s, size_t qbit, int order, double *w0_out)
{
    if (qbit >= s->nr_bits) return ORC_ERR_INVALID_QBIT;
    size_t mask = (size_t)1 << (s->nr_bits - 1 - qbit);
    for (size_t c = 0; c < s->ncols; c++)
        w0_out[c] = order == 0 ? ref_marginal0(s, c, qbit) : canon_total(s, c, mask, 0);
    return ORC_OK;
}

void orc_column_totals(const orc_state *s, int order, double *out)
{
    size_t N = (size_t)1 << s->nr_bits, C = s->ncols;
    for (size_t c = 0; c < C; c++) {
        if (order == 0) {
            double t = 0.0;
            for (size_t i = 0; i < N; i++) t += cnorm_sqr(s->states[i * C + c]);
            out[c] = t;
        } else out[c] = canon_total(s, c, 0, 0);
    }
}

/* ------------------------------------------------------------------ */
/* measurement                                                          */
/* ------------------------------------------------------------------ */

/* vectorstate.rs:91-104 */
static void collapse_col(orc_cplx *ns, size_t C, size_t col, size_t nr_bits, size_t block_size,
                         size_t nr_blocks, size_t offset, double norm_sq)
{
    size_t N = (size_t)1 << nr_bits;
    size_t off = offset;
    for (size_t b = 0; b < nr_blocks; b++) {
        for (size_t i = off; i < off + block_size; i++) { ns[i * C + col].re = 0.0; ns[i * C + col].im = 0.0; }
        off += 2 * block_size;
    }
    orc_cplx f = { 1.0 / sqrt(norm_sq), 0.0 };
    for (size_t i = 0; i < N; i++) ns[i * C + col] = cmul(ns[i * C + col], f);
}

static int measure_common(orc_state *s, size_t qbit, size_t cbit, uint64_t *res, size_t res_len,
                          orc_rng *rng, int order, int collapse)
{
    if (qbit >= s->nr_bits) return ORC_ERR_INVALID_QBIT;
    if (res_len < s->nr_shots) return ORC_ERR_NOT_ENOUGH_SPACE;
    size_t N = (size_t)1 << s->nr_bits, C = s->ncols;
    size_t block_size = (size_t)1 << (s->nr_bits - qbit - 1);
    size_t nr_blocks = (size_t)1 << qbit;
    double *w0s = malloc(C * sizeof(double));
    size_t *n0s = malloc(C * sizeof(size_t));
    orc_marginal0(s, qbit, order, w0s);
    size_t new_nr = 0;
    uint64_t one_mask = (uint64_t)1 << cbit, zero_mask = ~one_mask;
    if (collapse) {
        /* vectorstate.rs:263-275: all Binomial draws first, in column order */
        for (size_t c = 0; c < C; c++) {
            size_t cnt = s->counts[c];
            n0s[c] = (size_t)orc_binomial(rng, cnt, w0s[c] < 1.0 ? w0s[c] : 1.0);
            new_nr += (n0s[c] == 0 || n0s[c] == cnt) ? 1 : 2;
        }
        orc_cplx *ns = calloc(N * new_nr, sizeof(orc_cplx));
        size_t *nc = malloc(new_nr * sizeof(size_t));
        size_t ni = 0, start = 0;
        for (size_t c = 0; c < C; c++) {
            double w0 = w0s[c];
            size_t n0 = n0s[c], cnt = s->counts[c];
            for (size_t j = start; j < start + n0; j++) res[j] &= zero_mask;
            for (size_t j = start + n0; j < start + cnt; j++) res[j] |= one_mask;
            start += cnt;
            for (size_t i = 0; i < N; i++) ns[i * new_nr + ni] = s->states[i * C + c];
            if (n0 == cnt) {
                collapse_col(ns, new_nr, ni, s->nr_bits, block_size, nr_blocks, block_size, w0);
                nc[ni] = cnt;
            } else if (n0 == 0) {
                collapse_col(ns, new_nr, ni, s->nr_bits, block_size, nr_blocks, 0, 1.0 - w0);
                nc[ni] = cnt;
            } else {
                collapse_col(ns, new_nr, ni, s->nr_bits, block_size, nr_blocks, block_size, w0);
                nc[ni] = n0;
                ni++;
                for (size_t i = 0; i < N; i++) ns[i * new_nr + ni] = s->states[i * C + c];
                collapse_col(ns, new_nr, ni, s->nr_bits, block_size, nr_blocks, 0, 1.0 - w0);
                nc[ni] = cnt - n0;
            }
            ni++;
        }
        free(s->states); s->states = ns;
        free(s->counts); s->counts = nc;
        s->ncols = new_nr;
    } else {
        /* vectorstate.rs:375-390 */
        size_t start = 0;
        for (size_t c = 0; c < C; c++) {
            size_t cnt = s->counts[c];
            size_t n0 = (size_t)orc_binomial(rng, cnt, w0s[c] < 1.0 ? w0s[c] : 1.0);
            for (size_t j = start; j < start + n0; j++) res[j] &= zero_mask;
            for (size_t j = start + n0; j < start + cnt; j++) res[j] |= one_mask;
            start += cnt;
        }
    }
    free(w0s); free(n0s);
    return rng->exhausted ? ORC_ERR_RNG_EXHAUSTED : ORC_OK;
}

/* vectorstate.rs:237-329 */
int orc_measure_into(orc_state *s, size_t qbit, size_t cbit, uint64_t *res, size_t res_len,
                     orc_rng *rng, int order)
{
    return measure_common(s, qbit, cbit, res, res_len, rng, order, 1);
}
/* vectorstate.rs:346-393 */
int orc_peek_into(const orc_state *s, size_t qbit, size_t cbit, uint64_t *res, size_t res_len,
                  orc_rng *rng, int order)
{
    return measure_common((orc_state *)s, qbit, cbit, res, res_len, rng, order, 0);
}

static int cmp_double(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* vectorstate.rs:106-161.
 * order 0 (reference order): cumulative weights are a sequential running sum
 *   over amplitudes (rand 0.7 WeightedIndex::new), draws resolved one by one
 *   in draw order; distinct outcomes are emitted in first-occurrence order
 *   (the reference emits them in HashMap iteration order, which is
 *   unspecified -- compare as multisets).
 * order 1 (canonical): cumulative weights in the canonical blocked order;
 *   the draws of a column are sorted, so outcomes come grouped in ascending
 *   basis-index order. */
int orc_measure_all_into(orc_state *s, const size_t *cbits, size_t ncbits, uint64_t *res,
                         size_t res_len, int collapse, orc_rng *rng, int order)
{
    if (res_len < s->nr_shots) return ORC_ERR_NOT_ENOUGH_SPACE;
    if (ncbits != s->nr_bits) return ORC_ERR_INVALID_NR_MEASUREMENT_BITS;
    size_t N = (size_t)1 << s->nr_bits, C = s->ncols;
    size_t cap = s->nr_shots ? s->nr_shots : 1;
    size_t *sc_idx = malloc(cap * sizeof(size_t)), *sc_cnt = malloc(cap * sizeof(size_t));
    size_t nsc = 0;
    for (size_t c = 0; c < C; c++) {
        size_t cnt = s->counts[c];
        size_t first_group = nsc;
        if (order == 0) {
            double *cum = malloc((N > 1 ? N - 1 : 1) * sizeof(double));
            double total = cnorm_sqr(s->states[0 * C + c]);
            for (size_t i = 1; i < N; i++) { cum[i - 1] = total; total += cnorm_sqr(s->states[i * C + c]); }
            uni_f64 u = uniform_new(0.0, total);
            for (size_t j = 0; j < cnt; j++) {
                double chosen = uniform_sample(&u, rng);
                size_t lo = 0, hi = N - 1;      /* first i with cum[i] > chosen, over N-1 entries */
                while (lo < hi) { size_t mid = (lo + hi) / 2; if (cum[mid] <= chosen) lo = mid + 1; else hi = mid; }
                size_t g;
                for (g = first_group; g < nsc; g++) if (sc_idx[g] == lo) break;
                if (g == nsc) { sc_idx[nsc] = lo; sc_cnt[nsc] = 0; nsc++; }
                sc_cnt[g]++;
            }
            free(cum);
        } else {
            size_t nl;
            double *P = canon_leaf_prefix(s, c, 0, 0, &nl);
            size_t leaf = N / nl;
            double total = P[nl - 1];
            uni_f64 u = uniform_new(0.0, total);
            double *chosen = malloc((cnt ? cnt : 1) * sizeof(double));
            for (size_t j = 0; j < cnt; j++) chosen[j] = uniform_sample(&u, rng);
            qsort(chosen, cnt, sizeof(double), cmp_double);
            for (size_t j = 0; j < cnt; j++) {
                double ch = chosen[j];
                size_t lo = 0, hi = nl - 1;     /* number of leaf prefixes (excluding last) <= ch */
                while (lo < hi) { size_t mid = (lo + hi) / 2; if (P[mid] <= ch) lo = mid + 1; else hi = mid; }
                double run = lo == 0 ? 0.0 : P[lo - 1];
                size_t found = (size_t)-1, last_nz = (size_t)-1;
                for (size_t e = 0; e < leaf; e++) {
                    double p = cnorm_sqr(s->states[(lo * leaf + e) * C + c]);
                    if (p > 0.0) last_nz = e;
                    run += p;
                    if (ch < run) { found = e; break; }
                }
                if (found == (size_t)-1) found = last_nz == (size_t)-1 ? leaf - 1 : last_nz;
                size_t idx = lo * leaf + found;
                if (nsc > first_group && sc_idx[nsc - 1] == idx) sc_cnt[nsc - 1]++;
                else { sc_idx[nsc] = idx; sc_cnt[nsc] = 1; nsc++; }
            }
            free(chosen); free(P);
        }
    }
    uint64_t m = 0;
    for (size_t j = 0; j < ncbits; j++) m |= (uint64_t)1 << cbits[j];
    uint64_t mask = ~m;
    size_t off = 0;
    for (size_t g = 0; g < nsc; g++) {
        uint64_t rev = orc_reverse_bits((uint64_t)sc_idx[g], s->nr_bits);
        uint64_t word = orc_shuffle_bits(rev, cbits, ncbits);
        for (size_t j = off; j < off + sc_cnt[g]; j++) res[j] = (res[j] & mask) | word;
        off += sc_cnt[g];
    }
    if (collapse) {
        /* vectorstate.rs:150-158 builds a dense (2^n, n_distinct) matrix: 16 B * 2^n * n_distinct, which the host may refuse */
        orc_cplx *ns = calloc(N * (nsc ? nsc : 1), sizeof(orc_cplx));
        if (!ns) { free(sc_idx); free(sc_cnt); return ORC_ERR_OUT_OF_MEMORY; }
        free(s->states); free(s->counts);
        s->states = ns;
        s->counts = malloc((nsc ? nsc : 1) * sizeof(size_t));
        for (size_t g = 0; g < nsc; g++) { s->states[sc_idx[g] * nsc + g].re = 1.0; s->counts[g] = sc_cnt[g]; }
        s->ncols = nsc;
    }
    free(sc_idx); free(sc_cnt);
    return rng->exhausted ? ORC_ERR_RNG_EXHAUSTED : ORC_OK;
}

/* vectorstate.rs:402-408 */
int orc_reset(orc_state *s, size_t bit, orc_rng *rng, int order, int mode)
{
    uint64_t *m = calloc(s->nr_shots ? s->nr_shots : 1, sizeof(uint64_t));
    int e = orc_measure_into(s, bit, 0, m, s->nr_shots, rng, order);
    if (e) { free(m); return e; }
    uint8_t *control = malloc(s->nr_shots ? s->nr_shots : 1);
    for (size_t j = 0; j < s->nr_shots; j++) control[j] = m[j] != 0;
    const double X[8] = { 0, 0, 1, 0, 1, 0, 0, 0 };
    e = orc_apply_conditional_gate(s, control, s->nr_shots, X, &bit, 1, mode);
    free(m); free(control);
    return e;
}

/* ------------------------------------------------------------------ */
/* gate matrices (src/gates/ *.rs `matrix()`), name table composite.rs:287-445 */
/* ------------------------------------------------------------------ */
static orc_cplx C_(double re, double im) { orc_cplx c = { re, im }; return c; }
static orc_cplx polar(double r, double th) { return C_(r * cos(th), r * sin(th)); }
static orc_cplx cneg(orc_cplx a) { return C_(-a.re, -a.im); }

static int base_matrix(const char *nm, const double *p, size_t np, orc_cplx *m)
{
    const double x = 0.70710678118654752440; /* FRAC_1_SQRT_2 */
    const orc_cplx z = { 0, 0 }, o = { 1, 0 }, i = { 0, 1 };
#define NP(n) do { if (np != (n)) return -2; } while (0)
    if (!strcmp(nm, "h")) { NP(0); m[0] = C_(x, 0); m[1] = C_(x, 0); m[2] = C_(x, 0); m[3] = C_(-x, -0.0); return 1; }      /* hadamard.rs:89-93 */
    if (!strcmp(nm, "i")) { NP(0); m[0] = o; m[1] = z; m[2] = z; m[3] = o; return 1; }                                      /* identity.rs */
    if (!strcmp(nm, "x")) { NP(0); m[0] = z; m[1] = o; m[2] = o; m[3] = z; return 1; }                                      /* x.rs */
    if (!strcmp(nm, "y")) { NP(0); m[0] = z; m[1] = cneg(i); m[2] = i; m[3] = z; return 1; }                                /* y.rs:53-58 */
    if (!strcmp(nm, "z")) { NP(0); m[0] = o; m[1] = z; m[2] = z; m[3] = cneg(o); return 1; }                                /* z.rs */
    if (!strcmp(nm, "s")) { NP(0); m[0] = o; m[1] = z; m[2] = z; m[3] = i; return 1; }                                      /* s.rs:60-66 */
    if (!strcmp(nm, "sdg")) { NP(0); m[0] = o; m[1] = z; m[2] = z; m[3] = cneg(i); return 1; }
    if (!strcmp(nm, "t")) { NP(0); m[0] = o; m[1] = z; m[2] = z; m[3] = C_(x, x); return 1; }                               /* t.rs:52-59: x + x*i */
    if (!strcmp(nm, "tdg")) { NP(0); m[0] = o; m[1] = z; m[2] = z; m[3] = C_(x, -x); return 1; }
    if (!strcmp(nm, "v")) { NP(0); m[0] = C_(.5, .5); m[1] = C_(.5, -.5); m[2] = C_(.5, -.5); m[3] = C_(.5, .5); return 1; } /* v.rs:58-63 */
    if (!strcmp(nm, "vdg")) { NP(0); m[0] = C_(.5, -.5); m[1] = C_(.5, .5); m[2] = C_(.5, .5); m[3] = C_(.5, -.5); return 1; }
    if (!strcmp(nm, "rx")) { NP(1); double h = 0.5 * p[0]; orc_cplx c = C_(cos(h), 0), si = C_(0, sin(h));                  /* rx.rs:64-70 */
        m[0] = c; m[1] = cneg(si); m[2] = cneg(si); m[3] = c; return 1; }
    if (!strcmp(nm, "ry")) { NP(1); double h = 0.5 * p[0]; orc_cplx c = C_(cos(h), 0), s = C_(sin(h), 0);                   /* ry.rs:64-70 */
        m[0] = c; m[1] = cneg(s); m[2] = s; m[3] = c; return 1; }
    if (!strcmp(nm, "rz")) { NP(1); orc_cplx q = polar(1.0, 0.5 * p[0]);                                                    /* rz.rs:65-70 */
        m[0] = C_(q.re, -q.im); m[1] = z; m[2] = z; m[3] = q; return 1; }
    if (!strcmp(nm, "u1")) { NP(1); m[0] = o; m[1] = z; m[2] = z; m[3] = polar(1.0, p[0]); return 1; }                      /* u1.rs:69-75 */
    if (!strcmp(nm, "u2")) { NP(2); double phi = p[0], lam = p[1];                                                          /* u2.rs:70-79 */
        m[0] = C_(x, 0); m[1] = cneg(polar(x, lam)); m[2] = polar(x, phi); m[3] = polar(x, phi + lam); return 1; }
    if (!strcmp(nm, "u3")) { NP(3); double h = 0.5 * p[0], phi = p[1], lam = p[2]; double c = cos(h), s = sin(h);           /* u3.rs:74-84 */
        m[0] = C_(c, 0); m[1] = cneg(polar(s, lam)); m[2] = polar(s, phi); m[3] = polar(c, phi + lam); return 1; }
    if (!strcmp(nm, "swap")) { NP(0); for (int a = 0; a < 16; a++) m[a] = z;                                                /* swap.rs:78-88 */
        m[0] = o; m[1 * 4 + 2] = o; m[2 * 4 + 1] = o; m[15] = o; return 2; }
#undef NP
    return -1;
}

int orc_gate_matrix(const char *name, const double *params, size_t nparams, double *out)
{
    char nm[32];
    size_t L = strlen(name);
    if (L >= sizeof nm) return -1;
    for (size_t a = 0; a <= L; a++) nm[a] = (char)((name[a] >= 'A' && name[a] <= 'Z') ? name[a] + 32 : name[a]);
    static const char *const table[] = { "h","i","s","sdg","t","tdg","v","vdg","x","y","z","rx","ry","rz","u1","u2","u3",
        "cx","cy","cz","ch","cs","csdg","ct","ctdg","cv","cvdg","swap","crx","cry","crz","cu1","cu2","cu3",
        "ccx","ccz","ccrx","ccry","ccrz", NULL };
    int known = 0;
    for (int a = 0; table[a]; a++) if (!strcmp(table[a], nm)) known = 1;
    if (!known) return -1;
    orc_cplx base[16];
    int nb = base_matrix(nm, params, nparams, base);
    int nctl = 0;
    const char *rest = nm;
    /* C<G> = I (+) G, control is the first bit (controlled.rs:60-69) */
    while (nb == -1 && rest[0] == 'c' && rest[1] != '\0') {
        rest++; nctl++;
        nb = base_matrix(rest, params, nparams, base);
    }
    if (nb < 0) return nb;
    if (nctl > 0 && nb != 1) return -1;
    if (nctl > 2) return -1;
    size_t G = (size_t)1 << (nb + nctl), g0 = (size_t)1 << nb;
    orc_cplx *o = (orc_cplx *)out;
    for (size_t a = 0; a < G * G; a++) { o[a].re = 0; o[a].im = 0; }
    for (size_t a = 0; a < G; a++) o[a * G + a].re = 1.0;
    for (size_t r = 0; r < g0; r++)
        for (size_t c = 0; c < g0; c++) o[(G - g0 + r) * G + (G - g0 + c)] = base[r * g0 + c];
    return nb + nctl;
}
