// kernels.cu -- sm_100a kernels of the statevector engine.
//
//  sweep_kernel        fused gate application: one HBM read + one HBM write per
//                      amplitude, gates applied in registers / shared memory
//                      (replaces gates.rs:121-152 + Gate::apply_mat_slice
//                      gates.rs:273-326, one sweep per *group* of gates).
//  generic_gate_kernel unfused k-target dense gate with controls (fallback for
//                      what the sweep kernel cannot express).
//  leaf_totals / block_scan / top_chain / resolve_draws
//                      canonical-order |amp|^2 reductions, prefix scan and
//                      batched weighted sampling (vectorstate.rs:119-133,
//                      249-261).
//  collapse_kernel     zero one half + renormalise, 1 read -> 1 or 2 columns
//                      (vectorstate.rs:91-104, 299-320).
//
// All data is complex128 = double2, HBM-resident, one contiguous buffer of
// 2^n amplitudes per branch column.
#include "kernels.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace q1t {

__constant__ SweepProgram c_prog;

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
// the same with the 32-bit shared-window address already at hand (no generic -> shared conversion per request)
__device__ __forceinline__ void cp_async16_s(unsigned saddr, unsigned long long gaddr)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(gaddr));
}
__device__ __forceinline__ void st_global_cs_b(unsigned long long gaddr, double2 v)
{
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};\n" ::"l"(gaddr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all()
{
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 ld_shared_f64x2(unsigned saddr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void st_global_cs(double2 *p, double2 v)
{
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};\n" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// ---- TMA tile loads (cp.async.bulk.tensor + mbarrier), ladder kernel on dense sweeps ----
struct TmaMaps { CUtensorMap m[kMaxTmaCols]; };     // one tensor map per column of the launch, in the kernel parameters

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
// all threads: wait until the phase with the given parity has completed.  A tile that never arrives (a
// malformed tensor map) traps after ~2 s instead of hanging the device.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done = 0;
    for (unsigned spins = 0; !done; ++spins) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && spins > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void tma_load_5d(unsigned dst, const CUtensorMap *map, unsigned bar, int c4)
{
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %2, %2, %2, %3}], [%4];\n"
                 ::"r"(dst), "l"(map), "r"(0), "r"(c4), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_5d(const CUtensorMap *map, int c4)
{
    asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %1, %1, %1, %2}];\n" ::"l"(map), "r"(0), "r"(c4) : "memory");
}

// ((v & mask) << shift) with shift of either sign (bit-deposit runs of a thread index)
__device__ __forceinline__ unsigned long long run_bits(unsigned v, const BitRun &r)
{
    const unsigned long long m = v & r.mask;
    return r.shift >= 0 ? m << r.shift : m >> -r.shift;
}

// insert a zero bit at position J into p
template <int J>
__device__ __forceinline__ constexpr int ins0(int p)
{
    return ((p >> J) << (J + 1)) | (p & ((1 << J) - 1));
}

// ---------------------------------------------------------------------------
// register-level ops on the 2^R amplitudes of one thread.
// Every body exists twice: a branch-free one for cslot == 0 (no control among
// the register bits, the common case) and one with a per-pair test.
// ---------------------------------------------------------------------------
#define Q1T_FOR_PAIRS(BODY)                                              \
    if (cs == 0u) {                                                      \
        _Pragma("unroll") for (int p = 0; p < kSlots / 2; ++p) {         \
            const int s0 = ins0<J>(p), s1 = s0 | (1 << J);               \
            BODY                                                         \
        }                                                                \
    } else {                                                             \
        _Pragma("unroll") for (int p = 0; p < kSlots / 2; ++p) {         \
            const int s0 = ins0<J>(p), s1 = s0 | (1 << J);               \
            if ((s0 & cs) == cs) { BODY }                                \
        }                                                                \
    }

template <int J>
__device__ __forceinline__ void g1_generic(double2 (&a)[kSlots], const OpDesc &op)
{
    const double m00r = op.m[0], m00i = op.m[1], m01r = op.m[2], m01i = op.m[3];
    const double m10r = op.m[4], m10i = op.m[5], m11r = op.m[6], m11i = op.m[7];
    const unsigned cs = op.cslot;
    Q1T_FOR_PAIRS(
        const double2 x = a[s0]; const double2 y = a[s1];
        double2 u; double2 v;
        u.x = fma(m00r, x.x, fma(-m00i, x.y, fma(m01r, y.x, -(m01i * y.y))));
        u.y = fma(m00r, x.y, fma(m00i, x.x, fma(m01r, y.y, m01i * y.x)));
        v.x = fma(m10r, x.x, fma(-m10i, x.y, fma(m11r, y.x, -(m11i * y.y))));
        v.y = fma(m10r, x.y, fma(m10i, x.x, fma(m11r, y.y, m11i * y.x)));
        a[s0] = u; a[s1] = v;)
}

template <int J>
__device__ __forceinline__ void g1_hadamard(double2 (&a)[kSlots], const OpDesc &op)
{
    const double c = op.m[0];
    const unsigned cs = op.cslot;
    Q1T_FOR_PAIRS(
        const double2 x = a[s0]; const double2 y = a[s1];
        a[s0] = make_double2((x.x + y.x) * c, (x.y + y.y) * c);
        a[s1] = make_double2((x.x - y.x) * c, (x.y - y.y) * c);)
}

// uncontrolled Hadamard without its 1/sqrt(2): the factor is applied once per
// sweep at the store (SweepProgram::scale)
template <int J>
__device__ __forceinline__ void h_unnorm(double2 (&a)[kSlots])
{
#pragma unroll
    for (int p = 0; p < kSlots / 2; ++p) {
        const int s0 = ins0<J>(p), s1 = s0 | (1 << J);
        const double2 x = a[s0], y = a[s1];
        a[s0] = make_double2(x.x + y.x, x.y + y.y);
        a[s1] = make_double2(x.x - y.x, x.y - y.y);
    }
}

template <int J>
__device__ __forceinline__ void g1_antidiag(double2 (&a)[kSlots], const OpDesc &op)
{
    const double2 m01 = make_double2(op.m[2], op.m[3]), m10 = make_double2(op.m[4], op.m[5]);
    const unsigned cs = op.cslot;
    Q1T_FOR_PAIRS(
        const double2 x = a[s0]; const double2 y = a[s1];
        a[s0] = cmul(m01, y);
        a[s1] = cmul(m10, x);)
}

template <int J>
__device__ __forceinline__ void g1_swapx(double2 (&a)[kSlots], const OpDesc &op)
{
    const unsigned cs = op.cslot;
    Q1T_FOR_PAIRS(
        const double2 x = a[s0];
        a[s0] = a[s1];
        a[s1] = x;)
}

template <int J>
__device__ __forceinline__ void g1_diag(double2 (&a)[kSlots], const OpDesc &op)
{
    const double2 m00 = make_double2(op.m[0], op.m[1]), m11 = make_double2(op.m[6], op.m[7]);
    const unsigned cs = op.cslot;
    Q1T_FOR_PAIRS(
        a[s0] = cmul(m00, a[s0]);
        a[s1] = cmul(m11, a[s1]);)
}

// per-slot phase factors of a PHASE op: f[u] = F * prod_{other slot bits set in u} q[i]
__device__ __forceinline__ void phase_factors(double2 (&f)[kSlots / 2], const OpDesc &op, double2 F)
{
    f[0] = F;
#pragma unroll
    for (int i = 0; i < kRegBits - 1; ++i) {
        if (op.flags & (1u << i)) {
            const double2 q = make_double2(op.m[2 * i], op.m[2 * i + 1]);
#pragma unroll
            for (int u = 0; u < (1 << i); ++u) f[u | (1 << i)] = cmul(f[u], q);
        } else {
#pragma unroll
            for (int u = 0; u < (1 << i); ++u) f[u | (1 << i)] = f[u];
        }
    }
}

// PHASE: slots with bit J set are multiplied by f[u]; optionally every slot is
// multiplied by the constant c0 (a global phase term).
template <int J>
__device__ __forceinline__ void phase_apply(double2 (&a)[kSlots], const OpDesc &op, double2 F)
{
    double2 f[kSlots / 2];
    phase_factors(f, op, F);
#pragma unroll
    for (int u = 0; u < kSlots / 2; ++u) {
        const int s1 = ins0<J>(u) | (1 << J);
        a[s1] = cmul(a[s1], f[u]);
    }
    if (op.flags & kFlagC0) {
        const double2 c0 = make_double2(op.m[8], op.m[9]);
#pragma unroll
        for (int u = 0; u < kSlots / 2; ++u) {
            const int s0 = ins0<J>(u);
            a[s0] = cmul(a[s0], c0);
        }
    }
}

// PHASE followed by the unnormalised Hadamard on the same slot bit:
//   (x, y) -> (x + f*y, x - f*y), 8 FP64 instructions per pair
template <int J>
__device__ __forceinline__ void phase_h_apply(double2 (&a)[kSlots], const OpDesc &op, double2 F)
{
    double2 f[kSlots / 2];
    phase_factors(f, op, F);
#pragma unroll
    for (int u = 0; u < kSlots / 2; ++u) {
        const int s0 = ins0<J>(u), s1 = s0 | (1 << J);
        const double2 x = a[s0], y = a[s1];
        const double tr = fma(y.x, f[u].x, -(y.y * f[u].y));
        const double ti = fma(y.x, f[u].y, y.y * f[u].x);
        a[s0] = make_double2(x.x + tr, x.y + ti);
        a[s1] = make_double2(x.x - tr, x.y - ti);
    }
}

// ---------------------------------------------------------------------------
// ROUND_PH: NS fused (phase, Hadamard-butterfly) steps on the top NS slot bits
// kRegBits-NS..kRegBits-1 (lower slots are padding that only acts as partner bits),
// straight-line code.  Step j multiplies the bit-j=1 half by
//   F_j(thread, tile) * prod_{i<j, slot bit i set} q_{j,i}
// and then applies the unnormalised butterfly -- a radix-2^NS decimation stage
// of the QFT/FFT with per-thread twiddles, but for arbitrary phase coefficients.
// ---------------------------------------------------------------------------
// where a round's amplitudes come from and go to
struct RoundIO {
    char *tile_b;                 // shared-memory tile
    unsigned swT;                 // swizzled byte offset of this thread
    const double2 *gsrc;          // != null: read from global (direct load), slot offsets P.dl_slot
    double2 *gdst;                // != null: write to global (direct store), slot offsets P.ds_slot
    bool zero_fill;               // generated input on the direct path
    unsigned gen_slot;            // slot that holds the basis element (0xffffffff: none)
    double scale;
};

__device__ __forceinline__ void round_load(double2 (&a)[kSlots], const RoundIO &io, const RoundDesc &R, const SweepProgram &P)
{
    if (io.gsrc) {
#pragma unroll
        for (int s = 0; s < kSlots; ++s) a[s] = __ldcs(io.gsrc + P.dl_slot[s]);
    } else if (io.zero_fill) {
#pragma unroll
        for (int s = 0; s < kSlots; ++s) a[s] = make_double2(s == (int)io.gen_slot ? P.gen_scale : 0.0, 0.0);
    } else {
#pragma unroll
        for (int s = 0; s < kSlots; ++s) a[s] = *reinterpret_cast<const double2 *>(io.tile_b + (io.swT ^ R.sw_slot[s]));
    }
}

__device__ __forceinline__ void round_store(const double2 (&a)[kSlots], const RoundIO &io, const RoundDesc &R, const SweepProgram &P)
{
    if (io.gdst) {
#pragma unroll
        for (int s = 0; s < kSlots; ++s) st_global_cs(io.gdst + P.ds_slot[s], make_double2(a[s].x * io.scale, a[s].y * io.scale));
    } else {
#pragma unroll
        for (int s = 0; s < kSlots; ++s) *reinterpret_cast<double2 *>(io.tile_b + (io.swT ^ R.sw_slot[s])) = a[s];
    }
}

// Step J of a ladder round: the pairs of slot bit J whose already-processed slot bits (< J) read
// u get the factor f(u) = F * prod_{i in u} q_i.  The factors are produced depth-first (at most
// J + 1 of them alive) instead of as a 2^J-entry table, which keeps 32 amplitudes + factors
// within the register budget of 3 CTAs per SM.
template <int J, int I>
struct LadderStep {
    static __device__ __forceinline__ void run(double2 (&a)[kSlots], const double *qm, double2 f, int u)
    {
        LadderStep<J, I - 1>::run(a, qm, f, u);
        const double2 q = make_double2(qm[2 * I], qm[2 * I + 1]);
        LadderStep<J, I - 1>::run(a, qm, cmul(f, q), u | (1 << I));
    }
};
template <int J>
struct LadderStep<J, -1> {
    static __device__ __forceinline__ void run(double2 (&a)[kSlots], const double *, double2 f, int u)
    {
#pragma unroll
        for (int h = 0; h < (kSlots >> (J + 1)); ++h) {
            const int s0 = (h << (J + 1)) | u, s1 = s0 | (1 << J);
            const double2 x = a[s0], y = a[s1];
            // (x + f y, x - f y) in six FMAs instead of 2 mul + 2 fma + 4 add: the sum by two chained FMAs per
            // component, the difference as 2 x - sum (one FMA).  The FP64 pipe is the busiest unit of the kernel.
            const double pr = fma(y.x, f.x, fma(-y.y, f.y, x.x));
            const double pi = fma(y.x, f.y, fma(y.y, f.x, x.y));
            a[s0] = make_double2(pr, pi);
            a[s1] = make_double2(fma(2.0, x.x, -pr), fma(2.0, x.y, -pi));
        }
    }
};

template <int NS, int J, bool LO_SHARED = false>
struct LadderSteps {
    static __device__ __forceinline__ void run(double2 (&a)[kSlots], const RoundDesc &R, const SweepProgram &P,
                                               const PhaseTab *__restrict__ ptabs, const double2 *s_hiF, int he_bits, unsigned il, unsigned ih,
                                               const double2 *s_lo = nullptr)
    {
        const OpDesc &op = P.ops[R.op_begin + J - (kRegBits - NS)];
        const double2 lo = LO_SHARED ? s_lo[(op.phase_id << kThrLoBits) + il]
                                     : __ldg(reinterpret_cast<const double2 *>(ptabs[op.phase_id].lo) + il);
        const double2 hi = s_hiF[(op.phase_id << he_bits) + ih];          // hi[ih] * tile factor
        LadderStep<J, J - 1>::run(a, op.m, cmul(lo, hi), 0);
        LadderSteps<NS, J + 1, LO_SHARED>::run(a, R, P, ptabs, s_hiF, he_bits, il, ih, s_lo);
    }
};
template <int NS, bool LO_SHARED>
struct LadderSteps<NS, kRegBits, LO_SHARED> {
    static __device__ __forceinline__ void run(double2 (&)[kSlots], const RoundDesc &, const SweepProgram &,
                                               const PhaseTab *__restrict__, const double2 *, int, unsigned, unsigned,
                                               const double2 * = nullptr) {}
};

// One body for every number of steps: step J runs when J >= first = kRegBits - nsteps (a uniform branch per step).
// The five straight-line instantiations of LadderSteps<NS, ..> are 34 KB of code of which a sweep with rounds of
// different lengths executes two or three; the instruction caches (32 KB L1.5) hold one body (profiles/r2_ladder_lean.md).
template <int J>
struct LadderU {
    static __device__ __forceinline__ void run(double2 (&a)[kSlots], const RoundDesc &R, const SweepProgram &P, const double2 *s_hiF,
                                               int he_bits, unsigned il, unsigned ih, const double2 *s_lo, int first)
    {
        if (J >= first) {
            const OpDesc &op = P.ops[R.op_begin + J - first];
            const double2 lo = s_lo[(op.phase_id << kThrLoBits) + il];
            const double2 hi = s_hiF[(op.phase_id << he_bits) + ih];          // hi[ih] * tile factor
            LadderStep<J, J - 1>::run(a, op.m, cmul(lo, hi), 0);
        }
        LadderU<J + 1>::run(a, R, P, s_hiF, he_bits, il, ih, s_lo, first);
    }
};
template <>
struct LadderU<kRegBits> {
    static __device__ __forceinline__ void run(double2 (&)[kSlots], const RoundDesc &, const SweepProgram &, const double2 *, int, unsigned,
                                               unsigned, const double2 *, int) {}
};

template <int NS>
__device__ __forceinline__ void round_ph(const RoundIO &io, const RoundDesc &R, const SweepProgram &P,
                                         const PhaseTab *__restrict__ ptabs, const double2 *s_hiF, int he_bits, unsigned tid)
{
    double2 a[kSlots];
    const unsigned il = tid & ((1u << kThrLoBits) - 1u), ih = tid >> kThrLoBits;
    round_load(a, io, R, P);
    LadderSteps<NS, kRegBits - NS>::run(a, R, P, ptabs, s_hiF, he_bits, il, ih);
    round_store(a, io, R, P);
}

// LINPHASE: product of independent one-qubit phases on any set of index bits (the linear part
// of the phase polynomial): thread/tile bits through F, register bits through m[0..3]
__device__ __forceinline__ void linphase_apply(double2 (&a)[kSlots], const OpDesc &op, double2 F)
{
    double2 f[kSlots];
    f[0] = F;
#pragma unroll
    for (int i = 0; i < kRegBits; ++i) {
        const double2 q = make_double2(op.m[2 * i], op.m[2 * i + 1]);
#pragma unroll
        for (int u = 0; u < (1 << i); ++u) f[u | (1 << i)] = cmul(f[u], q);
    }
#pragma unroll
    for (int s = 0; s < kSlots; ++s) a[s] = cmul(a[s], f[s]);
}

#if Q1T_REG_BITS >= 5
#define Q1T_DISPATCH_J(fn, ...)          \
    switch (op.j) {                      \
    case 0: fn<0>(__VA_ARGS__); break;   \
    case 1: fn<1>(__VA_ARGS__); break;   \
    case 2: fn<2>(__VA_ARGS__); break;   \
    case 3: fn<3>(__VA_ARGS__); break;   \
    default: fn<4>(__VA_ARGS__); break;  \
    }
#else
#define Q1T_DISPATCH_J(fn, ...)          \
    switch (op.j) {                      \
    case 0: fn<0>(__VA_ARGS__); break;   \
    case 1: fn<1>(__VA_ARGS__); break;   \
    case 2: fn<2>(__VA_ARGS__); break;   \
    default: fn<3>(__VA_ARGS__); break;  \
    }
#endif

__device__ __forceinline__ unsigned long long outer_base(const uint64_t (&tab)[kOuterChunks][1 << kOuterChunkBits],
                                                         unsigned long long o, int n_outer)
{
    unsigned long long r = tab[0][o & 63ull];
    if (n_outer > 6) r |= tab[1][(o >> 6) & 63ull];
    if (n_outer > 12) r |= tab[2][(o >> 12) & 63ull];
    if (n_outer > 18) r |= tab[3][(o >> 18) & 63ull];
    if (n_outer > 24) r |= tab[4][(o >> 24) & 63ull];
    return r;
}

// ---------------------------------------------------------------------------
// the sweep kernel
// ---------------------------------------------------------------------------
// INTERP = false: programs whose rounds are all ROUND_PH ladders (or that have no rounds at all):
// the op interpreter is compiled out, which leaves a small straight-line kernel.
// GPROG: the program is read from global memory (gprog) instead of the constant bank.  A batch replayed as a CUDA
// graph keeps its programs device-resident, so a replay consists of kernel nodes only -- no 33 KB upload per sweep.
template <bool INTERP, int MAXT, int MINB, bool GPROG>
__global__ void __launch_bounds__(MAXT, MINB)
sweep_kernel(const double2 *const *__restrict__ src_cols, double2 *const *__restrict__ dst_cols,
             const PhaseTab *__restrict__ ptabs, const unsigned long long *__restrict__ gen_idx, const SweepProgram *__restrict__ gprog)
{
    extern __shared__ double2 tile[];

    const SweepProgram &P = *(GPROG ? gprog : &c_prog);
    const int T = P.T, TB = P.TB;
    const unsigned tid = threadIdx.x;
    const unsigned long long o = blockIdx.x;
    const int col = blockIdx.y;
    char *const tile_b = reinterpret_cast<char *>(tile);
    // per-tile phase factors behind the tile: s_hiF[pid][ih] = hi[ih] * exp(i*pi*angle(outer bits))
    double2 *const s_hiF = tile + (1u << T);

    const bool dload = P.direct_load != 0, dstore = P.direct_store != 0 && P.nrounds > 0;
    const unsigned long long obase_src = outer_base(P.o_src, o, P.n_outer);

    // ---- bring the tile in, unless round 0 reads global memory itself ----
    if (!dload) {
        const unsigned sw_tid = tile_swizzle(tid) * 16u;
        if (!P.generate) {
            // coalesced: consecutive tid -> consecutive source addresses
            unsigned long long soff = obase_src;
            for (int k = 0; k < P.ld_nruns; ++k) soff |= (unsigned long long)(tid & P.ld_runs[k].mask) << P.ld_runs[k].shift;
            const double2 *__restrict__ src = src_cols[col] + soff;
#pragma unroll
            for (int i = 0; i < kSlots; ++i) cp_async16(tile_b + (sw_tid ^ P.ld_sw_hi[i]), src + P.ld_hi[i]);
        } else {
#pragma unroll
            for (int i = 0; i < kSlots; ++i) *reinterpret_cast<double2 *>(tile_b + (sw_tid ^ P.ld_sw_hi[i])) = make_double2(0.0, 0.0);
        }
    }
    // per-tile phase factors: the part that depends on the outer index bits only, folded into the
    // table of the high thread bits (one complex multiply per thread and step less)
    const int he_bits = TB > kThrLoBits ? TB - kThrLoBits : 0;
    {
        for (int e = tid; e < (P.nphase << he_bits); e += blockDim.x) {
            const int pid = e >> he_bits, ih = e & ((1 << he_bits) - 1);
            const PhaseTab &pt = ptabs[pid];
            double ang = pt.base;
            for (int i = 0; i < P.n_outer; ++i)
                if ((o >> i) & 1ull) ang += pt.outer_coef[i];
            double sn, cs;
            sincospi(ang, &sn, &cs);
            const double2 hi = __ldg(reinterpret_cast<const double2 *>(pt.hi) + ih);
            s_hiF[e] = cmul(hi, make_double2(cs, sn));
        }
    }
    // the basis element of a generated input (it lives in exactly one tile of the grid)
    unsigned gen_l = 0xffffffffu;
    if (P.generate) {
        const unsigned long long g = gen_idx[col];
        if ((g & ~P.tile_mask_src) == obase_src) {
            gen_l = 0;
            for (int i = 0; i < T; ++i) gen_l |= (unsigned)((g >> P.tsrc[i]) & 1ull) << i;
        }
    }
    if (!dload) {
        if (!P.generate) cp_async_commit_wait_all();
        else {
            __syncthreads();
            if (tid == 0 && gen_l != 0xffffffffu) tile[tile_swizzle(gen_l)] = make_double2(P.gen_scale, 0.0);
        }
    }
    __syncthreads();

    const unsigned long long vhi = o << T;
    for (int r = 0; r < P.nrounds; ++r) {
        const RoundDesc &R = P.rounds[r];
        unsigned thrL = 0;
        for (int k = 0; k < R.nruns; ++k) thrL |= (unsigned)run_bits(tid, R.runs[k]);
        RoundIO io;
        io.tile_b = tile_b;
        io.swT = tile_swizzle(thrL) * 16u;
        io.gsrc = nullptr;
        io.gdst = nullptr;
        io.zero_fill = false;
        io.gen_slot = 0xffffffffu;
        io.scale = P.scale;
        if (r == 0 && dload) {
            if (!P.generate) {
                unsigned long long off = obase_src;
                for (int k = 0; k < P.dl_nruns; ++k) off |= (unsigned long long)(tid & P.dl_runs[k].mask) << P.dl_runs[k].shift;
                io.gsrc = src_cols[col] + off;
            } else {
                io.zero_fill = true;
                if (gen_l != 0xffffffffu) {
                    unsigned regmask = 0, slot = 0;
                    for (int j = 0; j < kRegBits; ++j) {
                        regmask |= 1u << R.reg_tb[j];
                        slot |= ((gen_l >> R.reg_tb[j]) & 1u) << j;
                    }
                    if ((gen_l & ~regmask) == thrL) io.gen_slot = slot;
                }
            }
        } else if (r > 0) {
            if (R.sync_before == 2) __syncthreads();
            else __syncwarp();
        }
        if (r + 1 == P.nrounds && dstore) {
            unsigned long long off = outer_base(P.o_dst, o, P.n_outer);
            for (int k = 0; k < P.ds_nruns; ++k) off |= run_bits(tid, P.ds_runs[k]);
            io.gdst = dst_cols[col] + off;
        }

        if (!INTERP || R.kind == ROUND_PH) {
            switch (R.nsteps) {
            case 1: round_ph<1>(io, R, P, ptabs, s_hiF, he_bits, tid); break;
            case 2: round_ph<2>(io, R, P, ptabs, s_hiF, he_bits, tid); break;
            case 3: round_ph<3>(io, R, P, ptabs, s_hiF, he_bits, tid); break;
#if Q1T_REG_BITS >= 5
            case 4: round_ph<4>(io, R, P, ptabs, s_hiF, he_bits, tid); break;
            default: round_ph<5>(io, R, P, ptabs, s_hiF, he_bits, tid); break;
#else
            default: round_ph<4>(io, R, P, ptabs, s_hiF, he_bits, tid); break;
#endif
            }
            continue;
        }
        if (!INTERP) continue;
        double2 a[kSlots];
        round_load(a, io, R, P);
        const unsigned long long vbase = vhi | thrL;
        for (int k = R.op_begin; k < R.op_end; ++k) {
            const OpDesc &op = P.ops[k];
            const unsigned kind = op.kind;
            if (kind == OP_PHASE_H || kind == OP_PHASE || kind == OP_LINPHASE) {
                const PhaseTab &pt = ptabs[op.phase_id];
                const unsigned il = tid & ((1u << kThrLoBits) - 1u), ih = tid >> kThrLoBits;
                const double2 lo = __ldg(reinterpret_cast<const double2 *>(pt.lo) + il);
                const double2 F = cmul(lo, s_hiF[(op.phase_id << he_bits) + ih]);
                if (kind == OP_PHASE_H) { Q1T_DISPATCH_J(phase_h_apply, a, op, F) }
                else if (kind == OP_LINPHASE) linphase_apply(a, op, F);
                else { Q1T_DISPATCH_J(phase_apply, a, op, F) }
            } else if (kind == OP_H_UNNORM) {
                Q1T_DISPATCH_J(h_unnorm, a)
            } else {
                if ((vbase & op.cmask) != op.cmask) continue;
                switch (kind) {
                case OP_G1_HADAMARD: Q1T_DISPATCH_J(g1_hadamard, a, op) break;
                case OP_G1_ANTIDIAG: Q1T_DISPATCH_J(g1_antidiag, a, op) break;
                case OP_G1_SWAPX: Q1T_DISPATCH_J(g1_swapx, a, op) break;
                case OP_G1_DIAG: Q1T_DISPATCH_J(g1_diag, a, op) break;
                default: Q1T_DISPATCH_J(g1_generic, a, op) break;
                }
            }
        }
        round_store(a, io, R, P);
    }
    if (dstore) return;
    const double scale = P.scale;
    __syncthreads();

    // ---- store the tile (coalesced in the destination layout) ----
    unsigned long long doff = outer_base(P.o_dst, o, P.n_outer);
    unsigned l_lo = 0;
    for (int k = 0; k < P.st_nruns; ++k) {
        doff |= (unsigned long long)(tid & P.st_runs[k].mask) << P.st_runs[k].shift;
        const int sh = P.st_lruns[k].shift;
        const unsigned v = tid & P.st_lruns[k].mask;
        l_lo |= sh >= 0 ? v << sh : v >> -sh;
    }
    double2 *__restrict__ dst = dst_cols[col] + doff;
    const unsigned sw_lo = tile_swizzle(l_lo) * 16u;
    if (scale == 1.0) {
#pragma unroll
        for (int i = 0; i < kSlots; ++i)
            st_global_cs(dst + P.st_off_hi[i], *reinterpret_cast<const double2 *>(tile_b + (sw_lo ^ P.st_l_hi[i])));
    } else {
#pragma unroll
        for (int i = 0; i < kSlots; ++i) {
            const double2 v = *reinterpret_cast<const double2 *>(tile_b + (sw_lo ^ P.st_l_hi[i]));
            st_global_cs(dst + P.st_off_hi[i], make_double2(v.x * scale, v.y * scale));
        }
    }
}

// ---------------------------------------------------------------------------
// Broadcast sweeps (ladder kernel, support tracking).  A ladder step computes (x, y) -> (x + f y, x - f y) on
// its slot bit.  When that bit is still pinned to 0 the y half is zero: the phase drops out and the step copies
// x into the other half.  If that holds for every step of every round of a sweep (first gates on fresh |0>
// qubits: a QFT or an H layer on a basis state whose target bits are 0), the whole sweep is
//     out[l] = in[l with the target bits cleared]
// -- no rounds, no phase tables: the staged store pass reads the few live inputs straight from the tile.
// The live inputs of a tile have all target bits clear, so the tile positions with target bits set are free:
// two target bits select one of four slots, and the inputs of the next three tiles are in flight while a
// tile is stored (under a saturating write stream a read takes several tile times to come back).
// Separate no-inline function: keeps the register allocation of the dense path as it was.
// ---------------------------------------------------------------------------
__device__ __noinline__ void ladder_broadcast_tiles(const double2 *__restrict__ src, double2 *__restrict__ dst,
                                                    double *__restrict__ leaf_col, char *tile_b, const unsigned long long *s_off,
                                                    unsigned bc_keep, unsigned sup_mt, unsigned sup_vt, unsigned long long sup_g,
                                                    unsigned long long sup_mo)
{
    const SweepProgram &P = c_prog;
    const int T = P.T, TB = P.TB;
    const unsigned tid = threadIdx.x;
    const unsigned long long ntiles = 1ull << P.n_outer;
    const unsigned long long soff_t = s_off[2 * tid], doff_t = s_off[2 * tid + 1];
    const double scale = P.scale;
    const bool leaf_fuse = leaf_col != nullptr;
    const bool zero_fill = P.sup_mode == 2;
    const unsigned sw_tid = tile_swizzle(tid) * 16u;
    const unsigned tile_s = static_cast<unsigned>(__cvta_generic_to_shared(tile_b));
    const unsigned tgt = ~bc_keep & ((1u << T) - 1u);
    const unsigned tb_a = __ffs(tgt) - 1, tb_b = __ffs(tgt & (tgt - 1u)) - 1;
    auto slot_pat = [&](unsigned k) { return ((k & 1u) << tb_a) | (((k >> 1) & 1u) << tb_b); };
    // Tiles are walked in DESTINATION order (w_src / w_dst): the CTAs of the grid then write a few sequential
    // streams instead of chunks scattered by the relabelling.
    // first tile at or after t (stride: the grid) that holds data; tiles passed on the way are zero-filled when
    // the column has to leave the batch dense
    auto next_from = [&](unsigned long long t) {
        for (; t < ntiles; t += gridDim.x) {
            const unsigned long long ob = outer_base(P.w_src, t, P.n_outer);
            if (!sup_mo || ((ob ^ sup_g) & sup_mo) == 0ull) break;
            if (zero_fill) {
                const unsigned long long db = outer_base(P.w_dst, t, P.n_outer) | doff_t;
                const double2 z = make_double2(0.0, 0.0);
#pragma unroll
                for (int i = 0; i < kSlots; ++i) st_global_cs(dst + db + P.st_off_hi[i], z);
                if (leaf_fuse && (tid & 31) == 0) leaf_col[db >> 10] = 0.0;
            }
        }
        return t;
    };
    // which of its 32 load slots a thread has to fetch: the elements inside the support (for most threads: none)
    unsigned ldmask = 0;
#pragma unroll
    for (int i = 0; i < kSlots; ++i) {
        const unsigned e = tid | ((unsigned)i << TB);
        ldmask |= (((e ^ sup_vt) & sup_mt) == 0u ? 1u : 0u) << i;
    }
    auto issue_slot = [&](unsigned long long t, unsigned k) {
        if (ldmask != 0u && t < ntiles) {
            const double2 *__restrict__ p = src + (outer_base(P.w_src, t, P.n_outer) | soff_t);
            const unsigned sw = sw_tid ^ (tile_swizzle(slot_pat(k)) * 16u);
            for (unsigned m = ldmask; m != 0u; m &= m - 1u) {
                const int i = __ffs(m) - 1;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(tile_s + (sw ^ P.ld_sw_hi[i])), "l"(p + P.ld_hi[i]));
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);     // one group per tile, empty ones included
    };
    unsigned l_lo = 0;
    for (int k = 0; k < P.st_nruns; ++k) {
        const int sh = P.st_lruns[k].shift;
        const unsigned t = tid & P.st_lruns[k].mask;
        l_lo |= sh >= 0 ? t << sh : t >> -sh;
    }
    // per thread and slot, once: where the source element sits in the tile (slot pattern not yet applied) and
    // whether it can be non-zero at all
    const unsigned chk = sup_mt & bc_keep;
    unsigned so[kSlots];
    unsigned valid = 0;
#pragma unroll
    for (int i = 0; i < kSlots; ++i) {
        const unsigned l = l_lo | tile_swizzle(P.st_l_hi[i] >> 4);     // the swizzle is an involution
        so[i] = tile_swizzle(l & bc_keep) * 16u;
        valid |= (((l ^ sup_vt) & chk) == 0u ? 1u : 0u) << i;
    }
    unsigned long long o = next_from(blockIdx.x);
    unsigned long long o1 = o < ntiles ? next_from(o + gridDim.x) : o, o2 = o1 < ntiles ? next_from(o1 + gridDim.x) : o1,
                       o3 = o2 < ntiles ? next_from(o2 + gridDim.x) : o2;
    issue_slot(o, 0);
    issue_slot(o1, 1);
    issue_slot(o2, 2);
    for (unsigned k = 0; o < ntiles; ++k) {
        asm volatile("cp.async.wait_group 2;\n" ::: "memory");
        __syncthreads();                   // the inputs of tile o are visible; every thread is done with slot (k + 3) & 3
        issue_slot(o3, k + 3);
        const unsigned patsw = tile_swizzle(slot_pat(k)) * 16u;           // the swizzle is linear over xor
        const unsigned long long db = outer_base(P.w_dst, o, P.n_outer) | doff_t;
        double2 *__restrict__ q = dst + db;
        double leaf_acc = 0.0;
#pragma unroll
        for (int i = 0; i < kSlots; ++i) {
            // element l of the tile = the input element with the target bits cleared, or zero where a
            // pinned bit outside the targets differs from the basis index
            double2 x = make_double2(0.0, 0.0);
            if ((valid >> i) & 1u) x = ld_shared_f64x2(tile_s + (so[i] ^ patsw));
            if (scale != 1.0) x = make_double2(x.x * scale, x.y * scale);
            if (leaf_fuse) leaf_acc = __dadd_rn(leaf_acc, __dadd_rn(__dmul_rn(x.x, x.x), __dmul_rn(x.y, x.y)));
            st_global_cs(q + P.st_off_hi[i], x);
        }
        if (leaf_fuse) {
            // lane l has summed elements l, l+32, ... of its warp's leaf in increasing order: the canonical butterfly
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) leaf_acc = __dadd_rn(leaf_acc, __shfl_xor_sync(0xffffffffu, leaf_acc, off));
            if ((tid & 31) == 0) leaf_col[db >> 10] = leaf_acc;
        }
        o = o1; o1 = o2; o2 = o3;
        if (o3 < ntiles) o3 = next_from(o3 + gridDim.x);
    }
}

// ---------------------------------------------------------------------------
// the ladder kernel: sweeps whose rounds are all ROUND_PH ladders (QFT-like circuits).
//
// Persistent CTAs (a few per SM) walk over the tiles.  A CTA is a serial chain
// per tile (tables -> loads -> rounds -> stores) and only 2-3 CTAs fit on an SM,
// so the chain is software-pipelined: as soon as every thread of the CTA holds
// the amplitudes of the tile's LAST round in registers the shared-memory tile is
// dead, and the cp.async loads and phase tables of the NEXT tile are issued into
// it before the last round's arithmetic and global stores.
// ---------------------------------------------------------------------------
#ifdef Q1T_PHASE_CLOCKS
// debug build: cycles per phase of the tile loop, per warp of one CTA, printed at exit
#define PCLK_DECL long long pclk[16] = { 0 }; long long pclk_t = clock64();
#define PCLK(i) do { const long long now__ = clock64(); pclk[i] += now__ - pclk_t; pclk_t = now__; } while (0)
#else
#define PCLK_DECL
#define PCLK(i) do { } while (0)
#endif

// Compile-time switches of the lean dense path (A/B builds: python -m q1tsim_b200.build --variant=_x -DQ1T_LADDER_UNIFIED=0 ..)
#ifndef Q1T_LADDER_UNIFIED
#define Q1T_LADDER_UNIFIED 0      // (measured: 7.79 vs 7.23 ms per dense sweep -- slower, off) one ladder body with guarded steps instead of one straight-line body per round length
#endif
#ifndef Q1T_LADDER_STATIC_ROUNDS
#define Q1T_LADDER_STATIC_ROUNDS 0
#endif
#ifndef Q1T_LADDER_BYTE_ADDR
#define Q1T_LADDER_BYTE_ADDR 1    // byte-offset tables + 32-bit shared addresses in the load issue and the direct store
#endif
// DENSE: instantiation for sweeps over dense columns (no input generation, no support tracking, no broadcast): those
// paths and their per-round predicates are compiled out (TMA implies it)
// REMOTE: the tile loads are gathered from the peers' shards (RemoteGather, kernels.h): the fused remap read
template <int MAXT, int MINB, bool TMA, bool DENSE, bool REMOTE>
__global__ void __launch_bounds__(MAXT, MINB)
ladder_kernel(const double2 *const *__restrict__ src_cols, double2 *const *__restrict__ dst_cols,
              const PhaseTab *__restrict__ ptabs, const unsigned long long *__restrict__ gen_idx, double *__restrict__ leaf_out,
              const __grid_constant__ TmaMaps tmaps, const RemoteGather rg)
{
    extern __shared__ __align__(1024) double2 tile[];
    const SweepProgram &P = c_prog;
    const int T = P.T, TB = P.TB;
    const unsigned tid = threadIdx.x;
    const int col = blockIdx.y;
    const unsigned long long ntiles = 1ull << P.n_outer;
    char *const tile_b = reinterpret_cast<char *>(tile);
    const int he_bits = TB > kThrLoBits ? TB - kThrLoBits : 0;
    const int ntab = P.nphase << he_bits;
    // tile, then 16 bytes for the mbarrier of the TMA mode, then the tables
    double2 *const s_hiF = tile + (1u << T) + 1;         // [2][ntab]: hi[ih] * exp(i*pi*angle(outer bits)), double-buffered
    double2 *const s_tileF = s_hiF + 2 * ntab;           // [nphase]
    double2 *const s_lo = s_tileF + P.nphase;            // [nphase][16]  copies of the PhaseTab rows (the CTA is persistent)
    double *const s_coef = reinterpret_cast<double *>(s_lo + (P.nphase << kThrLoBits));   // [nphase][n_outer + 1], last = base
    const int ncoef = P.n_outer + 1;
    // per-thread source / destination offsets: kept in shared memory, not in registers -- with 32
    // amplitudes per thread the compiler spilled them, and a local-memory reload is an L2 round trip
    unsigned long long *const s_off = reinterpret_cast<unsigned long long *>(s_coef + ((P.nphase * ncoef + 1) & ~1));
    unsigned long long *const s_peer = s_off + 2 * blockDim.x;     // REMOTE: [P] base address of every rank's current shard buffer
    // TMA mode (dense sweeps): the tile arrives by one cp.async.bulk.tensor request issued by thread 0 and
    // lives in shared memory in TMA order (SweepProgram::tma_*); completion is signalled on an mbarrier
    // (a separate instantiation: dense sweeps only, so input generation, support tracking and broadcast
    // sweeps are compiled out of it)
    constexpr bool tma = TMA;
    constexpr bool dense = TMA || DENSE;
#define Q1T_TILE_S0 static_cast<unsigned>(__cvta_generic_to_shared(tile))
#define Q1T_S_BAR (Q1T_TILE_S0 + (16u << T))
    const bool generate = !dense && P.generate != 0;
    const bool staged_store = P.direct_store == 0;
    const double scale = P.scale;
    // fused canonical leaf totals (one leaf of 1024 amplitudes per warp in the staged store pass)
    const bool leaf_fuse = P.leaf_fuse != 0 && leaf_out != nullptr && staged_store;
    double *__restrict__ const leaf_col = leaf_fuse ? leaf_out + ((unsigned long long)col << (P.n - 10)) : nullptr;
    const double2 *__restrict__ const src = src_cols[col];
    double2 *__restrict__ const dst = dst_cols[col];

    // thread-constant address parts
    const unsigned sw_tid = tile_swizzle(tid) * 16u;
    unsigned long long soff_t = 0;
    for (int k = 0; k < P.ld_nruns; ++k) soff_t |= (unsigned long long)(tid & P.ld_runs[k].mask) << P.ld_runs[k].shift;
    unsigned long long doff_t = 0;
    unsigned sw_lo = 0;
    if (staged_store) {
        unsigned l_lo = 0;
        for (int k = 0; k < P.st_nruns; ++k) {
            doff_t |= (unsigned long long)(tid & P.st_runs[k].mask) << P.st_runs[k].shift;
            const int sh = P.st_lruns[k].shift;
            const unsigned v = tid & P.st_lruns[k].mask;
            l_lo |= sh >= 0 ? v << sh : v >> -sh;
        }
        sw_lo = (tma ? tma_swizzle(l_lo) : tile_swizzle(l_lo)) * 16u;
    } else {
        for (int k = 0; k < P.ds_nruns; ++k) doff_t |= run_bits(tid, P.ds_runs[k]);
    }
    if (tma) {
        if (tid == 0) {
            if (Q1T_TILE_S0 & 1023u) __trap();     // SWIZZLE_128B needs the 1024-byte alignment
            mbar_init(Q1T_S_BAR, 1);
        }
        __syncthreads();
    }
    // support tracking: amplitudes that differ from the basis index in a bit of sup_mask are zero
    const int sup_mode = dense ? 0 : P.sup_mode;
    const unsigned long long sup_g = (generate || sup_mode) ? gen_idx[col] : 0ull;
    const unsigned long long gen_g = sup_g;
    const unsigned long long sup_mo = sup_mode ? (P.sup_mask & ~P.tile_mask_src) : 0ull;   // pinned outer bits
    unsigned sup_mt = 0, sup_vt = 0;                                                     // pinned tile bits (tile-local)
    if (sup_mode) {
        for (int i = 0; i < T; ++i) {
            sup_mt |= (unsigned)((P.sup_mask >> P.tsrc[i]) & 1ull) << i;
            sup_vt |= (unsigned)((sup_g >> P.tsrc[i]) & 1ull) << i;
        }
        sup_vt &= sup_mt;
    }
    s_off[2 * tid] = soff_t;            // only ever read back by the same thread
    s_off[2 * tid + 1] = doff_t;
    unsigned long long rg_lpmask = 0, rg_mybits = 0;
    unsigned rg_rank0 = 0;
    if (REMOTE) {
        unsigned gbmask = 0;
        for (int j = 0; j < rg.k; ++j) {
            rg_lpmask |= 1ull << rg.lp[j];
            rg_mybits |= (unsigned long long)((rg.rank >> rg.gb[j]) & 1) << rg.lp[j];
            gbmask |= 1u << rg.gb[j];
        }
        rg_rank0 = (unsigned)rg.rank & ~gbmask;
        for (int r = tid; r < rg.P; r += blockDim.x)
            s_peer[r] = r == rg.rank ? reinterpret_cast<unsigned long long>(src_cols[col])
                                     : reinterpret_cast<unsigned long long>(rg.peer_buf[2 * r + (int)(rg.my_mail[2 * r + 1] & 1ull)]);
        __syncthreads();
    }
    // broadcast sweep (see ladder_broadcast_tiles)?  bc_keep = tile-local mask of the non-target bits,
    // all ones when the sweep does not qualify for this column
    unsigned bc_keep = 0xffffffffu;
    if (sup_mode && !generate && staged_store && P.nrounds > 0) {
        unsigned tgt = 0;
        bool ok = true;
        for (int r = 0; r < P.nrounds; ++r) {
            const RoundDesc &R = P.rounds[r];
            for (int j = kRegBits - R.nsteps; j < kRegBits; ++j) {
                const unsigned tb = R.reg_tb[j];
                if (!((R.smask >> j) & 1u) || ((sup_vt >> tb) & 1u) || ((tgt >> tb) & 1u)) ok = false;
                tgt |= 1u << tb;
            }
        }
        if (ok && __popc(tgt) >= 2) bc_keep = ~tgt & ((1u << T) - 1u);
    }
    const bool bcast = bc_keep != 0xffffffffu;
    auto in_support = [&](unsigned long long o) {
        return ((outer_base(P.o_src, o, P.n_outer) ^ sup_g) & sup_mo) == 0ull;
    };
    auto write_zero_tile = [&](unsigned long long o) {
        double2 *__restrict__ q = dst + (outer_base(P.o_dst, o, P.n_outer) | s_off[2 * tid + 1]);
        const double2 z = make_double2(0.0, 0.0);
        if (staged_store) {
#pragma unroll
            for (int i = 0; i < kSlots; ++i) st_global_cs(q + P.st_off_hi[i], z);
            if (leaf_fuse && (tid & 31) == 0) leaf_col[(outer_base(P.o_dst, o, P.n_outer) | s_off[2 * tid + 1]) >> 10] = 0.0;
        } else {
#pragma unroll
            for (int s = 0; s < kSlots; ++s) st_global_cs(q + P.ds_slot[s], z);
        }
    };
    // sup_mode 1 with pinned outer bits: only ntiles >> popc(sup_mo) tiles hold data and the others are not
    // touched at all, so the CTAs walk a compact counter and deposit it into the free outer bits (testing
    // every tile of the grid cost 0.1 ms per launch at n = 30, for sweeps that touch one or 512 tiles)
    const bool compact = sup_mode == 1 && sup_mo != 0ull;
    auto walk_to_tile = [&](unsigned long long j) {
        unsigned long long t = 0;
        int b = 0;
        for (int i = 0; i < P.n_outer; ++i) {
            const int sp = P.osrc[i];
            if ((sup_mo >> sp) & 1ull) t |= ((sup_g >> sp) & 1ull) << i;
            else { t |= ((j >> b) & 1ull) << i; ++b; }
        }
        return t;
    };
    unsigned long long wj = blockIdx.x;
    // next tile of this CTA that holds data; tiles passed on the way are zero-filled in mode 2
    auto advance = [&](unsigned long long o) {
        if (compact) {
            wj += gridDim.x;
            return wj < (ntiles >> __popcll(sup_mo)) ? walk_to_tile(wj) : ntiles;
        }
        for (o += gridDim.x; o < ntiles; o += gridDim.x) {
            if (!sup_mo || in_support(o)) break;
            if (sup_mode == 2) write_zero_tile(o);
        }
        return o;
    };

    auto issue_loads = [&](unsigned long long o) {
        if (tma) {
            if (tid == 0) {
                // the buffer was read and written through the generic proxy up to the barrier just passed
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                mbar_expect_tx(Q1T_S_BAR, 16u << T);
                const int line = (int)(outer_base(P.o_src, o, P.n_outer) >> 3);
                for (int q = 0; q < P.tma_nreq; ++q)
                    tma_load_5d(Q1T_TILE_S0 + q * P.tma_req_bytes, &tmaps.m[col], Q1T_S_BAR, line + (int)P.tma_req_line[q]);
                if (P.prefetch_ahead > 0) {
                    const unsigned long long oa = o + (unsigned long long)P.prefetch_ahead * gridDim.x;
                    if (oa < ntiles) {
                        const int la = (int)(outer_base(P.o_src, oa, P.n_outer) >> 3);
                        for (int q = 0; q < P.tma_nreq; ++q) tma_prefetch_5d(&tmaps.m[col], la + (int)P.tma_req_line[q]);
                    }
                }
            }
            return;
        }
        const double2 *__restrict__ p = src + (outer_base(P.o_src, o, P.n_outer) | s_off[2 * tid]);
        unsigned sw = sw_tid;
        asm volatile("" : "+r"(sw));      // keeps the 32 destination addresses from being hoisted out of the tile loop (and spilled)
        if (REMOTE) {
            // element l of this rank's new shard lives on rank r[gb := l_lp] at index l[lp := r_gb]
            const unsigned long long t = outer_base(P.o_src, o, P.n_outer) | s_off[2 * tid];
            unsigned sr_t = rg_rank0;
            for (int j = 0; j < rg.k; ++j) sr_t |= (unsigned)((t >> rg.lp[j]) & 1ull) << rg.gb[j];
            const unsigned long long t2 = ((t & ~rg_lpmask) | rg_mybits) << 4;
            const unsigned s0 = Q1T_TILE_S0;
#pragma unroll
            for (int i = 0; i < kSlots; ++i) {
                const unsigned long long hi = P.ld_hi[i];
                unsigned sr = sr_t;
                for (int j = 0; j < rg.k; ++j) sr |= (unsigned)((hi >> rg.lp[j]) & 1ull) << rg.gb[j];
                cp_async16_s(s0 + (sw ^ P.ld_sw_hi[i]), s_peer[sr] + t2 + ((hi & ~rg_lpmask) << 4));
            }
        } else if (sup_mt == 0u) {
#if Q1T_LADDER_BYTE_ADDR
            // nine instructions per request (two constant loads, xor, + shared base, 64-bit index add, 64-bit scale-and-
            // add, request) cut to six: byte offsets from the program, the address kept as an integer
            unsigned long long pb = reinterpret_cast<unsigned long long>(p);
            asm volatile("" : "+l"(pb));
            const unsigned s0 = Q1T_TILE_S0;
#pragma unroll
            for (int i = 0; i < kSlots; ++i) cp_async16_s(s0 + (sw ^ P.ld_sw_hi[i]), pb + P.ld_hi_b[i]);
#else
#pragma unroll
            for (int i = 0; i < kSlots; ++i) cp_async16(tile_b + (sw ^ P.ld_sw_hi[i]), p + P.ld_hi[i]);
#endif
        } else {
            // only the elements inside the support exist in memory; the others are zeros that no round
            // reads (its loads are masked by zmask / smask), so they are not even written to the tile
#pragma unroll
            for (int i = 0; i < kSlots; ++i) {
                const unsigned e = tid | ((unsigned)i << TB);
                if (((e ^ sup_vt) & sup_mt) == 0u) cp_async16(tile_b + (sw ^ P.ld_sw_hi[i]), p + P.ld_hi[i]);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    // phase 1: angle of every phase op for tile o.  One warp per op, lanes over the outer bits,
    // butterfly sum; then one sincospi per lane.  (All warps take part: a single warp doing this
    // alone kept the other three waiting at the barrier.)
    auto tables_phase1 = [&](unsigned long long o) {
        if (blockDim.x < 32) {                 // tiny tiles (n < 10): no full warp, plain loop
            for (int pid = tid; pid < P.nphase; pid += blockDim.x) {
                double ang = s_coef[pid * ncoef + P.n_outer];
                for (int i = 0; i < P.n_outer; ++i)
                    if ((o >> i) & 1ull) ang += s_coef[pid * ncoef + i];
                double sn, cs;
                sincospi(ang, &sn, &cs);
                s_tileF[pid] = make_double2(cs, sn);
            }
            return;
        }
        // four ops per warp and pass: lane = (op slot p, bit group g); every lane adds the coefficients of
        // outer bits g, g+8, g+16, ..., three butterfly stages finish the sums, then one sincospi per op
        const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
        const int p = lane >> 3, g = lane & 7;
        for (int base = 0; base < P.nphase; base += 4 * nwarps) {
            const int pid = base + warp * 4 + p;
            double part = 0.0;
            if (pid < P.nphase) {
                if (g == 0) part = s_coef[pid * ncoef + P.n_outer];
                for (int i = g; i < P.n_outer; i += 8)
                    if ((o >> i) & 1ull) part += s_coef[pid * ncoef + i];
            }
#pragma unroll
            for (int off = 4; off >= 1; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
            if (g == 0 && pid < P.nphase) {
                double sn, cs;
                sincospi(part, &sn, &cs);
                s_tileF[pid] = make_double2(cs, sn);
            }
        }
    };
    // phase 2: s_hiF[e] = hi[e] * tile factor.  Entry tid of hi[] never changes: it stays in registers
    const double2 my_hi = (int)tid < ntab ? __ldg(reinterpret_cast<const double2 *>(ptabs[tid >> he_bits].hi) + (tid & ((1 << he_bits) - 1)))
                                          : make_double2(0.0, 0.0);
    auto tables_phase2 = [&](int buf) {
        if ((int)tid < ntab) s_hiF[buf * ntab + tid] = cmul(my_hi, s_tileF[tid >> he_bits]);
        for (int e = tid + blockDim.x; e < ntab; e += blockDim.x) {
            const double2 hi = __ldg(reinterpret_cast<const double2 *>(ptabs[e >> he_bits].hi) + (e & ((1 << he_bits) - 1)));
            s_hiF[buf * ntab + e] = cmul(hi, s_tileF[e >> he_bits]);
        }
    };

    if (bcast) {
        ladder_broadcast_tiles(src, dst, leaf_col, tile_b, s_off, bc_keep, sup_mt, sup_vt, sup_g, sup_mo);
        return;
    }
    unsigned long long o = blockIdx.x;
    if (compact) {
        if (o >= (ntiles >> __popcll(sup_mo))) return;
        o = walk_to_tile(o);
    } else {
        if (o >= ntiles) return;
        if (sup_mo && !in_support(o)) {
            if (sup_mode == 2) write_zero_tile(o);
            o = advance(o);
            if (o >= ntiles) return;
        }
    }
    if (!generate) issue_loads(o);
    for (int e = tid; e < (P.nphase << kThrLoBits); e += blockDim.x)
        s_lo[e] = __ldg(reinterpret_cast<const double2 *>(ptabs[e >> kThrLoBits].lo) + (e & ((1 << kThrLoBits) - 1)));
    for (int e = tid; e < P.nphase * ncoef; e += blockDim.x) {
        const int pid = e / ncoef, i = e - pid * ncoef;
        s_coef[e] = i < P.n_outer ? ptabs[pid].outer_coef[i] : ptabs[pid].base;
    }
    __syncthreads();
    tables_phase1(o);
    __syncthreads();
    tables_phase2(0);
    int buf = 0;
    unsigned long long o_next = 0;
    PCLK_DECL
    for (; o < ntiles; o = o_next, buf ^= 1) {
        o_next = advance(o);
        const bool has_next = o_next < ntiles;
        PCLK(0);
        if (tma) mbar_wait(Q1T_S_BAR, (unsigned)buf);          // tile k of this CTA completes phase k: parity = k & 1 = buf
        else if (!generate) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        PCLK(1);
        __syncthreads();                       // tile o and its tables are visible to the whole CTA
        PCLK(2);
        const double2 *const hiF = s_hiF + buf * ntab;
#if Q1T_LADDER_STATIC_ROUNDS
        // dense instantiation: the round loop unrolled (launched for nrounds <= 3 only), so that the round's tables are
        // constant-bank operands at fixed offsets instead of uniform loads through a runtime round index
#pragma unroll (dense ? 3 : 1)
        for (int r = 0; dense ? r < 3 : r < P.nrounds; ++r) {
            if (dense && r >= P.nrounds) break;
#else
        for (int r = 0; r < P.nrounds; ++r) {
#endif
            const RoundDesc &R = P.rounds[r];
            const bool last = r + 1 == P.nrounds;
            unsigned thrL = 0;
            for (int k = 0; k < R.nruns; ++k) thrL |= (unsigned)run_bits(tid, R.runs[k]);
            const unsigned swT = (tma ? tma_swizzle(thrL) : tile_swizzle(thrL)) * 16u;
            PCLK(3);
            if (r > 0) {
                if (R.sync_before == 2) __syncthreads();
                else __syncwarp();
            }
            PCLK(4);
            // a thread that differs from the basis index in a still-pinned thread bit holds only zeros
            const bool all_zero = !dense && ((thrL ^ sup_vt) & R.zmask) != 0u;
            double2 a[kSlots];
            if (r == 0 && generate) {
                // the basis element lives in exactly one tile of the grid and one slot of one thread
                unsigned gen_slot = 0xffffffffu;
                if ((gen_g & ~P.tile_mask_src) == outer_base(P.o_src, o, P.n_outer)) {
                    unsigned gen_l = 0, regmask = 0, slot = 0;
                    for (int i = 0; i < T; ++i) gen_l |= (unsigned)((gen_g >> P.tsrc[i]) & 1ull) << i;
                    for (int j = 0; j < kRegBits; ++j) {
                        regmask |= 1u << R.reg_tb[j];
                        slot |= ((gen_l >> R.reg_tb[j]) & 1u) << j;
                    }
                    if ((gen_l & ~regmask) == thrL) gen_slot = slot;
                }
#pragma unroll
                for (int s = 0; s < kSlots; ++s) a[s] = make_double2(s == (int)gen_slot ? P.gen_scale : 0.0, 0.0);
            } else if (dense || (R.zmask == 0u && R.smask == 0u)) {
#pragma unroll
                for (int s = 0; s < kSlots; ++s) a[s] = *reinterpret_cast<const double2 *>(tile_b + (swT ^ R.sw_slot[s]));
            } else {
                // support tracking: read only what can be non-zero
                unsigned vs = 0;
                for (int j = 0; j < kRegBits; ++j) vs |= ((sup_vt >> R.reg_tb[j]) & 1u) << j;
                vs &= R.smask;
#pragma unroll
                for (int s = 0; s < kSlots; ++s) {
                    if (!all_zero && (((unsigned)s ^ vs) & R.smask) == 0u) a[s] = *reinterpret_cast<const double2 *>(tile_b + (swT ^ R.sw_slot[s]));
                    else a[s] = make_double2(0.0, 0.0);
                }
            }
            PCLK(5);
            if (last && !staged_store) {
                if (has_next) tables_phase1(o_next);
                __syncthreads();               // all amplitudes of the tile are in registers: the buffer is dead
                if (has_next) {
                    if (!generate) issue_loads(o_next);
                    tables_phase2(buf ^ 1);
                }
            }
            PCLK(6);
            const unsigned il = tid & ((1u << kThrLoBits) - 1u), ih = tid >> kThrLoBits;
#if Q1T_LADDER_UNIFIED
            if (!all_zero) LadderU<0>::run(a, R, P, hiF, he_bits, il, ih, s_lo, kRegBits - (int)R.nsteps);
#else
            if (!all_zero)
            switch (R.nsteps) {
            case 1: LadderSteps<1, kRegBits - 1, true>::run(a, R, P, ptabs, hiF, he_bits, il, ih, s_lo); break;
            case 2: LadderSteps<2, kRegBits - 2, true>::run(a, R, P, ptabs, hiF, he_bits, il, ih, s_lo); break;
            case 3: LadderSteps<3, kRegBits - 3, true>::run(a, R, P, ptabs, hiF, he_bits, il, ih, s_lo); break;
#if Q1T_REG_BITS >= 5
            case 4: LadderSteps<4, kRegBits - 4, true>::run(a, R, P, ptabs, hiF, he_bits, il, ih, s_lo); break;
            default: LadderSteps<5, kRegBits - 5, true>::run(a, R, P, ptabs, hiF, he_bits, il, ih, s_lo); break;
#else
            default: LadderSteps<4, kRegBits - 4, true>::run(a, R, P, ptabs, hiF, he_bits, il, ih, s_lo); break;
#endif
            }
#endif
            PCLK(7);
            if (last && !staged_store) {
                double2 *__restrict__ q = dst + (outer_base(P.o_dst, o, P.n_outer) | s_off[2 * tid + 1]);
                if (scale != 1.0) {
#pragma unroll
                    for (int s = 0; s < kSlots; ++s) a[s] = make_double2(a[s].x * scale, a[s].y * scale);
                }
#if Q1T_LADDER_BYTE_ADDR
                unsigned long long qb = reinterpret_cast<unsigned long long>(q);
                asm volatile("" : "+l"(qb));
#pragma unroll
                for (int s = 0; s < kSlots; ++s) st_global_cs_b(qb + P.ds_slot_b[s], a[s]);
#else
#pragma unroll
                for (int s = 0; s < kSlots; ++s) st_global_cs(q + P.ds_slot[s], a[s]);
#endif
            } else if (!all_zero || last) {
                // (zeros of an all-zero thread are never read by a later round; the store pass after
                //  the last round reads everything)
                // the 32 addresses are recomputed (a constant load and an xor each): kept from the round's loads they
                // do not fit beside 32 amplitudes and were spilled to local memory, 31 stores + 31 loads per round
                unsigned swS = swT;
                asm volatile("" : "+r"(swS));
#pragma unroll
                for (int s = 0; s < kSlots; ++s) *reinterpret_cast<double2 *>(tile_b + (swS ^ R.sw_slot[s])) = a[s];
            }
        }
        PCLK(8);
        if (staged_store) {
            // the destination layout differs from the last round's thread layout (fused relabel):
            // one more trip through shared memory, stores coalesced in destination order
            __syncthreads();
            PCLK(10);
            double2 *__restrict__ q = dst + (outer_base(P.o_dst, o, P.n_outer) | s_off[2 * tid + 1]);
            unsigned swl = sw_lo;
            asm volatile("" : "+r"(swl));     // as in issue_loads: no hoisting of the 32 addresses
#if Q1T_LADDER_BYTE_ADDR
            unsigned long long qb = reinterpret_cast<unsigned long long>(q);
            asm volatile("" : "+l"(qb));
#define Q1T_STAGED_ST(i, x) st_global_cs_b(qb + P.st_off_hi_b[i], x)
#else
#define Q1T_STAGED_ST(i, x) st_global_cs(q + P.st_off_hi[i], x)
#endif
            double leaf_acc = 0.0;
            {   // first half: read and store right away
                double2 v[kSlots / 2];
#pragma unroll
                for (int i = 0; i < kSlots / 2; ++i) v[i] = *reinterpret_cast<const double2 *>(tile_b + (swl ^ P.st_l_hi[i]));
#pragma unroll
                for (int i = 0; i < kSlots / 2; ++i) {
                    const double2 x = scale != 1.0 ? make_double2(v[i].x * scale, v[i].y * scale) : v[i];
                    if (leaf_fuse) leaf_acc = __dadd_rn(leaf_acc, __dadd_rn(__dmul_rn(x.x, x.x), __dmul_rn(x.y, x.y)));
                    Q1T_STAGED_ST(i, x);
                }
            }
            PCLK(11);
            {   // second half: once it is in registers the tile is dead -> prefetch, then store
                double2 v[kSlots / 2];
#pragma unroll
                for (int i = 0; i < kSlots / 2; ++i) v[i] = *reinterpret_cast<const double2 *>(tile_b + (swl ^ P.st_l_hi[kSlots / 2 + i]));
                if (has_next) tables_phase1(o_next);
                PCLK(12);
                __syncthreads();
                PCLK(13);
                if (has_next) {
                    if (!generate) issue_loads(o_next);
                    tables_phase2(buf ^ 1);
                }
                PCLK(14);
#pragma unroll
                for (int i = 0; i < kSlots / 2; ++i) {
                    const double2 x = scale != 1.0 ? make_double2(v[i].x * scale, v[i].y * scale) : v[i];
                    if (leaf_fuse) leaf_acc = __dadd_rn(leaf_acc, __dadd_rn(__dmul_rn(x.x, x.x), __dmul_rn(x.y, x.y)));
                    Q1T_STAGED_ST(kSlots / 2 + i, x);
                }
            }
            if (leaf_fuse) {
                // lane l has summed elements l, l+32, ... of its warp's leaf in increasing order: finish
                // with the canonical 5-stage butterfly (same arithmetic as leaf_totals_kernel)
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) leaf_acc = __dadd_rn(leaf_acc, __shfl_xor_sync(0xffffffffu, leaf_acc, off));
                if ((tid & 31) == 0) leaf_col[(outer_base(P.o_dst, o, P.n_outer) | s_off[2 * tid + 1]) >> 10] = leaf_acc;
            }
        }
        PCLK(9);
    }
#ifdef Q1T_PHASE_CLOCKS
    if (blockIdx.x == 7 && (tid & 31) == 0)
        printf("PCLK warp %d: adv %lld ldwait %lld topbar %lld | roundhdr %lld sync %lld lds %lld dead+pref %lld compute %lld | store/sts %lld staged %lld"
               " [bar %lld half1 %lld lds2+ph1 %lld bar %lld pref %lld stg2->9]\n",
               tid >> 5, pclk[0], pclk[1], pclk[2], pclk[3], pclk[4], pclk[5], pclk[6], pclk[7], pclk[8], pclk[9], pclk[10], pclk[11], pclk[12], pclk[13], pclk[14]);
#endif
}

// threads per CTA: 2^(T - kRegBits).  The ladder-only kernel for tiles up to 2^12 runs with several
// CTAs per SM; everything else (op interpreter, 2^13 tiles) with one.
constexpr int kSmallThreads = 1 << (12 - kRegBits);
constexpr int kMaxThreads = 1 << kMaxThrBits;
#ifndef Q1T_LADDER_MIN_CTAS
#define Q1T_LADDER_MIN_CTAS (Q1T_REG_BITS >= 5 ? 3 : 2)
#endif

bool sweep_uses_ladder_kernel(const SweepProgram &prog)
{
    static const bool pipe = !(std::getenv("Q1T_LADDER_PIPE") && std::atoi(std::getenv("Q1T_LADDER_PIPE")) == 0);
    if (!pipe || prog.nrounds <= 0 || (1 << prog.TB) > kSmallThreads) return false;
    for (int r = 0; r < prog.nrounds; ++r)
        if (prog.rounds[r].kind != ROUND_PH) return false;
    return true;
}

// tensor maps of the source columns for a program in TMA layout (SweepProgram::tma_*); the driver entry point
// is looked up through the runtime, so the library does not link against libcuda
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else cudaGetLastError();
    }
    return fn;
}
bool tma_make_maps(const SweepProgram &prog, const double2 *const *h_src_cols, int ncols, TmaMaps &out)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || prog.tma_nreq <= 0 || ncols > kMaxTmaCols || !h_src_cols) return false;
    std::memset(&out, 0, sizeof out);
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t box[5], estr[5] = { 1, 1, 1, 1, 1 };
    for (int i = 0; i < 5; ++i) { gdim[i] = prog.tma_gdim[i]; box[i] = prog.tma_box[i]; }
    for (int i = 0; i < 4; ++i) gstride[i] = prog.tma_gstride[i];
    for (int c = 0; c < ncols; ++c) {
        const CUresult r = fn(&out.m[c], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, const_cast<double2 *>(h_src_cols[c]), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
    }
    return true;
}
bool tma_available() { return encode_tiled_fn() != nullptr; }
bool tma_can_encode(const SweepProgram &prog, const double2 *const *h_src_cols, int ncols)
{
    TmaMaps tmp;
    const bool ok = tma_make_maps(prog, h_src_cols, ncols, tmp);
    if (!ok) {
        static bool warned = false;
        if (!warned) {
            warned = true;
            std::fprintf(stderr, "q1tsim_b200: cuTensorMapEncodeTiled rejected a tile description (box %u %u %u %u, strides %llu %llu %llu); "
                                 "this sweep shape keeps cp.async loads\n", prog.tma_box[1], prog.tma_box[2], prog.tma_box[3], prog.tma_box[4],
                         (unsigned long long)prog.tma_gstride[0], (unsigned long long)prog.tma_gstride[1], (unsigned long long)prog.tma_gstride[2]);
        }
    }
    return ok;
}

// c_prog is ONE constant bank per device and process.  States of one process that share a device (several shards on one
// GPU: how the single-process sharded state is tested on a one-GPU box) launch from different streams, so the upload of
// one's program must not overtake the other's running sweep: while more than one state is alive on a device, every
// (upload, kernel) pair waits for the previous one on that device.  The usual case -- one state per device -- pays one
// load.
namespace {
struct CprogGuard {
    std::mutex mu;
    int live[64] = {};            // states (streams) alive on the device
    cudaEvent_t evt[64] = {};
} g_cprog;
}  // namespace
bool cprog_device_shared(int dev) { return dev >= 0 && dev < 64 && g_cprog.live[dev] > 1; }
// a state (one stream) appears on / leaves a device.  The second one to appear waits for what the first has in flight.
void cprog_stream_register(int dev)
{
    if (dev < 0 || dev >= 64) return;
    std::lock_guard<std::mutex> lk(g_cprog.mu);
    if (++g_cprog.live[dev] == 2) {
        cudaDeviceSynchronize();
        if (!g_cprog.evt[dev]) cudaEventCreateWithFlags(&g_cprog.evt[dev], cudaEventDisableTiming);
        cudaGetLastError();
    }
}
void cprog_stream_unregister(int dev)
{
    if (dev < 0 || dev >= 64) return;
    std::lock_guard<std::mutex> lk(g_cprog.mu);
    if (g_cprog.live[dev] > 0) --g_cprog.live[dev];
}
static cudaError_t cprog_acquire(int dev, cudaStream_t stream)
{
    if (!cprog_device_shared(dev) || !g_cprog.evt[dev]) return cudaSuccess;
    return cudaStreamWaitEvent(stream, g_cprog.evt[dev], 0);       // (an event never recorded yet does not block)
}
static void cprog_release(int dev, cudaStream_t stream)
{
    if (cprog_device_shared(dev) && g_cprog.evt[dev]) cudaEventRecord(g_cprog.evt[dev], stream);
}

bool sweep_can_gather_remote(const SweepProgram &prog)
{
    return sweep_uses_ladder_kernel(prog) && prog.tma_nreq == 0 && !prog.generate && prog.sup_mode == 0 && prog.nrounds <= 3;
}

cudaError_t launch_sweep(const SweepProgram &prog, const double2 *const *d_src_cols, double2 *const *d_dst_cols,
                         int ncols, const PhaseTab *d_ptabs, const unsigned long long *d_gen_idx, cudaStream_t stream,
                         double *d_leaf_out, const double2 *const *h_src_cols, const SweepProgram *d_prog, const RemoteGather *remote)
{
    if (remote && remote->k > 0 && (!sweep_can_gather_remote(prog) || ncols != 1)) return cudaErrorInvalidValue;
    TmaMaps tmaps;
    if (prog.tma_nreq > 0) {
        // a program in TMA layout can only run in the ladder kernel with tensor maps of its source columns
        if (!sweep_uses_ladder_kernel(prog) || prog.generate || prog.sup_mode || !tma_make_maps(prog, h_src_cols, ncols, tmaps))
            return cudaErrorInvalidValue;
    } else std::memset(&tmaps, 0, sizeof tmaps);
    const bool ladder = sweep_uses_ladder_kernel(prog);
    const bool gprog = d_prog != nullptr && !ladder;          // (the ladder kernel keeps its program in the constant bank)
    cudaError_t e = cudaSuccess;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!gprog) {
        e = cprog_acquire(dev, stream);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyToSymbolAsync(c_prog, &prog, sizeof(SweepProgram), 0, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return e;
    }
    const int he_bits = prog.TB > kThrLoBits ? prog.TB - kThrLoBits : 0;
    const size_t smem = (sizeof(double2) << prog.T) + (sizeof(double2) * (size_t)prog.nphase << he_bits);
    typedef void (*SweepFn)(const double2 *const *, double2 *const *, const PhaseTab *, const unsigned long long *, const SweepProgram *);
    static const SweepFn sfn[6] = { sweep_kernel<true, kMaxThreads, 1, false>, sweep_kernel<false, kMaxThreads, 1, false>,
                                    sweep_kernel<false, kSmallThreads, Q1T_LADDER_MIN_CTAS, false>,
                                    sweep_kernel<true, kMaxThreads, 1, true>, sweep_kernel<false, kMaxThreads, 1, true>,
                                    sweep_kernel<false, kSmallThreads, Q1T_LADDER_MIN_CTAS, true> };
    static int sattr_dev_mask = 0;             // function attributes are per device
    if (!((sattr_dev_mask >> (dev & 31)) & 1)) {
        const int max_smem = (int)((sizeof(double2) << kMaxTileBits) + sizeof(double2) * kMaxPhase * kHiEntries);
        const int small_smem = (int)((sizeof(double2) << 12) + sizeof(double2) * kMaxPhase * kHiEntries);
        for (int v = 0; v < 6; ++v) {
            e = cudaFuncSetAttribute(sfn[v], cudaFuncAttributeMaxDynamicSharedMemorySize, (v % 3) == 2 ? small_smem : max_smem);
            if (e != cudaSuccess) return e;
        }
        for (int v = 2; v < 6; v += 3) {
            e = cudaFuncSetAttribute(sfn[v], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return e;
        }
        sattr_dev_mask |= 1 << (dev & 31);
    }
    bool ladders_only = true;
    for (int r = 0; r < prog.nrounds; ++r) ladders_only = ladders_only && prog.rounds[r].kind == ROUND_PH;
    dim3 grid((unsigned)(1ull << prog.n_outer), (unsigned)ncols, 1);
    dim3 block(1u << prog.TB, 1, 1);
    if (ladder) {
        typedef void (*LadderFn)(const double2 *const *, double2 *const *, const PhaseTab *, const unsigned long long *, double *, const TmaMaps,
                                 const RemoteGather);
        static const LadderFn fns[4] = { ladder_kernel<kSmallThreads, Q1T_LADDER_MIN_CTAS, false, false, false>,
                                         ladder_kernel<kSmallThreads, Q1T_LADDER_MIN_CTAS, true, true, false>,
                                         ladder_kernel<kSmallThreads, Q1T_LADDER_MIN_CTAS, false, true, false>,
                                         ladder_kernel<kSmallThreads, Q1T_LADDER_MIN_CTAS, false, true, true> };
        static int attr_dev_mask[4] = { 0, 0, 0, 0 };      // function attributes are per device
        static const bool dense_variant = !(std::getenv("Q1T_LADDER_DENSE") && std::atoi(std::getenv("Q1T_LADDER_DENSE")) == 0);
        int nsm = 148, occ = 0;
        const bool gather = remote && remote->k > 0;
        const int variant = gather ? 3 : prog.tma_nreq > 0 ? 1 : (dense_variant && !prog.generate && prog.sup_mode == 0 && prog.nrounds <= 3) ? 2 : 0;
        RemoteGather rg;
        std::memset(&rg, 0, sizeof rg);
        if (gather) rg = *remote;
        const LadderFn fn = fns[variant];
        if (!((attr_dev_mask[variant] >> (dev & 31)) & 1)) {
            const int max_lsmem = (int)((sizeof(double2) << 12) + sizeof(double2) * kMaxPhase * (3 * kHiEntries + 1 + (1 << kThrLoBits)) + sizeof(double) * kMaxPhase * (kMaxBits + 1) + 16 * kSmallThreads + 32 + 256);
            e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_lsmem);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return e;
            attr_dev_mask[variant] |= 1 << (dev & 31);
        }
        const size_t lsmem = (sizeof(double2) << prog.T) + 16 +
                             sizeof(double2) * (((size_t)prog.nphase << he_bits) * 2 + prog.nphase + ((size_t)prog.nphase << kThrLoBits)) +
                             sizeof(double) * (((size_t)prog.nphase * (prog.n_outer + 1) + 1) & ~(size_t)1) + 16 * (size_t)block.x + (gather ? 256 : 0);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, (int)block.x, lsmem);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
        // persistent grid: every column gets the same share of the resident CTAs
        unsigned long long per_col = (unsigned long long)nsm * occ / (unsigned)ncols;
        if (per_col < 1) per_col = 1;
        if (per_col > (1ull << prog.n_outer)) per_col = 1ull << prog.n_outer;
        dim3 pgrid((unsigned)per_col, (unsigned)ncols, 1);
        fn<<<pgrid, block, lsmem, stream>>>(d_src_cols, d_dst_cols, d_ptabs, d_gen_idx, d_leaf_out, tmaps, rg);
        e = cudaGetLastError();
        cprog_release(dev, stream);
        return e;
    }
    const int v = (ladders_only && (int)block.x <= kSmallThreads ? 2 : ladders_only ? 1 : 0) + (gprog ? 3 : 0);
    sfn[v]<<<grid, block, smem, stream>>>(d_src_cols, d_dst_cols, d_ptabs, d_gen_idx, d_prog);
    e = cudaGetLastError();
    if (!gprog) cprog_release(dev, stream);
    return e;
}

// ---------------------------------------------------------------------------
// generic (unfused) gate: one thread per group of 2^k amplitudes
// ---------------------------------------------------------------------------
__global__ void generic_gate_kernel(double2 *const *__restrict__ cols, int n, int k, GenericGateArgs g,
                                    const double2 *__restrict__ mat)
{
    const unsigned long long ngroups = 1ull << (n - k);
    double2 *__restrict__ st = cols[blockIdx.y];
    const int G = 1 << k;
    for (unsigned long long grp = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; grp < ngroups;
         grp += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long base = grp;
        for (int a = 0; a < k; ++a) {
            const int p = g.sorted_pos[a];
            base = ((base >> p) << (p + 1)) | (base & ((1ull << p) - 1ull));
        }
        if ((base & g.cmask) != g.cmask) continue;
        double2 in[1 << kMaxGenericBits];
        for (int h = 0; h < G; ++h) in[h] = st[base | g.offs[h]];
        for (int i = 0; i < G; ++i) {
            double2 acc = make_double2(0.0, 0.0);
            for (int h = 0; h < G; ++h) {
                const double2 m = mat[i * G + h];
                acc.x = fma(in[h].x, m.x, fma(-in[h].y, m.y, acc.x));
                acc.y = fma(in[h].x, m.y, fma(in[h].y, m.x, acc.y));
            }
            st[base | g.offs[i]] = acc;
        }
    }
}

// The same with the 2^K amplitudes of a group in REGISTERS (K <= 5: the runtime-indexed array above lives in local memory)
// and the matrix in shared memory (every thread reads the same entry: a broadcast).  One thread per group; consecutive
// threads hold consecutive groups, so a load of slot h is coalesced whenever the low index bits are not targets.
template <int K>
__global__ void __launch_bounds__(128)
generic_gate_reg_kernel(double2 *const *__restrict__ cols, int n, GenericGateArgs g, const double2 *__restrict__ mat)
{
    constexpr int G = 1 << K;
    __shared__ double2 s_mat[G * G];
    for (int e = threadIdx.x; e < G * G; e += blockDim.x) s_mat[e] = mat[e];
    __syncthreads();
    const unsigned long long ngroups = 1ull << (n - K);
    double2 *__restrict__ st = cols[blockIdx.y];
    for (unsigned long long grp = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; grp < ngroups;
         grp += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long base = grp;
#pragma unroll
        for (int a = 0; a < K; ++a) {
            const int p = g.sorted_pos[a];
            base = ((base >> p) << (p + 1)) | (base & ((1ull << p) - 1ull));
        }
        if ((base & g.cmask) != g.cmask) continue;
        double2 in[G];
#pragma unroll
        for (int h = 0; h < G; ++h) in[h] = st[base | g.offs[h]];
#pragma unroll 1
        for (int i = 0; i < G; ++i) {
            double2 acc = make_double2(0.0, 0.0);
#pragma unroll
            for (int h = 0; h < G; ++h) {
                const double2 m = s_mat[i * G + h];
                acc.x = fma(in[h].x, m.x, fma(-in[h].y, m.y, acc.x));
                acc.y = fma(in[h].x, m.y, fma(in[h].y, m.x, acc.y));
            }
            st[base | g.offs[i]] = acc;
        }
    }
}

cudaError_t launch_generic_gate(double2 *const *d_cols, int ncols, int n, int k, const GenericGateArgs &g,
                                const double2 *d_mat, cudaStream_t stream)
{
    const unsigned long long ngroups = 1ull << (n - k);
    unsigned blocks = (unsigned)((ngroups + 127) / 128);
    if (blocks > 148u * 32u) blocks = 148u * 32u;
    if (blocks == 0) blocks = 1;
    const dim3 grid(blocks, ncols);
    switch (k) {
    case 1: generic_gate_reg_kernel<1><<<grid, 128, 0, stream>>>(d_cols, n, g, d_mat); break;
    case 2: generic_gate_reg_kernel<2><<<grid, 128, 0, stream>>>(d_cols, n, g, d_mat); break;
    case 3: generic_gate_reg_kernel<3><<<grid, 128, 0, stream>>>(d_cols, n, g, d_mat); break;
    case 4: generic_gate_reg_kernel<4><<<grid, 128, 0, stream>>>(d_cols, n, g, d_mat); break;
    case 5: generic_gate_reg_kernel<5><<<grid, 128, 0, stream>>>(d_cols, n, g, d_mat); break;
    default: generic_gate_kernel<<<grid, 128, 0, stream>>>(d_cols, n, k, g, d_mat); break;
    }
    return cudaGetLastError();
}

// Dense blocks on 6..10 targets (gates.rs:310-325 takes any k).  kBigGroups groups of 2^k amplitudes are staged in shared
// memory; thread i accumulates output i of every staged group, reading column i of the matrix once per pass (from shared
// memory while the matrix fits, k <= 6, else coalesced from the transposed copy in L2) and the inputs as broadcasts.
constexpr int kBigGroups = 4;
__global__ void __launch_bounds__(256)
generic_gate_big_kernel(double2 *const *__restrict__ cols, GenericBigArgs a, const double2 *__restrict__ matT, int mat_in_smem)
{
    extern __shared__ double2 s_big[];
    const int k = a.k, G = 1 << k;
    double2 *const s_in = s_big;                                                        // [kBigGroups][G]
    unsigned long long *const s_offs = reinterpret_cast<unsigned long long *>(s_in + (kBigGroups << k));   // [G]
    unsigned long long *const s_base = s_offs + G;                                      // [kBigGroups]
    double2 *const s_mat = reinterpret_cast<double2 *>(s_base + kBigGroups);            // [G][G] when it fits
    for (int h = threadIdx.x; h < G; h += blockDim.x) {
        unsigned long long o = 0;
        for (int j = 0; j < k; ++j)
            if ((h >> (k - 1 - j)) & 1) o |= 1ull << a.pos[j];
        s_offs[h] = o;
    }
    if (mat_in_smem)
        for (int e = threadIdx.x; e < G * G; e += blockDim.x) s_mat[e] = matT[e];
    const double2 *__restrict__ mt = mat_in_smem ? s_mat : matT;
    double2 *__restrict__ st = cols[blockIdx.y];
    const unsigned long long ngroups = 1ull << (a.n - k);
    for (unsigned long long g0 = (unsigned long long)blockIdx.x * kBigGroups; g0 < ngroups; g0 += (unsigned long long)gridDim.x * kBigGroups) {
        __syncthreads();                      // tables written / the previous pass has read its inputs
        if (threadIdx.x < kBigGroups) {
            unsigned long long base = ~0ull;                                            // ~0: nothing to do for this slot
            const unsigned long long grp = g0 + threadIdx.x;
            if (grp < ngroups) {
                base = grp;
                for (int j = 0; j < k; ++j) {
                    const int p = a.sorted_pos[j];
                    base = ((base >> p) << (p + 1)) | (base & ((1ull << p) - 1ull));
                }
                if ((base & a.cmask) != a.cmask) base = ~0ull;
            }
            s_base[threadIdx.x] = base;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < (kBigGroups << k); e += blockDim.x) {
            const unsigned long long base = s_base[e >> k];
            s_in[e] = base != ~0ull ? st[base | s_offs[e & (G - 1)]] : make_double2(0.0, 0.0);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < G; i += blockDim.x) {
            double2 acc[kBigGroups];
#pragma unroll
            for (int gi = 0; gi < kBigGroups; ++gi) acc[gi] = make_double2(0.0, 0.0);
            for (int h = 0; h < G; ++h) {
                const double2 m = mt[(size_t)h * G + i];
#pragma unroll
                for (int gi = 0; gi < kBigGroups; ++gi) {
                    const double2 x = s_in[(gi << k) + h];
                    acc[gi].x = fma(x.x, m.x, fma(-x.y, m.y, acc[gi].x));
                    acc[gi].y = fma(x.x, m.y, fma(x.y, m.x, acc[gi].y));
                }
            }
            const unsigned long long off = s_offs[i];
#pragma unroll
            for (int gi = 0; gi < kBigGroups; ++gi)
                if (s_base[gi] != ~0ull) st[s_base[gi] | off] = acc[gi];               // (every input of the pass sits in shared memory)
        }
    }
}
cudaError_t launch_generic_gate_big(double2 *const *d_cols, int ncols, const GenericBigArgs &g, const double2 *d_matT, cudaStream_t stream)
{
    if (g.k < 1 || g.k > kMaxBigGenericBits || g.k > g.n) return cudaErrorInvalidValue;
    const size_t G = (size_t)1 << g.k;
    const bool mat_in_smem = g.k <= 6;
    const size_t smem = sizeof(double2) * kBigGroups * G + sizeof(unsigned long long) * (G + kBigGroups) + (mat_in_smem ? sizeof(double2) * G * G : 0);
    static int attr_mask = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((attr_mask >> (dev & 31)) & 1)) {
        const cudaError_t e = cudaFuncSetAttribute(generic_gate_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) return e;
        attr_mask |= 1 << (dev & 31);
    }
    const unsigned long long ngroups = 1ull << (g.n - g.k);
    unsigned long long blocks = (ngroups + kBigGroups - 1) / kBigGroups;
    if (blocks > 148ull * 4ull) blocks = 148ull * 4ull;
    const int threads = G < 256 ? (G < 32 ? 32 : (int)G) : 256;
    generic_gate_big_kernel<<<dim3((unsigned)blocks, ncols), threads, smem, stream>>>(d_cols, g, d_matT, mat_in_smem ? 1 : 0);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// canonical-order reductions (DESIGN.md "canonical reduction order")
//   leaf  = min(2^n, 1024) consecutive amplitudes, one warp per leaf:
//           lane l accumulates elements l, l+32, ... sequentially, then a
//           5-stage xor butterfly (16, 8, 4, 2, 1);  p = fl(fl(re*re)+fl(im*im))
//   block = 1024 leaves chained sequentially; block totals chained sequentially
// ---------------------------------------------------------------------------
__device__ __forceinline__ double norm_sqr_rn(double2 v)
{
    return __dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y));
}

__global__ void __launch_bounds__(256)
leaf_totals_kernel(const double2 *const *__restrict__ cols, double *__restrict__ leaf_out, int n, int leaf_bits,
                   unsigned long long mask, unsigned long long want)
{
    const unsigned long long nleaves = 1ull << (n - leaf_bits);
    const int lane = threadIdx.x & 31;
    const unsigned long long warp = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const double2 *__restrict__ st = cols[blockIdx.y];
    double *__restrict__ out = leaf_out + (unsigned long long)blockIdx.y * nleaves;
    const int leaf = 1 << leaf_bits;
    for (unsigned long long L = warp; L < nleaves; L += nwarps) {
        const unsigned long long first = L << leaf_bits;
        double acc = 0.0;
        // whole leaf masked out (mask bit above the leaf): total is exactly +0.0, skip the reads
        const unsigned long long hi_mask = mask & ~((unsigned long long)leaf - 1ull);
        if ((first & hi_mask) == (want & hi_mask)) {
            if (leaf >= 32) {
#pragma unroll 8
                for (int t = 0; t < leaf; t += 32) {
                    const unsigned long long idx = first + t + lane;
                    if ((idx & mask) == want) acc = __dadd_rn(acc, norm_sqr_rn(__ldcs(&st[idx])));
                }
            } else if (lane < leaf) {
                const unsigned long long idx = first + lane;
                if ((idx & mask) == want) acc = __dadd_rn(acc, norm_sqr_rn(st[idx]));
            }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
        if (lane == 0) out[L] = acc;
    }
}

// one warp per block of 1024 leaves: stage into shared memory, lane 0 chains sequentially
__global__ void __launch_bounds__(128)
block_scan_kernel(double *__restrict__ leaf_io, double *__restrict__ block_totals, unsigned long long nleaves,
                  unsigned long long nblocks)
{
    __shared__ double s[4][kCanonBlock];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long b = blockIdx.x * 4ull + w;
    double *__restrict__ leaf = leaf_io + (unsigned long long)blockIdx.y * nleaves;
    if (b < nblocks) {
        const unsigned long long first = b * kCanonBlock;
        const unsigned long long cnt = (nleaves - first) < kCanonBlock ? (nleaves - first) : kCanonBlock;
        for (unsigned i = lane; i < cnt; i += 32) s[w][i] = leaf[first + i];
        __syncwarp();
        if (lane == 0) {
            double run = 0.0;
            for (unsigned i = 0; i < cnt; ++i) {
                run = __dadd_rn(run, s[w][i]);
                s[w][i] = run;
            }
            block_totals[(unsigned long long)blockIdx.y * nblocks + b] = run;
        }
        __syncwarp();
        for (unsigned i = lane; i < cnt; i += 32) leaf[first + i] = s[w][i];
    }
}

// one CTA per column: inclusive chain over block totals (in place).  The totals are staged through shared memory
// with coalesced loads (a thread chaining straight from global memory pays a DRAM round trip per block), thread 0
// chains them sequentially -- the canonical order -- and the CTA writes the prefixes back.
__global__ void __launch_bounds__(256)
top_chain_kernel(double *__restrict__ block_io, unsigned long long nblocks, int ncols, double *__restrict__ totals_out)
{
    __shared__ double s[2048];
    const int c = blockIdx.x;
    if (c >= ncols) return;
    double *__restrict__ bt = block_io + (unsigned long long)c * nblocks;
    double run = 0.0;
    for (unsigned long long first = 0; first < nblocks; first += 2048) {
        const unsigned cnt = (unsigned)((nblocks - first) < 2048ull ? (nblocks - first) : 2048ull);
        for (unsigned i = threadIdx.x; i < cnt; i += blockDim.x) s[i] = bt[first + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            for (unsigned i = 0; i < cnt; ++i) {
                run = __dadd_rn(run, s[i]);
                s[i] = run;
            }
        }
        __syncthreads();
        for (unsigned i = threadIdx.x; i < cnt; i += blockDim.x) bt[first + i] = s[i];
        __syncthreads();
    }
    if (threadIdx.x == 0) totals_out[c] = run;
}

__device__ __forceinline__ double leaf_prefix(const double *__restrict__ inblock, const double *__restrict__ bpref,
                                              unsigned long long L, double base0)
{
    const unsigned long long b = L / kCanonBlock;
    return __dadd_rn(b ? bpref[b - 1] : base0, inblock[L]);
}

// one warp per draw (draws sorted ascending on the host): leaf by binary search over the canonical leaf prefixes
// (lane 0's result broadcast), then the in-leaf scan: the warp loads 32 consecutive amplitudes at a time (512 B,
// coalesced, the next chunk already in flight) and lane 0's running sum visits them in index order through shuffles --
// the same sequential chain as before, without a dependent DRAM round trip per 8 elements.
__global__ void __launch_bounds__(256)
resolve_draws_kernel(const double2 *__restrict__ st, const double *__restrict__ inblock,
                     const double *__restrict__ bpref, int n, int leaf_bits,
                     const double *__restrict__ chosen, unsigned long long ndraws,
                     unsigned long long *__restrict__ idx_out, double base, double base0)
{
    const int lane = threadIdx.x & 31;
    const unsigned long long j = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
    if (j >= ndraws) return;
    const double ch = chosen[j];
    const unsigned long long nleaves = 1ull << (n - leaf_bits);
    unsigned long long lo = 0, hi = nleaves - 1;
    while (lo < hi) {                                   // (every lane runs the same search: uniform loads)
        const unsigned long long mid = (lo + hi) >> 1;
        if (leaf_prefix(inblock, bpref, mid, base0) <= ch) lo = mid + 1; else hi = mid;
    }
    double run = lo == 0 ? base : leaf_prefix(inblock, bpref, lo - 1, base0);   // base: weight before this shard
    const unsigned leaf = 1u << leaf_bits;
    const double2 *__restrict__ p = st + (lo << leaf_bits);
    unsigned found = 0xffffffffu, last_nz = 0xffffffffu;
    double w_next = (unsigned)lane < leaf ? norm_sqr_rn(p[lane]) : 0.0;
    for (unsigned e0 = 0; e0 < leaf && found == 0xffffffffu; e0 += 32) {
        const double w = w_next;
        if (e0 + 32 < leaf) w_next = e0 + 32 + lane < leaf ? norm_sqr_rn(p[e0 + 32 + lane]) : 0.0;
        const unsigned m = (leaf - e0) < 32u ? (leaf - e0) : 32u;
        // the chain runs redundantly in every lane on shuffled values: no divergence, same order
#pragma unroll 8
        for (unsigned q = 0; q < 32; ++q) {
            const double wq = __shfl_sync(0xffffffffu, w, q);
            if (q < m && found == 0xffffffffu) {
                if (wq > 0.0) last_nz = e0 + q;
                run = __dadd_rn(run, wq);
                if (ch < run) found = e0 + q;
            }
        }
    }
    if (found == 0xffffffffu) found = last_nz == 0xffffffffu ? leaf - 1 : last_nz;
    if (lane == 0) idx_out[j] = (lo << leaf_bits) + found;
}

cudaError_t launch_leaf_totals(const double2 *const *d_cols, int ncols, double *d_leaf, int n,
                               unsigned long long mask, unsigned long long want, cudaStream_t stream)
{
    const int leaf_bits = n < kCanonLeafBits ? n : kCanonLeafBits;
    const unsigned long long nleaves = 1ull << (n - leaf_bits);
    unsigned long long blocks = (nleaves + 7) / 8;      // 8 warps per CTA
    if (blocks > 148ull * 64ull) blocks = 148ull * 64ull;
    leaf_totals_kernel<<<dim3((unsigned)blocks, ncols), 256, 0, stream>>>(d_cols, d_leaf, n, leaf_bits, mask, want);
    return cudaGetLastError();
}

// in-block inclusive prefixes (in place) and the raw block totals only: the chain over blocks continues on other ranks
cudaError_t launch_block_scan(double *d_leaf, double *d_block, int ncols, int n, cudaStream_t stream)
{
    const int leaf_bits = n < kCanonLeafBits ? n : kCanonLeafBits;
    const unsigned long long nleaves = 1ull << (n - leaf_bits);
    const unsigned long long nblocks = (nleaves + kCanonBlock - 1) / kCanonBlock;
    block_scan_kernel<<<dim3((unsigned)((nblocks + 3) / 4), ncols), 128, 0, stream>>>(d_leaf, d_block, nleaves, nblocks);
    return cudaGetLastError();
}

cudaError_t launch_scan(double *d_leaf, double *d_block, double *d_totals, int ncols, int n, cudaStream_t stream)
{
    const int leaf_bits = n < kCanonLeafBits ? n : kCanonLeafBits;
    const unsigned long long nleaves = 1ull << (n - leaf_bits);
    const unsigned long long nblocks = (nleaves + kCanonBlock - 1) / kCanonBlock;
    block_scan_kernel<<<dim3((unsigned)((nblocks + 3) / 4), ncols), 128, 0, stream>>>(d_leaf, d_block, nleaves, nblocks);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    top_chain_kernel<<<ncols, 256, 0, stream>>>(d_block, nblocks, ncols, d_totals);
    return cudaGetLastError();
}

cudaError_t launch_resolve_draws(const double2 *d_col, const double *d_leaf, const double *d_block, int n,
                                 const double *d_chosen, unsigned long long ndraws, unsigned long long *d_idx,
                                 double base, cudaStream_t stream, double base0)
{
    if (ndraws == 0) return cudaSuccess;
    const int leaf_bits = n < kCanonLeafBits ? n : kCanonLeafBits;
    resolve_draws_kernel<<<(unsigned)((ndraws + 7) / 8), 256, 0, stream>>>(d_col, d_leaf, d_block, n, leaf_bits, d_chosen,
                                                                        ndraws, d_idx, base, base0);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// collapse: out0 keeps the bit==0 half scaled by f0, out1 keeps the bit==1
// half scaled by f1 (either may be null); everything else is zero.
// vectorstate.rs:91-104 multiplies by Complex(1/sqrt(w), 0): both parts scaled.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
collapse_kernel(const double2 *in, double2 *out0, double2 *out1,   // in may alias out0 or out1
                unsigned long long N, unsigned long long bit, double f0, double f1)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < N;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const double2 v = in[i];
        const bool one = (i & bit) != 0;
        // (re*f - im*0, re*0 + im*f) as num-complex does; +0.0 terms dropped
        const double2 z = make_double2(0.0, 0.0);
        if (out0) out0[i] = one ? z : make_double2(__dmul_rn(v.x, f0), __dmul_rn(v.y, f0));
        if (out1) out1[i] = one ? make_double2(__dmul_rn(v.x, f1), __dmul_rn(v.y, f1)) : z;
    }
}

cudaError_t launch_collapse(const double2 *d_in, double2 *d_out0, double2 *d_out1, int n, int bitpos, double f0,
                            double f1, cudaStream_t stream)
{
    const unsigned long long N = 1ull << n;
    unsigned long long blocks = (N + 255) / 256;
    if (blocks > 148ull * 16ull) blocks = 148ull * 16ull;
    collapse_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_in, d_out0, d_out1, N, 1ull << bitpos, f0, f1);
    return cudaGetLastError();
}

// product state of per-qubit coefficient pairs, same multiplication chain as
// vectorstate.rs:62-83 / cmatrix.rs:40-51 (kron_mat): cur = b_q[bit_q] * cur,
// textbook complex product without FMA contraction.
__global__ void product_state_kernel(double2 *__restrict__ st, int n, const double2 *__restrict__ coefs)
{
    const unsigned long long N = 1ull << n;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < N;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        double2 cur = make_double2(1.0, 0.0);
        for (int q = 0; q < n; ++q) {
            const int bit = (int)((i >> (n - 1 - q)) & 1ull);
            const double2 b = coefs[2 * q + bit];
            const double re = __dsub_rn(__dmul_rn(b.x, cur.x), __dmul_rn(b.y, cur.y));
            const double im = __dadd_rn(__dmul_rn(b.x, cur.y), __dmul_rn(b.y, cur.x));
            cur = make_double2(re, im);
        }
        st[i] = cur;
    }
}
cudaError_t launch_product_state(double2 *d_col, int n, const double2 *d_coefs, cudaStream_t stream)
{
    const unsigned long long N = 1ull << n;
    unsigned long long blocks = (N + 255) / 256;
    if (blocks > 148ull * 16ull) blocks = 148ull * 16ull;
    product_state_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_col, n, d_coefs);
    return cudaGetLastError();
}

// out0 = in * f0, out1 = in * f1 (either output may be null, out0 may alias in): collapse of a
// qubit that is a rank bit of a sharded state -- the whole shard is kept (scaled) or zeroed
__global__ void __launch_bounds__(256)
scale2_kernel(const double2 *in, double2 *out0, double2 *out1, unsigned long long N, double f0, double f1)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < N;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const double2 v = in[i];
        if (out0) out0[i] = make_double2(__dmul_rn(v.x, f0), __dmul_rn(v.y, f0));
        if (out1) out1[i] = make_double2(__dmul_rn(v.x, f1), __dmul_rn(v.y, f1));
    }
}
cudaError_t launch_scale2(const double2 *d_in, double2 *d_out0, double2 *d_out1, int n, double f0, double f1, cudaStream_t stream)
{
    const unsigned long long N = 1ull << n;
    unsigned long long blocks = (N + 255) / 256;
    if (blocks > 148ull * 16ull) blocks = 148ull * 16ull;
    scale2_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_in, d_out0, d_out1, N, f0, f1);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// qubit remap between two ranks over NVLink peer memory (multi-GPU).
// Rank bit a of this rank and index bit L of the shard trade places: my element at
// index l (l_L = 1-a) is exchanged with the partner's element at l ^ (1<<L).  The
// 2^(n-1) pairs are split between the two ranks (this rank takes the half whose top
// remaining bit equals a), so every pair is swapped exactly once, in place, with
// no staging buffer: one remote read + one remote write per 16 bytes handled.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
peer_swap_kernel(double2 *__restrict__ mine, double2 *__restrict__ theirs, int n, int L, int a)
{
    const unsigned long long npairs = 1ull << (n - 2);       // pairs handled by this rank
    const unsigned long long sel = (unsigned long long)a << (n - 2);
    for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < npairs;
         p += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long q = sel | p;                 // n-1 bits: all index bits except L
        const unsigned long long lo = q & ((1ull << L) - 1ull);
        const unsigned long long base = ((q >> L) << (L + 1)) | lo;
        const unsigned long long lm = base | ((unsigned long long)(1 - a) << L);
        const unsigned long long lt = base | ((unsigned long long)a << L);
        const double2 x = mine[lm];
        const double2 y = theirs[lt];
        mine[lm] = y;
        theirs[lt] = x;
    }
}
cudaError_t launch_peer_swap(double2 *d_mine, double2 *d_theirs, int n, int L, int a, cudaStream_t stream)
{
    const unsigned long long npairs = 1ull << (n - 2);
    unsigned long long blocks = (npairs + 255) / 256;
    if (blocks > 148ull * 32ull) blocks = 148ull * 32ull;
    if (blocks == 0) blocks = 1;
    peer_swap_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_mine, d_theirs, n, L, a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// peer group (one process per GPU): device-side barrier and the multi-bit qubit remap
// over NVLink peer memory.
//
// Every rank owns a small mailbox in its own HBM that all peers have mapped (CUDA IPC): slot 2r holds the
// last barrier epoch rank r has reached, slot 2r+1 which of its two registered shard buffers is current.
// group_barrier_kernel: one lane per peer writes (current buffer, then epoch) into that peer's mailbox with
// system-scope release stores and then waits, with system-scope acquire loads, until the peer's epoch has
// arrived in its own mailbox.  Stream-ordered: the host never blocks, and everything enqueued before the
// barrier on any rank is complete and visible to everything enqueued after it on every rank.
// ---------------------------------------------------------------------------
void group_kernels_preload();
__global__ void group_barrier_kernel(unsigned long long *const *__restrict__ peer_mail, unsigned long long *my_mail, int P, int rank,
                                     unsigned long long epoch, unsigned long long cur)
{
    const int t = threadIdx.x;
    if (t >= P || t == rank) return;
    unsigned long long *theirs = peer_mail[t];
    // the buffer index goes into the slot of this epoch's parity: a rank that is already past its next barrier (which
    // publishes the buffer it moved INTO) must not change what a slower peer's kernel after THIS barrier still has to read
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;\n" ::"l"(theirs + 2 * rank + 1 + 64 * (epoch & 1ull)), "l"(cur) : "memory");
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(theirs + 2 * rank), "l"(epoch) : "memory");
    unsigned long long seen = 0;
    for (unsigned long long spins = 0;; ++spins) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(seen) : "l"(my_mail + 2 * t) : "memory");
        if (seen >= epoch) break;
        __nanosleep(100);
        if (spins > 200000000ull) __trap();      // ~20 s: a peer that died must not hang this device for good
    }
}
cudaError_t launch_group_barrier(unsigned long long *const *d_peer_mail, unsigned long long *d_my_mail, int P, int rank,
                                 unsigned long long epoch, unsigned long long cur, cudaStream_t stream)
{
    group_kernels_preload();
    group_barrier_kernel<<<1, 32, 0, stream>>>(d_peer_mail, d_my_mail, P, rank, epoch, cur);
    return cudaGetLastError();
}

// Multi-bit remap: k rank bits gb[j] and k index bits lp[j] of the shard trade places, in place, in ONE pass.
// Element (rank r, index l) goes to (r', l'): r' = r with bit gb[j] := l_lp[j], l' = l with bit lp[j] := r_gb[j]
// -- an involution, so the data moves in pairs between two ranks (or stays, when the two bit patterns agree).
// Of every pair, one rank does the swap: the lower rank for the pairs whose index bit `split` is 0, the higher
// one for the others; each rank therefore reads and writes (2^k - 1) 2^(n-k-1) remote amplitudes, spread
// over its 2^k - 1 partners (blockIdx.y), all NVLink directions busy at once.
__global__ void __launch_bounds__(256)
group_swap_kernel(double2 *__restrict__ mine, void *const *__restrict__ peer_buf, const unsigned long long *__restrict__ my_mail, GroupRemapArgs a)
{
    unsigned apat = 0;
    for (int j = 0; j < a.k; ++j) apat |= (unsigned)((a.rank >> a.gb[j]) & 1) << j;
    // interleave: consecutive CTAs serve different partners (all of them busy from the first wave on); else the
    // partners are served one after the other (blockIdx.y), each as one perfect matching of the ranks
    const unsigned npartners = (1u << a.k) - 1u;
    const unsigned py = a.interleave ? blockIdx.x % npartners : blockIdx.y;
    const unsigned long long bx = a.interleave ? blockIdx.x / npartners : blockIdx.x;
    const unsigned long long gx = a.interleave ? gridDim.x / npartners : gridDim.x;
    const unsigned bpat = py < apat ? py : py + 1;
    int partner = a.rank;
    unsigned long long mine_or = 0, theirs_or = 0;
    for (int j = 0; j < a.k; ++j) {
        const unsigned bj = (bpat >> j) & 1u, aj = (apat >> j) & 1u;
        partner = (partner & ~(1 << a.gb[j])) | ((int)bj << a.gb[j]);
        mine_or |= (unsigned long long)bj << a.lp[j];
        theirs_or |= (unsigned long long)aj << a.lp[j];
    }
    const unsigned long long c = a.rank < partner ? 0ull : 1ull;
    mine_or |= c << a.split;
    theirs_or |= c << a.split;
    const unsigned long long cur = my_mail[2 * partner + 1];          // published by the partner in the barrier before this kernel
    double2 *__restrict__ theirs = static_cast<double2 *>(peer_buf[2 * partner + (int)(cur & 1ull)]);
    const unsigned long long npairs = 1ull << (a.n - a.k - 1);
    for (unsigned long long t = bx * (unsigned long long)blockDim.x + threadIdx.x; t < npairs;
         t += gx * blockDim.x) {
        unsigned long long base = t;
        for (int i = 0; i <= a.k; ++i) {                              // open a zero bit at every swapped position (ascending)
            const int p = a.ins[i];
            base = ((base >> p) << (p + 1)) | (base & ((1ull << p) - 1ull));
        }
        const double2 x = mine[base | mine_or];
        const double2 y = theirs[base | theirs_or];
        mine[base | mine_or] = y;
        theirs[base | theirs_or] = x;
    }
}
cudaError_t launch_group_swap(double2 *d_mine, void *const *d_peer_buf, const unsigned long long *d_my_mail, const GroupRemapArgs &a,
                              cudaStream_t stream)
{
    const unsigned long long npairs = 1ull << (a.n - a.k - 1);
    unsigned long long blocks = (npairs + 255) / 256;
    const unsigned long long cap = (148ull * 16ull) / ((1u << a.k) - 1u) + 1;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    const unsigned np = (1u << a.k) - 1u;
    if (a.interleave) group_swap_kernel<<<dim3((unsigned)blocks * np, 1), 256, 0, stream>>>(d_mine, d_peer_buf, d_my_mail, a);
    else group_swap_kernel<<<dim3((unsigned)blocks, np), 256, 0, stream>>>(d_mine, d_peer_buf, d_my_mail, a);
    return cudaGetLastError();
}

// The remap as an out-of-place gather (option fused_remap, when no sweep follows that could read through the trade):
// dst[l] = shard of rank r[gb := l_lp] at index l[lp := r_gb].  Every rank only READS its peers' current buffers and
// writes its own other buffer, so it may run beside peers that do the same read inside a ladder sweep.
__global__ void __launch_bounds__(256)
group_gather_kernel(const double2 *__restrict__ mine, double2 *__restrict__ dst, RemoteGather rg, unsigned long long nelem)
{
    __shared__ unsigned long long s_peer[32];
    if (threadIdx.x < 32 && (int)threadIdx.x < rg.P)
        s_peer[threadIdx.x] = (int)threadIdx.x == rg.rank
                                  ? reinterpret_cast<unsigned long long>(mine)
                                  : reinterpret_cast<unsigned long long>(rg.peer_buf[2 * threadIdx.x + (int)(rg.my_mail[2 * threadIdx.x + 1] & 1ull)]);
    __syncthreads();
    unsigned long long lpmask = 0, mybits = 0;
    unsigned gbmask = 0;
    for (int j = 0; j < rg.k; ++j) {
        lpmask |= 1ull << rg.lp[j];
        mybits |= (unsigned long long)((rg.rank >> rg.gb[j]) & 1) << rg.lp[j];
        gbmask |= 1u << rg.gb[j];
    }
    const unsigned rank0 = (unsigned)rg.rank & ~gbmask;
    for (unsigned long long l = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; l < nelem; l += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned sr = rank0;
        for (int j = 0; j < rg.k; ++j) sr |= (unsigned)((l >> rg.lp[j]) & 1ull) << rg.gb[j];
        const double2 *__restrict__ sp = reinterpret_cast<const double2 *>(s_peer[sr]);
        st_global_cs(dst + l, __ldcs(sp + ((l & ~lpmask) | mybits)));
    }
}
cudaError_t launch_group_gather(const double2 *d_mine, double2 *d_dst, const RemoteGather &rg, int n, cudaStream_t stream)
{
    const unsigned long long nelem = 1ull << n;
    unsigned long long blocks = (nelem + 255) / 256;
    if (blocks > 148ull * 16ull) blocks = 148ull * 16ull;
    group_gather_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_mine, d_dst, rg, nelem);
    return cudaGetLastError();
}

// CUDA loads a kernel at its first launch (lazy module loading), and loading may synchronise the context: a swap kernel
// launched for the first time while a barrier kernel of ANOTHER shard of this process spins on the same device would wait
// for that barrier, which waits for a barrier this thread has not launched yet.  Both kernels are loaded before the first
// barrier goes out (per device).
void group_kernels_preload()
{
    static int done_mask = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if ((done_mask >> (dev & 31)) & 1) return;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, group_barrier_kernel);
    cudaFuncGetAttributes(&fa, group_swap_kernel);
    cudaFuncGetAttributes(&fa, group_gather_kernel);
    cudaFuncGetAttributes(&fa, ladder_kernel<kSmallThreads, Q1T_LADDER_MIN_CTAS, false, true, true>);
    cudaGetLastError();
    done_mask |= 1 << (dev & 31);
}

__global__ void set_basis_kernel(double2 *__restrict__ st, unsigned long long idx)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) st[idx] = make_double2(1.0, 0.0);
}
cudaError_t launch_set_basis(double2 *d_col, unsigned long long idx, cudaStream_t stream)
{
    set_basis_kernel<<<1, 32, 0, stream>>>(d_col, idx);
    return cudaGetLastError();
}

}  // namespace q1t
