// planner.cpp -- gate lowering + fusion planner (host only).
//
// Lowering consumes only `Gate::matrix()` (gates.rs:174), as SURVEY 8(b)
// prescribes: user gates define nothing else.  The unitary is classified into
//   POLY    diagonal 1/2-qubit unitary  -> phase polynomial (Z,S,T,RZ,U1,CZ,CS,CT,CU1,CRZ..)
//   G1      [controls] + one non-diagonal 2x2 target (H,X,Y,V,RX,RY,U2,U3,CX,CCX,CH,CU3..)
//   SWAP    qubit relabel, zero bytes moved (swap.rs:78-88)
//   GENERIC anything else (dense k-target block), unfused fallback kernel
#include "planner.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>

namespace q1t {

double wrap_half_turns(double a)
{
    if (a >= -1.0 && a < 1.0) return a;
    double r = std::fmod(a, 2.0);      // exact
    if (r >= 1.0) r -= 2.0;
    if (r < -1.0) r += 2.0;
    return r;
}

static void sincospi_host(double a, double &s, double &c)
{
    a = wrap_half_turns(a);
    const double t = 2.0 * a;          // exact multiples of 1/2 turn-halves get exact values
    if (t == std::floor(t)) {
        const int q = ((int)t % 4 + 4) % 4;
        static const double cs[4] = { 1, 0, -1, 0 }, sn[4] = { 0, 1, 0, -1 };
        c = cs[q]; s = sn[q];
        return;
    }
    s = std::sin(M_PI * a);
    c = std::cos(M_PI * a);
}

static inline uint64_t deposit(uint64_t v, const int *positions, int nbits)
{
    uint64_t r = 0;
    for (int i = 0; i < nbits; ++i) r |= ((v >> i) & 1ull) << positions[i];
    return r;
}

// tid bit i -> target bit pos[i]: group consecutive tid bits with equal displacement into
// (mask, shift) runs.  Ascending pos gives shifts >= 0; the per-round thread maps and the direct
// store may be in any order (negative shifts), which the kernels handle through run_bits()
static int make_runs(const int *pos, int nbits, BitRun *runs)
{
    int nr = 0;
    int i = 0;
    while (i < nbits) {
        int j = i;
        uint32_t mask = 0;
        while (j < nbits && pos[j] - j == pos[i] - i) { mask |= 1u << j; ++j; }
        runs[nr].mask = mask;
        runs[nr].shift = pos[i] - i;
        ++nr;
        i = j;
    }
    return nr;
}

static void fill_outer_tables(SweepProgram &P)
{
    for (int c = 0; c < kOuterChunks; ++c)
        for (int v = 0; v < (1 << kOuterChunkBits); ++v) {
            uint64_t so = 0, dof = 0;
            for (int b = 0; b < kOuterChunkBits; ++b) {
                const int i = c * kOuterChunkBits + b;
                if (i >= P.n_outer || !((v >> b) & 1)) continue;
                so |= 1ull << P.osrc[i];
                dof |= 1ull << P.odst[i];
            }
            P.o_src[c][v] = so;
            P.o_dst[c][v] = dof;
        }
    // destination-ordered walk: walk bit i = the outer bit with the i-th lowest destination position
    int ord[kMaxBits];
    for (int i = 0; i < P.n_outer; ++i) ord[i] = i;
    std::sort(ord, ord + P.n_outer, [&](int a, int b) { return P.odst[a] < P.odst[b]; });
    for (int c = 0; c < kOuterChunks; ++c)
        for (int v = 0; v < (1 << kOuterChunkBits); ++v) {
            uint64_t so = 0, dof = 0;
            for (int b = 0; b < kOuterChunkBits; ++b) {
                const int i = c * kOuterChunkBits + b;
                if (i >= P.n_outer || !((v >> b) & 1)) continue;
                so |= 1ull << P.osrc[ord[i]];
                dof |= 1ull << P.odst[ord[i]];
            }
            P.w_src[c][v] = so;
            P.w_dst[c][v] = dof;
        }
}

// Decide which rounds can skip shared-memory staging and how wide each inter-round barrier must be.
//  - round 0 may read straight from global memory when tile bits 0,1,2 (128 contiguous bytes) are its
//    three lowest thread bits; the last round may write straight to global memory when the three
//    lowest destination bits are its three lowest thread bits;
//  - between two rounds a __syncwarp is enough when every warp-index bit (tid bit >= 5) stands for
//    the same tile bit in both rounds: each warp then re-reads only what it wrote itself.
static void setup_direct(SweepProgram &P)
{
    P.tile_mask_src = 0;
    for (int i = 0; i < P.T; ++i) P.tile_mask_src |= 1ull << P.tsrc[i];
    const int C = P.coalesce;
    P.direct_load = P.direct_store = 0;
    P.dl_nruns = P.ds_nruns = 0;
    for (int r = 0; r < P.nrounds; ++r) {
        RoundDesc &R = P.rounds[r];
        R.sync_before = 2;
        if (r > 0) {
            const RoundDesc &Q = P.rounds[r - 1];
            bool warp_local = true;
            for (int i = 5; i < P.TB; ++i)
                if (R.thr_tb[i] != Q.thr_tb[i]) warp_local = false;
            if (warp_local) R.sync_before = 1;
        }
    }
    if (P.nrounds == 0 || P.TB < C) return;
    {
        const RoundDesc &R = P.rounds[0];
        bool ok = true;
        for (int i = 0; i < C; ++i) ok = ok && R.thr_tb[i] == i && P.tsrc[i] == i;
        for (int i = 1; i < P.TB; ++i) ok = ok && R.thr_tb[i] > R.thr_tb[i - 1];
        if (ok) {
            int pos[kMaxThrBits + 1];
            for (int i = 0; i < P.TB; ++i) pos[i] = P.tsrc[R.thr_tb[i]];
            P.dl_nruns = make_runs(pos, P.TB, P.dl_runs);
            for (int s = 0; s < kSlots; ++s) {
                uint64_t off = 0;
                for (int j = 0; j < kRegBits; ++j)
                    if ((s >> j) & 1) off |= 1ull << P.tsrc[R.reg_tb[j]];
                P.dl_slot[s] = off;
            }
            P.direct_load = 1;
        }
    }
    {
        const RoundDesc &R = P.rounds[P.nrounds - 1];
        bool ok = true;
        int pos[kMaxThrBits + 1];
        for (int i = 0; i < P.TB; ++i) {
            pos[i] = P.tdst[R.thr_tb[i]];
            if (i < C && pos[i] != i) ok = false;
        }
        if (ok) {
            P.ds_nruns = make_runs(pos, P.TB, P.ds_runs);
            for (int s = 0; s < kSlots; ++s) {
                uint64_t off = 0;
                for (int j = 0; j < kRegBits; ++j)
                    if ((s >> j) & 1) off |= 1ull << P.tdst[R.reg_tb[j]];
                P.ds_slot[s] = off;
            }
            P.direct_store = 1;
        }
    }
}

// ---------------------------------------------------------------------------
// lowering
// ---------------------------------------------------------------------------
bool lower_gate(const cplx *mat, int k, const int *phys, LoweredGate &out, std::string &err)
{
    out = LoweredGate();
    const int G = 1 << k;
    for (int a = 0; a < k; ++a)
        for (int b = a + 1; b < k; ++b)
            if (phys[a] == phys[b]) { err = "duplicate qubit index in gate bits"; return false; }

    bool diagonal = true;
    for (int r = 0; r < G && diagonal; ++r)
        for (int c = 0; c < G; ++c)
            if (r != c && (mat[r * G + c].real() != 0.0 || mat[r * G + c].imag() != 0.0)) { diagonal = false; break; }

    if (diagonal && k <= 2) {
        bool unit = true;
        double a[4] = { 0, 0, 0, 0 };
        for (int g = 0; g < G; ++g) {
            const double mod = std::abs(mat[g * G + g]);
            if (std::fabs(mod - 1.0) > 1e-9) unit = false;
            a[g] = std::atan2(mat[g * G + g].imag(), mat[g * G + g].real()) / M_PI;
        }
        if (unit) {
            out.kind = LoweredGate::POLY;
            out.nb = k;
            if (k == 1) {
                out.b[0] = phys[0];
                out.c0 = wrap_half_turns(a[0]);
                out.lin[0] = wrap_half_turns(a[1] - a[0]);
            } else {
                out.b[0] = phys[0]; out.b[1] = phys[1];
                out.c0 = wrap_half_turns(a[0]);
                out.lin[0] = wrap_half_turns(a[2] - a[0]);
                out.lin[1] = wrap_half_turns(a[1] - a[0]);
                out.quad = wrap_half_turns(a[3] - a[2] - a[1] + a[0]);
            }
            return true;
        }
    }

    // control extraction
    std::vector<int> act(k);
    for (int j = 0; j < k; ++j) act[j] = j;
    std::vector<cplx> M(mat, mat + (size_t)G * G);
    uint64_t cmask = 0;
    int last_ctl = -1;
    bool found = true;
    while (found && !act.empty()) {
        found = false;
        const int na = (int)act.size(), D = 1 << na;
        for (int a = 0; a < na && !found; ++a) {
            const int bit = 1 << (na - 1 - a);
            bool ctl = true;
            for (int r = 0; r < D && ctl; ++r)
                for (int c = 0; c < D; ++c) {
                    if ((r & bit) && (c & bit)) continue;
                    const cplx e = M[(size_t)r * D + c];
                    const double want = r == c ? 1.0 : 0.0;
                    if (e.real() != want || e.imag() != 0.0) { ctl = false; break; }
                }
            if (!ctl) continue;
            // reduce to the bit==1 block
            std::vector<cplx> M2((size_t)(D / 2) * (D / 2));
            int rr = 0;
            for (int r = 0; r < D; ++r) {
                if (!(r & bit)) continue;
                int cc = 0;
                for (int c = 0; c < D; ++c) {
                    if (!(c & bit)) continue;
                    M2[(size_t)rr * (D / 2) + cc] = M[(size_t)r * D + c];
                    ++cc;
                }
                ++rr;
            }
            M.swap(M2);
            cmask |= 1ull << phys[act[a]];
            last_ctl = act[a];
            act.erase(act.begin() + a);
            found = true;
        }
    }
    if (act.empty()) {
        const cplx s = M[0];
        if (s.real() == 1.0 && s.imag() == 0.0) {   // identity gate
            out.kind = LoweredGate::POLY;
            out.nb = 0;
            return true;
        }
        // all bits were controls of a scalar phase: make the last one the target of diag(1, s)
        out.kind = LoweredGate::G1;
        out.target = phys[last_ctl];
        out.cmask = cmask & ~(1ull << phys[last_ctl]);
        const double mm[8] = { 1, 0, 0, 0, 0, 0, s.real(), s.imag() };
        std::memcpy(out.m, mm, sizeof mm);
        return true;
    }
    if (act.size() == 1) {
        out.kind = LoweredGate::G1;
        out.target = phys[act[0]];
        out.cmask = cmask;
        for (int e = 0; e < 4; ++e) { out.m[2 * e] = M[e].real(); out.m[2 * e + 1] = M[e].imag(); }
        return true;
    }
    if (act.size() == 2 && cmask == 0) {
        static const double sw[16] = { 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1 };
        bool is_swap = true;
        for (int e = 0; e < 16; ++e)
            if (M[e].real() != sw[e] || M[e].imag() != 0.0) is_swap = false;
        if (is_swap) {
            out.kind = LoweredGate::SWAP;
            out.b[0] = phys[act[0]]; out.b[1] = phys[act[1]];
            return true;
        }
    }
    if ((int)act.size() > 10) { err = "dense gate blocks on more than 10 target qubits are not supported"; return false; }
    out.kind = LoweredGate::GENERIC;
    out.cmask = cmask;
    for (int a : act) out.pos.push_back(phys[a]);
    out.mat = M;
    return true;
}

// ---------------------------------------------------------------------------
// pending diagonal terms
// ---------------------------------------------------------------------------
bool PendingDiag::touches(int t) const
{
    if (lin[t] != 0.0) return true;
    for (int p = 0; p < n; ++p)
        if (quad[(size_t)t * n + p] != 0.0) return true;
    return false;
}
bool PendingDiag::empty() const
{
    if (c0 != 0.0) return false;
    for (int t = 0; t < n; ++t)
        if (touches(t)) return false;
    return true;
}

// ---------------------------------------------------------------------------
// planner
// ---------------------------------------------------------------------------
Planner::Planner(int n, int tile_bits, int coalesce_bits, bool balance, int mid_relabel)
    : n_(n), T_(std::min(tile_bits, n)), C_(std::max(1, std::min(coalesce_bits, 3))), balance_(balance), mid_relabel_(mid_relabel)
{
    cur_.resize(n);
    inv_.resize(n);
    for (int p = 0; p < n; ++p) cur_[p] = inv_[p] = p;
    pd_.init(n);
    open_sweep();
}

void Planner::open_sweep()
{
    tile_.clear();
    rounds_.clear();
    nops_ = nphase_ = 0;
    for (int p = 0; p < C_ && p < n_; ++p) tile_.push_back(inv_[p]);  // coalescing bits: 8 amplitudes = 128 B (4 = 64 B)
}

bool Planner::in_tile(int p) const { return std::find(tile_.begin(), tile_.end(), p) != tile_.end(); }

std::vector<int> Planner::upcoming(size_t max_count) const
{
    std::vector<int> out;
    for (size_t i = la_at_; i < la_.size() && out.size() < max_count; ++i)
        if (std::find(out.begin(), out.end(), la_[i]) == out.end()) out.push_back(la_[i]);
    return out;
}

bool Planner::has_pending() const { return !rounds_.empty() || !pd_.empty(); }

bool Planner::place_target(int t)
{
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (!rounds_.empty()) {
            RoundB &r = rounds_.back();
            if (std::find(r.regs.begin(), r.regs.end(), t) != r.regs.end()) return true;
            if ((int)r.regs.size() < kRegBits && (in_tile(t) || (int)tile_.size() < T_)) {
                if (!in_tile(t)) tile_.push_back(t);
                r.regs.push_back(t);
                return true;
            }
        }
        // balance: a further round whose register bits the tile can no longer supply (fewer than
        // kRegBits free tile bits left) costs a full shared-memory round trip for a few steps; a
        // fresh sweep takes those steps along for free if it has to come anyway
        const bool starve = balance_ && attempt == 0 && !rounds_.empty() && !in_tile(t) &&
                            T_ - (int)tile_.size() < kRegBits && (int)rounds_.back().regs.size() >= kRegBits;
        if (!starve && (int)rounds_.size() < kMaxRounds && (in_tile(t) || (int)tile_.size() < T_)) {
            if (!in_tile(t)) tile_.push_back(t);
            RoundB r;
            r.regs.push_back(t);
            rounds_.push_back(r);
            return true;
        }
        close_sweep();
        open_sweep();
    }
    return false;
}

void Planner::emit_op(const OpB &op)
{
    rounds_.back().ops.push_back(op);
    ++nops_;
    if (op.is_phase) ++nphase_;
}

void Planner::emit_phase_for(int t)
{
    OpB op;
    op.is_phase = true;
    op.target = t;
    op.base = pd_.lin[t];
    for (int p = 0; p < n_; ++p) {
        if (p == t) continue;
        const double v = pd_.q(t, p);
        if (v != 0.0) op.partners.push_back(std::make_pair(p, v));
        pd_.q(t, p) = 0.0;
        pd_.q(p, t) = 0.0;
    }
    pd_.lin[t] = 0.0;
    if (pd_.c0 != 0.0) { op.has_c0 = true; op.c0 = pd_.c0; pd_.c0 = 0.0; }
    if (op.base == 0.0 && op.partners.empty() && !op.has_c0) return;
    emit_op(op);
}

void Planner::add(const LoweredGate &g)
{
    if (g.kind == LoweredGate::POLY) {
        pd_.c0 = wrap_half_turns(pd_.c0 + g.c0);
        for (int i = 0; i < g.nb; ++i) pd_.lin[g.b[i]] = wrap_half_turns(pd_.lin[g.b[i]] + g.lin[i]);
        if (g.nb == 2 && g.quad != 0.0) {
            const double v = wrap_half_turns(pd_.q(g.b[0], g.b[1]) + g.quad);
            pd_.q(g.b[0], g.b[1]) = v;
            pd_.q(g.b[1], g.b[0]) = v;
        }
        return;
    }
    // G1
    const int t = g.target;
    const size_t need_ops = 2;
    struct Count { size_t &c; ~Count() { ++c; } } count_this_gate{ la_at_ };      // (the gate still counts as upcoming while it is placed)
    if (nops_ + need_ops > (size_t)kMaxOps || nphase_ + 1 > (size_t)kMaxPhase) {
        close_sweep();
        open_sweep();
    }
    place_target(t);
    if (pd_.touches(t)) emit_phase_for(t);
    OpB op;
    op.target = t;
    op.cmask = g.cmask;
    std::memcpy(op.m, g.m, sizeof op.m);
    const double *m = g.m;
    const bool real = m[1] == 0 && m[3] == 0 && m[5] == 0 && m[7] == 0;
    if (real && m[0] == m[2] && m[0] == m[4] && m[0] == -m[6]) op.kind = OP_G1_HADAMARD;
    else if (m[0] == 0 && m[1] == 0 && m[6] == 0 && m[7] == 0) {
        op.kind = (m[2] == 1 && m[3] == 0 && m[4] == 1 && m[5] == 0) ? OP_G1_SWAPX : OP_G1_ANTIDIAG;
    } else if (m[2] == 0 && m[3] == 0 && m[4] == 0 && m[5] == 0) op.kind = OP_G1_DIAG;
    else op.kind = OP_G1_GENERIC;
    emit_op(op);
}

void Planner::flush_diag_touching(uint64_t bits_mask)
{
    for (int t = 0; t < n_; ++t) {
        if (!((bits_mask >> t) & 1ull)) continue;
        if (!pd_.touches(t)) continue;
        if (nops_ + 1 > (size_t)kMaxOps || nphase_ + 1 > (size_t)kMaxPhase) { close_sweep(); open_sweep(); }
        place_target(t);
        emit_phase_for(t);
    }
}

void Planner::finish()
{
    for (;;) {
        // pick the bit with the most pending partners; prefer one already in registers / the tile
        int best = -1, best_score = -1;
        for (int t = 0; t < n_; ++t) {
            if (!pd_.touches(t)) continue;
            int deg = 0;
            for (int p = 0; p < n_; ++p)
                if (p != t && pd_.q(t, p) != 0.0) ++deg;
            int score = deg * 4;
            if (!rounds_.empty() && std::find(rounds_.back().regs.begin(), rounds_.back().regs.end(), t) != rounds_.back().regs.end()) score += 3;
            else if (in_tile(t)) score += 1;
            if (score > best_score) { best_score = score; best = t; }
        }
        if (best < 0 || best_score < 4) break;        // only bits with quadratic partners need their own op
        if (nops_ + 1 > (size_t)kMaxOps || nphase_ + 1 > (size_t)kMaxPhase) { close_sweep(); open_sweep(); }
        place_target(best);
        emit_phase_for(best);
    }
    // what is left is separable: c0 + sum_b lin_b x_b  ->  one LINPHASE op in a round of its own
    bool any_lin = pd_.c0 != 0.0;
    for (int t = 0; t < n_; ++t) any_lin = any_lin || pd_.lin[t] != 0.0;
    if (any_lin) {
        if (nops_ + 1 > (size_t)kMaxOps || nphase_ + 1 > (size_t)kMaxPhase || (int)rounds_.size() >= kMaxRounds) { close_sweep(); open_sweep(); }
        OpB op;
        op.is_phase = true;
        op.is_lin = true;
        op.base = pd_.c0;
        for (int t = 0; t < n_; ++t)
            if (pd_.lin[t] != 0.0) op.partners.push_back(std::make_pair(t, pd_.lin[t]));
        pd_.c0 = 0.0;
        std::fill(pd_.lin.begin(), pd_.lin.end(), 0.0);
        RoundB r;                                     // register bits: any tile bit will do
        r.regs.push_back(tile_.empty() ? 0 : tile_.back());
        op.target = r.regs[0];
        rounds_.push_back(r);
        emit_op(op);
    }
    close_sweep();
    open_sweep();
}

std::vector<PlannedSweep> Planner::take()
{
    std::vector<PlannedSweep> out;
    out.swap(done_);
    return out;
}

void Planner::close_sweep()
{
    if (rounds_.empty()) return;
    PlannedSweep ps;
    SweepProgram &P = ps.prog;
    std::memset(&P, 0, sizeof P);
    const int T = T_;
    // everything below works on CURRENT positions: the planner's bookkeeping (tile_, rounds_, pending terms) stays on
    // the positions the gates arrive with, cur_ says where those are after the relabelling stores planned so far
    std::vector<RoundB> rounds_x = rounds_;
    for (RoundB &rb : rounds_x) {
        for (int &p : rb.regs) p = cur_[p];
        for (OpB &ob : rb.ops) {
            ob.target = cur_[ob.target];
            uint64_t cm = 0;
            for (int p = 0; p < n_; ++p)
                if ((ob.cmask >> p) & 1ull) cm |= 1ull << cur_[p];
            ob.cmask = cm;
            for (auto &pr : ob.partners) pr.first = cur_[pr.first];
        }
    }
    // pad the tile with the lowest unused positions (mid_relabel 2: with the upcoming targets first -- passengers that the
    // relabelling store below can move into the low positions)
    std::vector<int> tile;
    for (int p : tile_) tile.push_back(cur_[p]);
    std::vector<int> next_targets;
    if (mid_relabel_ >= 2 && !la_.empty()) {
        for (int p : upcoming((size_t)T)) next_targets.push_back(cur_[p]);
        for (size_t i = 0; i < next_targets.size() && (int)tile.size() < T; ++i)
            if (std::find(tile.begin(), tile.end(), next_targets[i]) == tile.end()) tile.push_back(next_targets[i]);
    }
    for (int p = 0; p < n_ && (int)tile.size() < T; ++p)
        if (std::find(tile.begin(), tile.end(), p) == tile.end()) tile.push_back(p);
    std::sort(tile.begin(), tile.end());
    std::vector<int> outer;
    for (int p = 0; p < n_; ++p)
        if (!std::binary_search(tile.begin(), tile.end(), p)) outer.push_back(p);
    P.n = n_; P.T = T; P.TB = T - kRegBits; P.n_outer = n_ - T;
    P.coalesce = std::min(C_, n_);
    P.relabel = 0;
    int tile_index[kMaxBits], outer_index[kMaxBits];
    for (int p = 0; p < kMaxBits; ++p) { tile_index[p] = -1; outer_index[p] = -1; }
    for (int i = 0; i < T; ++i) { P.tsrc[i] = P.tdst[i] = (uint8_t)tile[i]; tile_index[tile[i]] = i; P.st_tb[i] = (uint8_t)i; }
    for (int i = 0; i < P.n_outer; ++i) { P.osrc[i] = P.odst[i] = (uint8_t)outer[i]; outer_index[outer[i]] = i; }
    fill_outer_tables(P);
    {
        int pos[kMaxThrBits + 1];
        for (int i = 0; i < P.TB; ++i) pos[i] = tile[i];
        P.ld_nruns = make_runs(pos, P.TB, P.ld_runs);
        P.st_nruns = make_runs(pos, P.TB, P.st_runs);
        for (int i = 0; i < P.TB; ++i) pos[i] = i;
        const int nl = make_runs(pos, P.TB, P.st_lruns);     // identity: one run
        for (int k = nl; k < P.st_nruns; ++k) { P.st_lruns[k].mask = 0; P.st_lruns[k].shift = 0; }
    }
    for (int i = 0; i < kSlots; ++i) {
        uint64_t off = 0;
        for (int b = 0; b < kRegBits; ++b)
            if ((i >> b) & 1) off |= 1ull << tile[P.TB + b];
        P.ld_hi[i] = off;
        P.ld_sw_hi[i] = tile_swizzle((uint32_t)i << P.TB) * 16u;
        P.st_off_hi[i] = off;
        P.st_l_hi[i] = tile_swizzle((uint32_t)i << P.TB) * 16u;
    }
    P.nrounds = (int)rounds_.size();
    int nops = 0, nphase = 0;
    // force_lane: tile bits that must be thread bits 0..C-1 of the LAST round, in this order (the sources of the low
    // destination positions of a relabelling store, so that it can stay a direct store); null: the default order
    auto build_rounds = [&](const int *force_lane) {
    nops = 0; nphase = 0;
    P.scale = 1.0;
    ps.ptabs.clear();
    ps.touched = 0;
    for (int r = 0; r < P.nrounds; ++r) {
        RoundB &rb = rounds_x[r];
        RoundDesc &R = P.rounds[r];
        // register bits as tile-bit indices.  Slot bits kRegBits-cnt.. are the round's targets in
        // placement order, so that a chain of (phase, Hadamard) steps is a ladder whatever the
        // positions of its bits; the low slots are padding (the highest free tile bits), which a
        // ladder step treats like already-processed partner bits
        std::vector<int> regs;
        for (int p : rb.regs) regs.push_back(tile_index[p]);
        {
            std::vector<int> pads;
            for (int tb = T - 1; tb >= 0 && (int)(regs.size() + pads.size()) < kRegBits; --tb)
                if (std::find(regs.begin(), regs.end(), tb) == regs.end()) pads.push_back(tb);
            regs.insert(regs.begin(), pads.begin(), pads.end());
        }
        int slot_of_tb[kMaxTileBits + 3];
        for (int tb = 0; tb < T; ++tb) slot_of_tb[tb] = -1;
        for (int j = 0; j < kRegBits; ++j) { R.reg_tb[j] = (uint8_t)regs[j]; slot_of_tb[regs[j]] = j; }
        // thread bits: the remaining tile bits (tid bit i -> tile bit thr[i]).  The eight lanes of a
        // quarter warp make one 128-byte shared-memory wavefront; tile_swizzle() folds tile bit b
        // into bank-group bit b % 3, so tid bits 0..2 take the lowest tile bits with three different
        // residues (conflict-free LDS/STS.128); the others follow in ascending order
        std::vector<int> ordered;
        {
            std::vector<int> rest;
            for (int tb = 0; tb < T; ++tb)
                if (slot_of_tb[tb] < 0) rest.push_back(tb);
            bool used[3] = { false, false, false };
            if (force_lane && r == P.nrounds - 1) {
                bool ok = true;
                bool res[3] = { false, false, false };
                for (int i = 0; i < P.coalesce; ++i) {
                    ok = ok && slot_of_tb[force_lane[i]] < 0 && !res[force_lane[i] % 3];
                    res[force_lane[i] % 3] = true;
                }
                if (ok)
                    for (int i = 0; i < P.coalesce; ++i) {
                        ordered.push_back(force_lane[i]);
                        used[force_lane[i] % 3] = true;
                        rest.erase(std::find(rest.begin(), rest.end(), force_lane[i]));
                    }
            }
            for (size_t i = 0; i < rest.size() && ordered.size() < 3;) {
                if (!used[rest[i] % 3]) {
                    used[rest[i] % 3] = true;
                    ordered.push_back(rest[i]);
                    rest.erase(rest.begin() + i);
                } else ++i;
            }
            ordered.insert(ordered.end(), rest.begin(), rest.end());
        }
        int thr_index_of_tb[kMaxTileBits + 3];
        for (int tb = 0; tb < T; ++tb) thr_index_of_tb[tb] = -1;
        for (int i = 0; i < P.TB; ++i) { R.thr_tb[i] = (uint8_t)ordered[i]; thr_index_of_tb[ordered[i]] = i; }
        R.nruns = (uint8_t)make_runs(ordered.data(), P.TB, R.runs);
        for (int s = 0; s < kSlots; ++s) {
            uint32_t l = 0;
            for (int j = 0; j < kRegBits; ++j)
                if ((s >> j) & 1) l |= 1u << regs[j];
            R.sw_slot[s] = tile_swizzle(l) * 16u;
        }
        R.op_begin = (uint16_t)nops;
        for (size_t oi = 0; oi < rb.ops.size(); ++oi) {
            const OpB &ob = rb.ops[oi];
            OpDesc &op = P.ops[nops];
            std::memset(&op, 0, sizeof op);
            const int j = slot_of_tb[tile_index[ob.target]];
            op.j = (uint8_t)j;
            if (!ob.is_phase) ps.touched |= 1ull << ob.target;
            const bool plain_h = !ob.is_phase && ob.kind == OP_G1_HADAMARD && ob.cmask == 0;
            if (plain_h) {
                // uncontrolled Hadamard: butterfly only, c = 1/sqrt(2) deferred to the store
                op.kind = OP_H_UNNORM;
                P.scale *= ob.m[0];
            } else if (!ob.is_phase) {
                op.kind = (uint8_t)ob.kind;
                std::memcpy(op.m, ob.m, sizeof ob.m);
                for (int p = 0; p < n_; ++p) {
                    if (!((ob.cmask >> p) & 1ull)) continue;
                    if (tile_index[p] >= 0) {
                        const int tb = tile_index[p];
                        if (slot_of_tb[tb] >= 0) op.cslot |= 1u << slot_of_tb[tb];
                        else op.cmask |= 1ull << tb;
                    } else op.cmask |= 1ull << (T + outer_index[p]);
                }
            } else if (ob.is_lin) {
                op.kind = OP_LINPHASE;
                op.phase_id = (uint32_t)nphase;
                PhaseTab pt;
                std::memset(&pt, 0, sizeof pt);
                double lo_ang[1 << kThrLoBits] = { 0 }, hi_ang[1 << (kMaxThrBits - kThrLoBits)] = { 0 };
                pt.base = wrap_half_turns(ob.base);
                for (int q = 0; q < kRegBits; ++q) { op.m[2 * q] = 1.0; op.m[2 * q + 1] = 0.0; }
                for (const auto &pr : ob.partners) {
                    const int p = pr.first;
                    const double v = pr.second;
                    if (tile_index[p] >= 0) {
                        const int tb = tile_index[p];
                        if (slot_of_tb[tb] >= 0) sincospi_host(v, op.m[2 * slot_of_tb[tb] + 1], op.m[2 * slot_of_tb[tb]]);
                        else {
                            const int ti = thr_index_of_tb[tb];
                            if (ti < kThrLoBits) {
                                for (int x = 0; x < (1 << kThrLoBits); ++x)
                                    if ((x >> ti) & 1) lo_ang[x] += v;
                            } else {
                                for (int x = 0; x < (1 << (kMaxThrBits - kThrLoBits)); ++x)
                                    if ((x >> (ti - kThrLoBits)) & 1) hi_ang[x] += v;
                            }
                        }
                    } else pt.outer_coef[outer_index[p]] = v;
                }
                for (int x = 0; x < (1 << kThrLoBits); ++x) sincospi_host(lo_ang[x], pt.lo[2 * x + 1], pt.lo[2 * x]);
                for (int x = 0; x < (1 << (kMaxThrBits - kThrLoBits)); ++x) sincospi_host(hi_ang[x], pt.hi[2 * x + 1], pt.hi[2 * x]);
                ps.ptabs.push_back(pt);
                ++nphase;
            } else {
                op.kind = OP_PHASE;
                if (!ob.has_c0 && oi + 1 < rb.ops.size()) {
                    const OpB &nx = rb.ops[oi + 1];
                    if (!nx.is_phase && nx.kind == OP_G1_HADAMARD && nx.cmask == 0 && nx.target == ob.target) {
                        op.kind = OP_PHASE_H;       // phase then butterfly on the same bit, fused
                        P.scale *= nx.m[0];
                        ps.touched |= 1ull << nx.target;
                        ++oi;
                    }
                }
                op.phase_id = (uint32_t)nphase;
                PhaseTab pt;
                std::memset(&pt, 0, sizeof pt);
                double lo_ang[1 << kThrLoBits] = { 0 }, hi_ang[1 << (kMaxThrBits - kThrLoBits)] = { 0 };
                pt.base = wrap_half_turns(ob.base + (ob.has_c0 ? ob.c0 : 0.0));
                if (ob.has_c0) {
                    op.flags |= kFlagC0;
                    sincospi_host(ob.c0, op.m[9], op.m[8]);
                }
                for (const auto &pr : ob.partners) {
                    const int p = pr.first;
                    const double v = pr.second;
                    if (tile_index[p] >= 0) {
                        const int tb = tile_index[p];
                        if (slot_of_tb[tb] >= 0) {
                            int js = slot_of_tb[tb];
                            const int qi = js < j ? js : js - 1;      // index among the other slot bits
                            op.flags |= 1u << qi;
                            sincospi_host(v, op.m[2 * qi + 1], op.m[2 * qi]);
                        } else {
                            const int ti = thr_index_of_tb[tb];
                            if (ti < kThrLoBits) {
                                for (int x = 0; x < (1 << kThrLoBits); ++x)
                                    if ((x >> ti) & 1) lo_ang[x] += v;
                            } else {
                                for (int x = 0; x < (1 << (kMaxThrBits - kThrLoBits)); ++x)
                                    if ((x >> (ti - kThrLoBits)) & 1) hi_ang[x] += v;
                            }
                        }
                    } else pt.outer_coef[outer_index[p]] = v;
                }
                for (int x = 0; x < (1 << kThrLoBits); ++x) sincospi_host(lo_ang[x], pt.lo[2 * x + 1], pt.lo[2 * x]);
                for (int x = 0; x < (1 << (kMaxThrBits - kThrLoBits)); ++x) sincospi_host(hi_ang[x], pt.hi[2 * x + 1], pt.hi[2 * x]);
                ps.ptabs.push_back(pt);
                ++nphase;
            }
            ++nops;
        }
        R.op_end = (uint16_t)nops;
        // ROUND_PH: the round is a ladder of fused (phase, butterfly) steps on slot bits 0,1,2,.. in
        // order, each phase only coupling to slot bits already processed -> straight-line kernel path
        R.kind = ROUND_GENERIC;
        R.nsteps = 0;
        const int cnt = nops - R.op_begin;
        const int j0 = kRegBits - cnt;                       // first slot bit of the ladder
        bool ladder = cnt >= 1 && cnt <= kRegBits;
        for (int i = 0; ladder && i < cnt; ++i) {
            const OpDesc &op = P.ops[R.op_begin + i];
            if (op.j != j0 + i) ladder = false;
            else if (op.kind == OP_H_UNNORM) continue;
            else if (op.kind != OP_PHASE_H) ladder = false;
            else if (op.flags >> op.j) ladder = false;       // partner among later slot bits, or a c0 term
        }
        if (ladder && nphase + cnt <= kMaxPhase) {
            for (int i = 0; i < cnt; ++i) {
                OpDesc &op = P.ops[R.op_begin + i];
                if (op.kind == OP_H_UNNORM) {                // bare butterfly = phase step with unit factors
                    PhaseTab pt;
                    std::memset(&pt, 0, sizeof pt);
                    for (int x = 0; x < (1 << kThrLoBits); ++x) pt.lo[2 * x] = 1.0;
                    for (int x = 0; x < (1 << (kMaxThrBits - kThrLoBits)); ++x) pt.hi[2 * x] = 1.0;
                    op.kind = OP_PHASE_H;
                    op.phase_id = (uint32_t)nphase;
                    op.flags = 0;
                    ps.ptabs.push_back(pt);
                    ++nphase;
                }
                for (int q = 0; q < op.j; ++q)
                    if (!(op.flags & (1u << q))) { op.m[2 * q] = 1.0; op.m[2 * q + 1] = 0.0; }
            }
            R.kind = ROUND_PH;
            R.nsteps = (uint8_t)cnt;
        }
    }
    };
    build_rounds(nullptr);
    P.nops = nops;
    P.nphase = nphase;
    setup_direct(P);
    if (mid_relabel_ && n_ > T) {
        bool ladder = P.nrounds > 0, contiguous = true;
        for (int r = 0; r < P.nrounds; ++r) ladder = ladder && P.rounds[r].kind == ROUND_PH;
        for (int i = 0; i < T; ++i) contiguous = contiguous && tile[i] == i;
        // mid_relabel 2: the next targets that ride in this tile go to the low C positions
        std::vector<int> newlow;
        if (ladder)
            for (size_t i = 0; i < next_targets.size() && (int)newlow.size() < C_; ++i)
                if (std::binary_search(tile.begin(), tile.end(), next_targets[i])) newlow.push_back(next_targets[i]);
        bool rotate = false;
        for (int p : newlow) rotate = rotate || p >= C_;
        if (ladder && (!contiguous || rotate)) {
            // destination of every tile bit: positions 0..T-1; newlow first, then whoever sits in a position it can keep,
            // then the rest into the holes; the non-tile bits below T move up into the vacated positions
            std::vector<int> dstpos(n_, -1);
            std::vector<char> taken(n_, 0);
            int at = 0;
            for (int p : newlow) { dstpos[p] = at; taken[at] = 1; ++at; }
            // low positions not claimed by a next target keep a low occupant: one whose tile-bit index has a residue mod 3
            // that the lanes so far do not have (tile_swizzle: the three lane bits of a round want three residues)
            {
                bool res[3] = { false, false, false };
                for (int p : newlow) res[(int)(std::lower_bound(tile.begin(), tile.end(), p) - tile.begin()) % 3] = true;
                for (int pass = 0; pass < 2; ++pass)
                    for (int lowpos = 0; lowpos < C_ && lowpos < n_; ++lowpos) {
                        if (taken[lowpos]) continue;
                        // candidates: low occupants without a destination yet; pass 0 wants a new residue and its own place
                        int pick = -1;
                        for (int b = 0; b < C_ && b < n_; ++b) {
                            if (dstpos[b] >= 0) continue;
                            const int rb3 = b % 3;                        // (low bits are tile bits 0..C-1: index = position)
                            if (pass == 0 && (res[rb3] || b != lowpos)) continue;
                            if (pass == 1 && pick >= 0 && !res[rb3]) { pick = b; break; }
                            if (pick < 0) pick = b;
                            if (!res[rb3]) break;
                        }
                        if (pick >= 0) { dstpos[pick] = lowpos; taken[lowpos] = 1; res[pick % 3] = true; }
                    }
            }
            for (int i = 0; i < T; ++i) {                                        // tile bits already inside [C, T): stay
                const int p = tile[i];
                if (dstpos[p] < 0 && p >= C_ && p < T && !taken[p]) { dstpos[p] = p; taken[p] = 1; }
            }
            for (int i = 0; i < T; ++i) {                                        // the other tile bits: lowest free position below T
                const int p = tile[i];
                if (dstpos[p] >= 0) continue;
                int q = 0;
                while (q < T && taken[q]) ++q;
                dstpos[p] = q; taken[q] = 1;
            }
            std::vector<int> vacated;                                            // positions >= T that tile bits have left
            for (int i = 0; i < T; ++i)
                if (tile[i] >= T) vacated.push_back(tile[i]);
            size_t v = 0;
            for (int p = 0; p < n_; ++p) {
                if (dstpos[p] >= 0) continue;
                if (p < T) dstpos[p] = vacated[v++];                             // a non-tile bit below T: up
                else dstpos[p] = p;
            }
            bool ident = true;
            for (int p = 0; p < n_; ++p) ident = ident && dstpos[p] == p;
            if (!ident && can_fuse_relabel(P, dstpos)) {
                // the last round's lanes become the sources of the low destination positions, so that the relabelling
                // store stays a direct store (setup_direct) instead of one more trip through shared memory
                int force[3] = { -1, -1, -1 };
                bool have = true;
                for (int i = 0; i < P.coalesce; ++i) {
                    for (int tb = 0; tb < T; ++tb)
                        if (dstpos[tile[tb]] == i) force[i] = tb;
                    have = have && force[i] >= 0;
                }
                bool differs = false;
                for (int i = 0; i < P.coalesce; ++i) differs = differs || force[i] != i;
                if (have && differs && mid_relabel_ >= 2) {
                    build_rounds(force);
                    P.nops = nops;
                    P.nphase = nphase;
                    setup_direct(P);
                }
                ps.mid_dstpos = dstpos;
                for (int p = 0; p < n_; ++p) cur_[p] = dstpos[cur_[p]];
                for (int p = 0; p < n_; ++p) inv_[cur_[p]] = p;
            }
        }
    }
    stats.sweeps += 1;
    stats.rounds += P.nrounds;
    stats.ops += nops;
    done_.push_back(ps);
    rounds_.clear();
}

// ---------------------------------------------------------------------------
// relabelling: source bit p goes to destination position dstpos[p]
// ---------------------------------------------------------------------------
// Fill the destination side of a program whose source side (tsrc, osrc, T, TB)
// is set.  The store is coalesced when the tile contains the sources of the low
// destination bits.
void set_relabel(SweepProgram &P, const std::vector<int> &dstpos, bool leaf_split)
{
    const int T = P.T;
    bool ident = true;
    for (int i = 0; i < T; ++i) { P.tdst[i] = (uint8_t)dstpos[P.tsrc[i]]; ident = ident && P.tdst[i] == P.tsrc[i]; }
    for (int i = 0; i < P.n_outer; ++i) { P.odst[i] = (uint8_t)dstpos[P.osrc[i]]; ident = ident && P.odst[i] == P.osrc[i]; }
    P.relabel = ident ? 0 : 1;
    std::vector<int> order(T);
    for (int i = 0; i < T; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return P.tdst[a] < P.tdst[b]; });
    for (int i = 0; i < T; ++i) P.st_tb[i] = (uint8_t)order[i];
    P.leaf_fuse = 0;
    if (leaf_split && T == 12 && P.TB == 7 && kRegBits == 5) {
        bool whole_leaves = true;
        for (int i = 0; i < 10; ++i) whole_leaves = whole_leaves && P.tdst[order[i]] == i;
        if (whole_leaves) {
            // thread bits: destination bits 0..4 (lanes) and the two tile bits above the leaf (warps);
            // slot bits: destination bits 5..9 -> slot i of lane l is element l + 32 i of the warp's leaf
            const int sel[12] = { 0, 1, 2, 3, 4, 10, 11, 5, 6, 7, 8, 9 };
            std::vector<int> o2(T);
            for (int i = 0; i < T; ++i) o2[i] = order[sel[i]];
            order = o2;
            P.leaf_fuse = 1;
        }
    }
    int pos[kMaxThrBits + 1];
    for (int i = 0; i < P.TB; ++i) pos[i] = P.tdst[order[i]];
    P.st_nruns = make_runs(pos, P.TB, P.st_runs);
    // tile index of the low part of f: tid bit m -> tile bit order[m]; split at the same places as st_runs
    // (each run is re-split until both mappings are affine inside it)
    {
        BitRun a[kMaxRuns * 2], b[kMaxRuns * 2];
        int n2 = 0;
        int i = 0;
        while (i < P.TB) {
            int j = i;
            uint32_t mask = 0;
            while (j < P.TB && pos[j] - j == pos[i] - i && order[j] - j == order[i] - i) { mask |= 1u << j; ++j; }
            a[n2].mask = mask; a[n2].shift = pos[i] - i;
            b[n2].mask = mask; b[n2].shift = order[i] - i;
            ++n2;
            i = j;
        }
        P.st_nruns = n2;
        for (int k = 0; k < n2 && k < kMaxRuns; ++k) { P.st_runs[k] = a[k]; P.st_lruns[k] = b[k]; }
    }
    for (int i = 0; i < kSlots; ++i) {
        uint64_t dof = 0;
        uint32_t l = 0;
        for (int b = 0; b < kRegBits; ++b) {
            if (!((i >> b) & 1)) continue;
            const int tb = order[P.TB + b];
            dof |= 1ull << P.tdst[tb];
            l |= 1u << tb;
        }
        P.st_off_hi[i] = dof;
        P.st_l_hi[i] = tile_swizzle(l) * 16u;
    }
    fill_outer_tables(P);
    setup_direct(P);
}

// can a relabel be fused into this (gate) sweep?  the sources of the three lowest
// destination bits must be tile bits so that stores stay 128-byte coalesced
bool can_fuse_relabel(const SweepProgram &P, const std::vector<int> &dstpos)
{
    int found = 0;
    for (int i = 0; i < P.T; ++i)
        if (dstpos[P.tsrc[i]] < P.coalesce) ++found;
    return found >= (P.n < P.coalesce ? P.n : P.coalesce);
}

// ---------------------------------------------------------------------------
// TMA tile layout (ladder kernel, dense sweeps).
//
// A cp.async.bulk.tensor load delivers the 512 lines (128 B each) of a tile in the order of the tensor
// map's box dimensions and, with CU_TENSOR_MAP_SWIZZLE_128B, xors the 16-byte chunk index of every
// line with the three lowest line-index bits.  That is a narrower swizzle than tile_swizzle() (which
// folds ALL line bits into the chunk bits): a quarter warp's LDS/STS.128 is conflict-free only if its
// three lane bits land on smem bits {0..5} with three different residues mod 3.  The line order is ours
// to choose (any order of the box dimensions), so: pick the three tile bits that become smem bits 3, 4, 5
// such that the lane bits every round (and the staged store pass) already uses qualify -- the thread
// maps of the rounds, hence the phase tables, stay as they are; only the shared-memory offset tables are
// rewritten.  Returns false (program untouched) when no such choice exists or the line order needs more
// box dimensions / requests than the kernel supports.
// ---------------------------------------------------------------------------
bool apply_tma_layout(SweepProgram &P)
{
    const int T = P.T, TB = P.TB;
    if (T != 12 || TB != 7 || kRegBits != 5 || P.coalesce != 3 || P.n < T + 1) return false;
    for (int i = 0; i < 3; ++i)
        if (P.tsrc[i] != i) return false;
    // lane triples that must be conflict-free
    std::vector<std::array<int, 3>> triples;
    for (int r = 0; r < P.nrounds; ++r) triples.push_back({ P.rounds[r].thr_tb[0], P.rounds[r].thr_tb[1], P.rounds[r].thr_tb[2] });
    // store pass: tid bit m -> tile bit st_thr[m], slot bit b -> tile bit st_reg[b] (recovered from the tables)
    int st_thr[kMaxThrBits], st_reg[kRegBits], st_dpos[kMaxThrBits];
    auto one_bit = [](uint64_t v) { int b = 0; while (b < 63 && !((v >> b) & 1ull)) ++b; return b; };
    for (int m = 0; m < TB; ++m) {
        uint64_t l = 0, d = 0;
        for (int k = 0; k < P.st_nruns; ++k) {
            const uint32_t v = (1u << m) & P.st_lruns[k].mask;
            const int sh = P.st_lruns[k].shift;
            l |= sh >= 0 ? (uint64_t)v << sh : (uint64_t)v >> -sh;
            d |= (uint64_t)((1u << m) & P.st_runs[k].mask) << P.st_runs[k].shift;
        }
        if (__builtin_popcountll(l) != 1 || __builtin_popcountll(d) != 1) return false;
        st_thr[m] = one_bit(l);
        st_dpos[m] = one_bit(d);
    }
    for (int b = 0; b < kRegBits; ++b) {
        const uint32_t l = tile_swizzle(P.st_l_hi[1 << b] >> 4);     // the swizzle is an involution
        if (__builtin_popcount(l) != 1) return false;
        st_reg[b] = one_bit(l);
    }
    const bool staged_store = P.direct_store == 0 || P.nrounds == 0;
    if (staged_store) triples.push_back({ st_thr[0], st_thr[1], st_thr[2] });
    int best[3] = { -1, -1, -1 }, best_cost = 1 << 30;
    int best_pi[kMaxTileBits + 3];
    struct Dims { int nruns; int run_len[16]; int run_pos[16]; };
    Dims best_dims;
    std::memset(&best_dims, 0, sizeof best_dims);
    for (int s3 = 3; s3 < T; ++s3)
        for (int s4 = 3; s4 < T; ++s4)
            for (int s5 = 3; s5 < T; ++s5) {
                if (s3 == s4 || s3 == s5 || s4 == s5) continue;
                int pi[kMaxTileBits + 3];
                for (int i = 0; i < T; ++i) pi[i] = -1;
                pi[0] = 0; pi[1] = 1; pi[2] = 2; pi[s3] = 3; pi[s4] = 4; pi[s5] = 5;
                int next = 6;
                for (int i = 3; i < T; ++i)
                    if (pi[i] < 0) pi[i] = next++;
                bool ok = true;
                for (const auto &t : triples) {
                    bool res[3] = { false, false, false };
                    for (int k = 0; k < 3; ++k) {
                        const int m = pi[t[k]];
                        if (m > 5 || res[m % 3]) { ok = false; break; }
                        res[m % 3] = true;
                    }
                    if (!ok) break;
                }
                if (!ok) continue;
                // line order -> runs of consecutive source bits (at most 8 bits = 256 lines per box dimension)
                int tb_of_m[kMaxTileBits + 3];
                for (int i = 0; i < T; ++i) tb_of_m[pi[i]] = i;
                Dims d;
                d.nruns = 0;
                for (int m = 3; m < T;) {
                    int len = 1;
                    while (m + len < T && len < 8 && P.tsrc[tb_of_m[m + len]] == P.tsrc[tb_of_m[m]] + len) ++len;
                    d.run_len[d.nruns] = len;
                    d.run_pos[d.nruns] = P.tsrc[tb_of_m[m]];
                    ++d.nruns;
                    m += len;
                }
                int extra_bits = 0;
                for (int k = 3; k < d.nruns; ++k) extra_bits += d.run_len[k];
                if ((1 << extra_bits) > kMaxTmaReq) continue;
                const int cost = (extra_bits << 8) + d.nruns;
                if (cost < best_cost) {
                    best_cost = cost;
                    best[0] = s3; best[1] = s4; best[2] = s5;
                    std::memcpy(best_pi, pi, sizeof pi);
                    best_dims = d;
                }
            }
    if (best[0] < 0) return false;
    const int *pi = best_pi;
    const Dims &d = best_dims;
    // tensor description
    int extra_bits = 0;
    for (int k = 3; k < d.nruns; ++k) extra_bits += d.run_len[k];
    P.tma_nreq = 1 << extra_bits;
    P.tma_req_bytes = (uint32_t)((sizeof(double) * 2) << (T - extra_bits));
    P.tma_box[0] = 16; P.tma_gdim[0] = 16;
    for (int k = 0; k < 3; ++k) {
        if (k < d.nruns) {
            P.tma_box[1 + k] = 1u << d.run_len[k];
            P.tma_gdim[1 + k] = 1ull << d.run_len[k];
            P.tma_gstride[k] = 16ull << d.run_pos[k];
        } else {
            P.tma_box[1 + k] = 1;
            P.tma_gdim[1 + k] = 1;
            P.tma_gstride[k] = 128;
        }
    }
    P.tma_box[4] = 1;
    P.tma_gdim[4] = 1ull << (P.n - 3);
    P.tma_gstride[3] = 128;
    {
        // the bits iterated by separate requests are the top smem bits, in run order
        int xpos[16], nx = 0;
        for (int k = 3; k < d.nruns; ++k)
            for (int j = 0; j < d.run_len[k]; ++j) xpos[nx++] = d.run_pos[k] + j;
        for (int q = 0; q < P.tma_nreq; ++q) {
            uint64_t line = 0;
            for (int j = 0; j < nx; ++j)
                if ((q >> j) & 1) line |= 1ull << (xpos[j] - 3);
            P.tma_req_line[q] = line;
        }
    }
    for (int i = 0; i < T; ++i) P.tma_pi[i] = (uint8_t)pi[i];
    // shared-memory offset tables in the new order
    for (int r = 0; r < P.nrounds; ++r) {
        RoundDesc &R = P.rounds[r];
        int pos[kMaxThrBits + 1];
        for (int i = 0; i < TB; ++i) pos[i] = pi[R.thr_tb[i]];
        R.nruns = (uint8_t)make_runs(pos, TB, R.runs);
        for (int s = 0; s < kSlots; ++s) {
            uint32_t m = 0;
            for (int j = 0; j < kRegBits; ++j)
                if ((s >> j) & 1) m |= 1u << pi[R.reg_tb[j]];
            R.sw_slot[s] = tma_swizzle(m) * 16u;
        }
    }
    {
        int n2 = 0, i = 0;
        BitRun a[kMaxRuns * 2], b[kMaxRuns * 2];
        while (i < TB) {
            int j = i;
            uint32_t mask = 0;
            while (j < TB && st_dpos[j] - j == st_dpos[i] - i && pi[st_thr[j]] - j == pi[st_thr[i]] - i) { mask |= 1u << j; ++j; }
            a[n2].mask = mask; a[n2].shift = st_dpos[i] - i;
            b[n2].mask = mask; b[n2].shift = pi[st_thr[i]] - i;
            ++n2;
            i = j;
        }
        if (n2 > kMaxRuns) { P.tma_nreq = 0; return false; }
        P.st_nruns = n2;
        for (int k = 0; k < n2; ++k) { P.st_runs[k] = a[k]; P.st_lruns[k] = b[k]; }
        for (int s = 0; s < kSlots; ++s) {
            uint32_t m = 0;
            for (int bb = 0; bb < kRegBits; ++bb)
                if ((s >> bb) & 1) m |= 1u << pi[st_reg[bb]];
            P.st_l_hi[s] = tma_swizzle(m) * 16u;
        }
    }
    return true;
}

// ---------------------------------------------------------------------------
// in-place relabelling.  A relabel sweep may run with source == destination when its permutation is
// TILE-CLOSED: it maps the tile's bit positions onto themselves and fixes every outer bit, so a CTA
// writes exactly the addresses it has read (and it has read all of them before its first store).
// Any bit permutation is a product of such passes: each pass takes the low `coalesce` bits (kept
// for 128-byte accesses, moved or not) plus as many whole cycles of the remaining permutation as fit
// into the tile; of a cycle that does not fit it takes a segment a1..am, puts a1..a(m-1) into their
// final places and parks am at a1, which shortens the cycle by m-1.  The bit reversal of QFT-30
// takes 4 passes at 12 tile bits.  Used when no second column buffer fits into device memory
// (128 GiB shards); otherwise the relabel is one out-of-place sweep.
// ---------------------------------------------------------------------------
std::vector<InplacePass> plan_inplace_relabel(int n, int tile_bits, int coalesce, const std::vector<int> &dstpos)
{
    std::vector<InplacePass> passes;
    const int T = std::min(tile_bits, n), c = std::min(coalesce, n);
    std::vector<int> R(dstpos);                  // data now at position p still has to go to R[p]
    for (;;) {
        std::vector<char> seen(n, 0), in_tile(n, 0);
        std::vector<std::vector<int>> cycles;
        for (int p = 0; p < n; ++p) {
            if (seen[p] || R[p] == p) continue;
            std::vector<int> cyc;
            for (int q = p; !seen[q]; q = R[q]) { seen[q] = 1; cyc.push_back(q); }
            cycles.push_back(cyc);
        }
        if (cycles.empty()) break;
        auto low_members = [&](const std::vector<int> &cy) { int k = 0; for (int q : cy) k += q < c; return k; };
        std::stable_sort(cycles.begin(), cycles.end(), [&](const std::vector<int> &a, const std::vector<int> &b) {
            const int la = low_members(a), lb = low_members(b);
            if ((la > 0) != (lb > 0)) return la > 0;             // cycles through the low bits first: they cost less
            return a.size() - la < b.size() - lb;                // then the cheapest
        });
        InplacePass ps;
        ps.dstpos.resize(n);
        for (int p = 0; p < n; ++p) ps.dstpos[p] = p;
        for (int b = 0; b < c; ++b) in_tile[b] = 1;
        int freeb = T - c;
        bool progress = false;
        for (const std::vector<int> &cy : cycles) {
            const int L = (int)cy.size();
            const int cost = L - low_members(cy);
            if (cost <= freeb) {                                 // the whole cycle
                for (int q : cy) { ps.dstpos[q] = R[q]; in_tile[q] = 1; }
                freeb -= cost;
                progress = true;
                continue;
            }
            // longest affordable segment: try every start, keep the longest (a start at a low bit is cheaper)
            int best_start = -1, best_len = 0;
            for (int s0 = 0; s0 < L; ++s0) {
                int len = 0, spent = 0;
                while (len < L - 1) {
                    const int q = cy[(s0 + len) % L];
                    const int add = q < c ? 0 : 1;
                    if (spent + add > freeb) break;
                    spent += add;
                    ++len;
                }
                if (len > best_len) { best_len = len; best_start = s0; }
            }
            if (best_len < 2) continue;
            for (int i = 0; i < best_len; ++i) {
                const int q = cy[(best_start + i) % L];
                if (q >= c) --freeb;
                in_tile[q] = 1;
                ps.dstpos[q] = i + 1 < best_len ? R[q] : cy[best_start];      // the last one is parked at the segment's head
            }
            progress = true;
        }
        if (!progress) { passes.clear(); return passes; }       // cannot happen for T >= coalesce + 2
        // fill the tile with the lowest free bits (fixed passengers: longer contiguous runs)
        for (int p = 0; p < n && freeb > 0; ++p)
            if (!in_tile[p]) { in_tile[p] = 1; --freeb; }
        for (int p = 0; p < n; ++p)
            if (in_tile[p]) ps.tile.push_back(p);
        // what is left: the data now at ps.dstpos[p] still has to go to R[p]
        std::vector<int> R2(n);
        for (int p = 0; p < n; ++p) R2[ps.dstpos[p]] = R[p];
        R.swap(R2);
        passes.push_back(ps);
    }
    return passes;
}

PlannedSweep build_permute_sweep(int n, int tile_bits, const std::vector<int> &dstpos, const std::vector<int> *forced_tile)
{
    PlannedSweep ps;
    ps.is_permute = true;
    SweepProgram &P = ps.prog;
    std::memset(&P, 0, sizeof P);
    const int T = forced_tile ? (int)forced_tile->size() : std::min(tile_bits, n);
    std::vector<int> srcpos_of_dst(n);
    for (int p = 0; p < n; ++p) srcpos_of_dst[dstpos[p]] = p;
    // tile = low source bits + sources of the low destination bits, alternating until full
    std::vector<int> tile;
    if (forced_tile) tile = *forced_tile;
    for (int i = 0; i < n && (int)tile.size() < T; ++i) {
        if (std::find(tile.begin(), tile.end(), i) == tile.end()) tile.push_back(i);
        if ((int)tile.size() >= T) break;
        const int s = srcpos_of_dst[i];
        if (std::find(tile.begin(), tile.end(), s) == tile.end()) tile.push_back(s);
    }
    std::sort(tile.begin(), tile.end());
    std::vector<int> outer;
    for (int p = 0; p < n; ++p)
        if (!std::binary_search(tile.begin(), tile.end(), p)) outer.push_back(p);
    P.n = n; P.T = T; P.TB = T - kRegBits; P.n_outer = n - T;
    P.coalesce = std::min(3, n);
    P.scale = 1.0;
    for (int i = 0; i < T; ++i) P.tsrc[i] = (uint8_t)tile[i];
    for (int i = 0; i < P.n_outer; ++i) P.osrc[i] = (uint8_t)outer[i];
    {
        int pos[kMaxThrBits + 1];
        for (int i = 0; i < P.TB; ++i) pos[i] = tile[i];
        P.ld_nruns = make_runs(pos, P.TB, P.ld_runs);
    }
    for (int i = 0; i < kSlots; ++i) {
        uint64_t so = 0;
        for (int b = 0; b < kRegBits; ++b)
            if ((i >> b) & 1) so |= 1ull << tile[P.TB + b];
        P.ld_hi[i] = so;
        P.ld_sw_hi[i] = tile_swizzle((uint32_t)i << P.TB) * 16u;
    }
    set_relabel(P, dstpos);
    P.nrounds = 0; P.nops = 0; P.nphase = 0;
    return ps;
}

}  // namespace q1t
