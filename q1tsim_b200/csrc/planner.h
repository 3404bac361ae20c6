// planner.h -- host-side lowering of (matrix, bits) gates and the fusion planner
// that packs them into sweep programs.  Pure host C++ (no CUDA calls) so that it
// is unit-testable without a device (q1t_plan_dry_run).
#pragma once
#include <complex>
#include <cstdint>
#include <string>
#include <vector>

#include "program.h"

namespace q1t {

typedef std::complex<double> cplx;

// A gate after lowering, expressed on PHYSICAL index-bit positions.
struct LoweredGate {
    enum Kind { POLY, G1, SWAP, GENERIC } kind = GENERIC;
    // POLY: diagonal unitary on <= 2 bits as a phase polynomial in half-turns:
    //   angle(x)/pi = c0 + lin0*x_b0 + lin1*x_b1 + quad*x_b0*x_b1
    int nb = 0;
    int b[2] = { 0, 0 };
    double c0 = 0, lin[2] = { 0, 0 }, quad = 0;
    // G1: 2x2 matrix on `target`, applied where all bits of cmask are 1
    int target = 0;
    uint64_t cmask = 0;
    double m[8] = { 0 };
    // SWAP: exchange positions b[0], b[1] (handled as a relabel by the engine)
    // GENERIC: dense matrix on pos[] (gate-index MSB first) with controls cmask
    std::vector<int> pos;
    std::vector<cplx> mat;
};

// (matrix, logical index bits MSB-first-of-gate-index) -> lowered form.
// `phys[j]` is the physical position of gate bit j.  Returns false if the gate
// cannot be represented (too many dense targets).
bool lower_gate(const cplx *mat, int k, const int *phys, LoweredGate &out, std::string &err);

struct PlannedSweep {
    SweepProgram prog;
    std::vector<PhaseTab> ptabs;
    bool is_permute = false;
    uint64_t touched = 0;        // physical bits acted on by non-diagonal gates (these lose a pinned basis value)
    // Planner(.., mid_relabel): this sweep stores its tile relabelled -- the data at position p goes to mid_dstpos[p] --
    // and every LATER sweep of the plan is expressed in the layout that leaves.  The executor applies
    // set_relabel(prog, mid_dstpos), runs the sweep out of place and composes its qubit map with mid_dstpos; on the LAST
    // sweep of a batch it is ignored (nothing depends on it).  Empty: in place, as planned.
    std::vector<int> mid_dstpos;
};

struct PlanStats {
    uint64_t sweeps = 0, rounds = 0, ops = 0;
};

// Accumulates diagonal (phase-polynomial) terms that have not been applied yet.
struct PendingDiag {
    int n = 0;
    double c0 = 0;
    std::vector<double> lin;     // n
    std::vector<double> quad;    // n*n symmetric
    void init(int nbits) { n = nbits; c0 = 0; lin.assign(n, 0.0); quad.assign((size_t)n * n, 0.0); }
    double &q(int a, int b) { return quad[(size_t)a * n + b]; }
    void add_quad(int a, int b, double v) { q(a, b) += v; q(b, a) += v; }
    bool touches(int t) const;
    bool empty() const;
};

class Planner {
public:
    // coalesce_bits: low index bits every tile keeps contiguous (3 = 128-byte runs, 2 = 64-byte runs)
    // balance: close a sweep rather than open a round that the tile cannot fill with register bits
    // mid_relabel: a ladder sweep whose tile is not the contiguous low block (its targets are high index bits: 128-byte
    // lines at a large stride, read AND written) stores its tile into the low block instead -- the tile's high bits trade
    // places with the non-tile bits below position T -- so that its writes are whole contiguous tiles
    // (PlannedSweep::mid_dstpos).  Gates keep arriving on the ORIGINAL positions; the planner tracks where they are now.
    // mid_relabel 2 (needs set_lookahead): the store also rotates the next targets into the LOW positions -- the coalescing
    // bits are in every tile anyway, so a sweep gets up to C targets for free: with balanced packing a QFT-30 runs as
    // 10 + 10 + 10 steps in two full rounds each instead of 12 + 9 + 9 in 3 + 2 + 2.  Spare tile slots are then filled
    // with the upcoming targets (they ride along as passengers and can be moved by the store).
    Planner(int n, int tile_bits, int coalesce_bits = 3, bool balance = false, int mid_relabel = 0);
    // the targets of the non-diagonal (G1) gates that will be add()ed, in order (original positions)
    void set_lookahead(const std::vector<int> &g1_targets) { la_ = g1_targets; la_at_ = 0; }
    // feed gates in program order
    void add(const LoweredGate &g);           // POLY or G1 only
    // flush everything that is pending (diagonal terms included) into sweeps
    void finish();
    // close the open sweep without touching pending diagonal terms
    void cut() { close_sweep(); open_sweep(); }
    // move the planned sweeps out (call after finish(), or at any time to drain
    // completed sweeps; the current partial sweep stays open unless finished)
    std::vector<PlannedSweep> take();
    bool has_pending() const;
    // make sure no pending diagonal term touches these physical bits (used
    // before a GENERIC gate); emits PHASE ops into the current sweep
    void flush_diag_touching(uint64_t bits_mask);
    PlanStats stats;

private:
    struct OpB {
        bool is_phase = false;
        bool is_lin = false;            // separable linear phase on all bits in `partners` (+ c0)
        int target = 0;                 // physical
        uint64_t cmask = 0;             // physical (G1)
        double m[8] = { 0 };
        int kind = OP_G1_GENERIC;
        // phase
        double base = 0;
        bool has_c0 = false;
        double c0 = 0;
        std::vector<std::pair<int, double>> partners;   // (physical bit, half-turns)
    };
    struct RoundB {
        std::vector<int> regs;          // physical
        std::vector<OpB> ops;
    };
    int n_, T_, C_;
    bool balance_;
    int mid_relabel_;
    std::vector<int> la_;               // set_lookahead
    size_t la_at_ = 0;                  // G1 gates add()ed so far
    std::vector<int> upcoming(size_t max_count) const;   // next distinct targets (original positions)
    std::vector<int> cur_, inv_;        // original position -> current position after the mid-plan relabels, and back
    std::vector<int> tile_;             // physical bits in the open sweep's tile
    std::vector<RoundB> rounds_;
    size_t nops_ = 0, nphase_ = 0;
    PendingDiag pd_;
    std::vector<PlannedSweep> done_;

    bool in_tile(int p) const;
    bool place_target(int t);           // make t a register bit of the current round (may open round/sweep)
    void emit_phase_for(int t);
    void emit_op(const OpB &op);
    void close_sweep();
    void open_sweep();
};

// relabelling sweep: dstpos[p] = destination position of source bit p
PlannedSweep build_permute_sweep(int n, int tile_bits, const std::vector<int> &dstpos, const std::vector<int> *forced_tile = nullptr);
// a bit permutation as a sequence of tile-closed passes that may run in place (source == destination)
struct InplacePass {
    std::vector<int> tile;      // bit positions of the tile, ascending; contains the low `coalesce` bits
    std::vector<int> dstpos;    // this pass moves the data at bit p to dstpos[p]; identity outside `tile`
};
std::vector<InplacePass> plan_inplace_relabel(int n, int tile_bits, int coalesce, const std::vector<int> &dstpos);
// turn the store side of an existing sweep into a relabelling one (out-of-place launch required)
// leaf_split: if the tile holds destination bits 0..9 (whole canonical leaves of 1024 amplitudes), lay the
// store pass out as one leaf per warp (lanes = destination bits 0..4, slots = bits 5..9) and set P.leaf_fuse
void set_relabel(SweepProgram &P, const std::vector<int> &dstpos, bool leaf_split = false);
bool can_fuse_relabel(const SweepProgram &P, const std::vector<int> &dstpos);

// re-express the shared-memory side of a dense ladder program in the order a TMA tensor load delivers the
// tile (planner.cpp); false = not possible, program untouched
bool apply_tma_layout(SweepProgram &P);

// reduce an angle in half-turns to [-1, 1)
double wrap_half_turns(double a);

}  // namespace q1t
