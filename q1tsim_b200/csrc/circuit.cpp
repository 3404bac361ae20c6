// circuit.cpp -- `Circuit` of the reference on the device engine.
//   builder          circuit.rs:161-554
//   execute family   circuit.rs:562-641
//   interpreter      circuit.rs:643-762 (do_execute_with)
//   histograms       circuit.rs:773-841
#include "circuit.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace q1t {

static CircuitError ok() { return CircuitError(); }
static CircuitError mkerr(int code, const std::string &m)
{
    CircuitError e;
    e.code = code;
    e.msg = m;
    return e;
}
static CircuitError invalid_qbit(size_t b)
{
    char buf[96];
    std::snprintf(buf, sizeof buf, "Invalid index %zu for a quantum bit", b);
    return mkerr(Q1T_ERR_INVALID_QBIT, buf);
}
static CircuitError invalid_cbit(size_t b)
{
    char buf[96];
    std::snprintf(buf, sizeof buf, "Invalid index %zu for a classical bit", b);
    return mkerr(Q1T_ERR_INVALID_CBIT, buf);
}

// description(): "H", "CX", "RX(1.2300)", "S†" ... (src/gates/*.rs `description`)
std::string GateSpec::description(bool with_values) const
{
    if (name.empty()) return user_desc.empty() ? "matrix gate" : user_desc;
    // every table name that starts with 'c' is C<..> of a base gate (controlled.rs:402-552)
    std::string inner = name, prefix;
    while (inner.size() > 1 && inner[0] == 'c') { prefix += "C"; inner = inner.substr(1); }
    std::string d;
    if (inner == "sdg") d = "S\xE2\x80\xA0";
    else if (inner == "tdg") d = "T\xE2\x80\xA0";
    else if (inner == "vdg") d = "V\xE2\x80\xA0";
    else if (inner == "swap") d = "Swap";
    else { d = inner; for (char &ch : d) if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 32); }
    if (!params.empty() && with_values) {
        d += "(";
        for (size_t i = 0; i < params.size(); ++i) {
            char buf[48];
            std::snprintf(buf, sizeof buf, "%s%.4f", i ? ", " : "", params[i].get());
            d += buf;
        }
        d += ")";
    }
    return prefix + d;
}

int GateSpec::evaluate(std::vector<std::complex<double>> &out) const
{
    if (name.empty()) { out = matrix; return (int)nr_bits; }
    double p[4] = { 0, 0, 0, 0 };
    for (size_t i = 0; i < params.size() && i < 4; ++i) p[i] = params[i].get();
    out.assign(64, std::complex<double>(0, 0));
    const int nb = builtin_gate_matrix(name.c_str(), p, params.size(), out.data());
    if (nb > 0) out.resize((size_t)1 << (2 * nb));
    return nb;
}

int gate_spec_from_name(const char *name, const Param *params, size_t nparams, GateSpec &out, std::string &err)
{
    out = GateSpec();
    std::string nm(name ? name : "");
    for (char &ch : nm) if (ch >= 'A' && ch <= 'Z') ch = (char)(ch + 32);
    // arity check the way the reference reports it (ffi.rs:216-305, error.rs ParseError)
    static const struct { const char *n; int np; } arity[] = {
        { "rx", 1 }, { "ry", 1 }, { "rz", 1 }, { "u1", 1 }, { "u2", 2 }, { "u3", 3 }, { "crx", 1 }, { "cry", 1 }, { "crz", 1 },
        { "cu1", 1 }, { "cu2", 2 }, { "cu3", 3 }, { "ccrx", 1 }, { "ccry", 1 }, { "ccrz", 1 }, { nullptr, 0 } };
    // (name -> number of qubits) is a fixed table: probe every name once per process, not once per gate added
    struct Known { std::string name; int nb, np; };
    static std::vector<Known> known;
    static std::mutex known_mu;
    int want = 0, nb = -1;
    {
        std::lock_guard<std::mutex> lk(known_mu);
        bool hit = false;
        for (const Known &k : known)
            if (k.name == nm) { nb = k.nb; want = k.np; hit = true; break; }
        if (!hit) {
            for (int a = 0; arity[a].n; ++a) if (nm == arity[a].n) want = arity[a].np;
            const double probe[4] = { 0.1, 0.2, 0.3, 0.4 };
            std::complex<double> tmp[64];
            nb = builtin_gate_matrix(nm.c_str(), probe, (size_t)want, tmp);
            if (nb != -1 && known.size() < 256) known.push_back({ nm, nb, want });
        }
    }
    if (nb == -1) { err = "Unknown gate \"" + std::string(name ? name : "") + "\""; return Q1T_ERR_PARSE; }
    out.name.swap(nm);
    if ((int)nparams != want && want > 0) {
        std::string up = out.name;
        for (char &ch : up) if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 32);
        char buf[160];
        std::snprintf(buf, sizeof buf, "Expected %d arguments to \"%s\" gate, got %zu", want, up.c_str(), nparams);
        err = buf;
        return Q1T_ERR_PARSE;
    }
    out.nr_bits = (size_t)nb;
    for (int i = 0; i < want; ++i) out.params.push_back(params[i]);
    return Q1T_OK;
}

// ---- builder -------------------------------------------------------------
CircuitError Circuit::add_gate(const GateSpec &g, const std::vector<size_t> &bits)
{
    for (size_t b : bits)
        if (b >= nr_qbits_) return invalid_qbit(b);          // circuit.rs:164-167
    ops_.emplace_back();
    CircuitOp &op = ops_.back();
    op.kind = CircuitOp::Gate;
    op.gate = g;
    op.bits = bits;
    return ok();
}

CircuitError Circuit::add_conditional_gate(const std::vector<size_t> &control, uint64_t target, const GateSpec &g,
                                           const std::vector<size_t> &bits)
{
    for (size_t b : control)
        if (b >= nr_cbits_) return invalid_cbit(b);          // circuit.rs:187-190
    for (size_t b : bits)
        if (b >= nr_qbits_) return invalid_qbit(b);
    CircuitOp op;
    op.kind = CircuitOp::ConditionalGate;
    op.gate = g;
    op.bits = bits;
    op.control = control;
    op.target = target;
    ops_.push_back(op);
    return ok();
}

CircuitError Circuit::measure_basis(size_t qbit, size_t cbit, Basis b)
{
    if (qbit >= nr_qbits_) return invalid_qbit(qbit);        // circuit.rs:209-216
    if (cbit >= nr_cbits_) return invalid_cbit(cbit);
    CircuitOp op;
    op.kind = CircuitOp::Measure;
    op.qbit = qbit; op.cbit = cbit; op.basis = b;
    ops_.push_back(op);
    return ok();
}

CircuitError Circuit::measure_all_basis(const std::vector<size_t> &cbits, Basis b)
{
    for (size_t c : cbits)
        if (c >= nr_cbits_) return invalid_cbit(c);          // circuit.rs:271-274
    CircuitOp op;
    op.kind = CircuitOp::MeasureAll;
    op.bits = cbits; op.basis = b;
    ops_.push_back(op);
    return ok();
}

CircuitError Circuit::peek_basis(size_t qbit, size_t cbit, Basis b)
{
    if (qbit >= nr_qbits_) return invalid_qbit(qbit);
    if (cbit >= nr_cbits_) return invalid_cbit(cbit);
    CircuitOp op;
    op.kind = CircuitOp::Peek;
    op.qbit = qbit; op.cbit = cbit; op.basis = b;
    ops_.push_back(op);
    return ok();
}

CircuitError Circuit::peek_all_basis(const std::vector<size_t> &cbits, Basis b)
{
    for (size_t c : cbits)
        if (c >= nr_cbits_) return invalid_cbit(c);
    CircuitOp op;
    op.kind = CircuitOp::PeekAll;
    op.bits = cbits; op.basis = b;
    ops_.push_back(op);
    return ok();
}

CircuitError Circuit::reset(size_t qbit)
{
    if (qbit >= nr_qbits_) return invalid_qbit(qbit);
    CircuitOp op;
    op.kind = CircuitOp::Reset;
    op.qbit = qbit;
    ops_.push_back(op);
    return ok();
}

void Circuit::reset_all()
{
    CircuitOp op;
    op.kind = CircuitOp::ResetAll;
    ops_.push_back(op);
}

CircuitError Circuit::barrier(const std::vector<size_t> &qbits)
{
    for (size_t b : qbits)
        if (b >= nr_qbits_) return invalid_qbit(b);
    CircuitOp op;
    op.kind = CircuitOp::Barrier;
    op.bits = qbits;
    ops_.push_back(op);
    return ok();
}

// ---- execution -------------------------------------------------------------
CircuitError Circuit::state_err(int rc)
{
    return mkerr(rc, q_state_ ? q_state_->last_error() : "no state");
}

// circuit.rs:562-600: fresh quantum state (always the statevector backend here),
// classical register cleared
// Q1T_HOST_PROFILE=1: one line per execute() on stderr with the host-side phases (microseconds)
static bool host_profile_on()
{
    static const bool on = std::getenv("Q1T_HOST_PROFILE") && std::atoi(std::getenv("Q1T_HOST_PROFILE")) != 0;
    return on;
}
static double now_us()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// execute() on a state sharded over several devices (circuit.rs:562-600 for circuits beyond one GPU).  The op list is known
// in full: it is walked on a data-less ShardedVectorState from two initial layouts -- |0..0> is symmetric, so any layout
// is a legal start -- and run from the cheaper one (see q1tsim_b200/sharded.py run_ops; a QFT ends canonical without a
// single exchange).
CircuitError Circuit::execute_sharded(size_t nr_shots, q1t_rng rng)
{
    static const double H[8] = { 0.70710678118654752440, 0, 0.70710678118654752440, 0, 0.70710678118654752440, 0, -0.70710678118654752440, -0.0 };
    static const double S[8] = { 1, 0, 0, 0, 0, 0, 0, 1 };
    static const double SDG[8] = { 1, 0, 0, 0, 0, 0, -0.0, -1 };
    static const double SWAPM[32] = { 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0 };
    for (const CircuitOp &op : ops_)
        if (op.kind != CircuitOp::Gate && op.kind != CircuitOp::MeasureAll && op.kind != CircuitOp::PeekAll && op.kind != CircuitOp::Barrier)
            return mkerr(Q1T_ERR_UNSUPPORTED, "a circuit on a sharded state may hold gates, measure_all / peek_all and barriers");
    // a collapsing measure_all leaves one basis-state column per outcome; the gates that restore an X or Y basis afterwards
    // (circuit.rs:705-735) would have to remap a multi-column state, which the peer group does not do
    for (const CircuitOp &op : ops_)
        if (op.kind == CircuitOp::MeasureAll && op.basis != Basis::Z)
            return mkerr(Q1T_ERR_UNSUPPORTED, "measure_all in the X or Y basis is not available on a sharded state (peek_all is; measure_all in the Z basis is)");
    q_state_.reset();
    // matrices of the gates (parameters are read now, at execution time)
    std::vector<std::vector<std::complex<double>>> mats(ops_.size());
    size_t nprefix = 0;
    bool in_prefix = true;
    for (size_t t = 0; t < ops_.size(); ++t) {
        if (ops_[t].kind != CircuitOp::Gate) { in_prefix = false; continue; }
        if (ops_[t].gate.evaluate(mats[t]) < 0) return mkerr(Q1T_ERR_PARSE, "invalid gate");
        if (in_prefix) nprefix = t + 1;
    }
    // dest[q]: the logical qubit the data labelled q ends as after the Swap relabels of the leading gate run
    std::vector<int> dest(nr_qbits_);
    for (size_t q = 0; q < nr_qbits_; ++q) dest[q] = (int)q;
    for (size_t t = nprefix; t-- > 0;) {
        const CircuitOp &op = ops_[t];
        if (op.bits.size() == 2 && mats[t].size() == 16 && std::memcmp(mats[t].data(), SWAPM, sizeof SWAPM) == 0)
            std::swap(dest[op.bits[0]], dest[op.bits[1]]);
    }
    auto walk = [&](ShardedVectorState &st, size_t upto, uint64_t *res, size_t nres) -> int {
        for (size_t t = 0; t < upto; ++t) {
            const CircuitOp &op = ops_[t];
            int rc = Q1T_OK;
            if (op.kind == CircuitOp::Gate)
                rc = st.apply_gate(reinterpret_cast<const double *>(mats[t].data()), (size_t)1 << op.bits.size(), op.bits.data(), op.bits.size(),
                                   op.gate.description().c_str());
            else if (op.kind == CircuitOp::MeasureAll || op.kind == CircuitOp::PeekAll) {
                if (op.basis == Basis::X) rc = st.apply_unary_gate_all(H, 2, "H");
                if (op.basis == Basis::Y) { rc = st.apply_unary_gate_all(SDG, 2, "Sdg"); if (!rc) rc = st.apply_unary_gate_all(H, 2, "H"); }
                if (!rc) rc = st.measure_all_into(op.bits.data(), op.bits.size(), res, nres, rng, op.kind == CircuitOp::MeasureAll);
                if (!rc && op.basis == Basis::X) rc = st.apply_unary_gate_all(H, 2, "H");
                if (!rc && op.basis == Basis::Y) { rc = st.apply_unary_gate_all(H, 2, "H"); if (!rc) rc = st.apply_unary_gate_all(S, 2, "S"); }
            }
            if (rc) return rc;
        }
        return Q1T_OK;
    };
    bool use_dest = false;
    {
        bool ident = true;
        for (size_t q = 0; q < nr_qbits_; ++q) ident = ident && dest[q] == (int)q;
        if (!ident && nprefix > 0) {
            uint64_t cost[2] = { 0, 0 };
            for (int c = 0; c < 2; ++c) {
                ShardedVectorState dry(nr_qbits_, nr_shots, devices, true);
                if (dry.init_zero_state()) return mkerr(Q1T_ERR_INVALID_ARGUMENT, dry.last_error());
                if (c == 1 && dry.set_initial_layout(dest)) return mkerr(Q1T_ERR_INVALID_ARGUMENT, dry.last_error());
                const int rc = walk(dry, nprefix, nullptr, 0);
                if (rc) return mkerr(rc, dry.last_error());
                if (dry.canonicalize()) return mkerr(Q1T_ERR_INVALID_ARGUMENT, dry.last_error());
                cost[c] = dry.remaps * 1000 + dry.local_relabels;
            }
            use_dest = cost[1] < cost[0];
        }
    }
    // the shards, their registered buffers and the peer mappings are kept from one execute() to the next (a fresh
    // state is |0..0> either way: reset_all)
    int rc;
    if (s_state_ && s_state_->nr_shots() == nr_shots && s_state_->nr_shards() == devices.size() && s_devices_ == devices)
        rc = s_state_->reset_all();
    else {
        s_state_.reset();
        s_state_.reset(new ShardedVectorState(nr_qbits_, nr_shots, devices));
        s_devices_ = devices;
        rc = s_state_->init_zero_state();
    }
    if (!rc && use_dest) rc = s_state_->set_initial_layout(dest);
    if (rc) {
        CircuitError e = mkerr(rc, s_state_->last_error());
        s_state_.reset();
        has_cstate_ = false;
        return e;
    }
    c_state_.assign(nr_shots, 0);
    has_cstate_ = true;
    rc = walk(*s_state_, ops_.size(), c_state_.data(), c_state_.size());
    if (!rc) rc = s_state_->canonicalize();         // the reference's execute() returns with the state fully evolved
    sharded_counters[0] = s_state_->remaps;
    sharded_counters[1] = s_state_->exchanges;
    sharded_counters[2] = s_state_->local_relabels;
    if (rc) return mkerr(rc, s_state_->last_error());
    return ok();
}

CircuitError Circuit::execute(size_t nr_shots, q1t_rng rng, const double *qubit_coefs)
{
    if (devices.size() >= 2) {
        if (qubit_coefs) return mkerr(Q1T_ERR_UNSUPPORTED, "execute_with a product state is not available on a sharded state");
        return execute_sharded(nr_shots, rng);
    }
    s_state_.reset();
    s_devices_.clear();
    // one device holds 1..34 qubits (capi.cpp make_state checks the same for the inner ABI)
    if (nr_qbits_ < 1 || nr_qbits_ > 34) return mkerr(Q1T_ERR_INVALID_ARGUMENT, "the number of qubits must be in 1..34 for one device");
    const double t0 = host_profile_on() ? now_us() : 0.0;
    q_state_.reset(new DeviceVectorState(nr_qbits_, nr_shots, device));
    const int rc = qubit_coefs ? q_state_->init_from_qubit_coefs(qubit_coefs) : q_state_->init_zero_state();
    if (rc) {
        CircuitError e = state_err(rc);
        q_state_.reset();
        has_cstate_ = false;
        return e;
    }
    c_state_.assign(nr_shots, 0);
    has_cstate_ = true;
    fresh_state_ = true;               // identity layout: a cached lowering of the leading gate run may be replayed
    if (host_profile_on()) std::fprintf(stderr, "q1t host profile: new state + init %.0f us\n", now_us() - t0);
    const CircuitError ce = do_execute(rng);
    if (q_state_) q_state_->record_lowered(nullptr);
    return ce;
}

// circuit.rs:618-641
CircuitError Circuit::reexecute(q1t_rng rng)
{
    if (!has_cstate_ || !q_state_) return mkerr(Q1T_ERR_NOT_EXECUTED, "The circuit has not been executed yet");
    return do_execute(rng);
}

CircuitError Circuit::set_cstate(const uint64_t *w, size_t n)
{
    if (!q_state_) {
        if (nr_qbits_ < 1 || nr_qbits_ > 34) return mkerr(Q1T_ERR_INVALID_ARGUMENT, "the number of qubits must be in 1..34 for one device");
        q_state_.reset(new DeviceVectorState(nr_qbits_, n, device));
        const int rc = q_state_->init_zero_state();
        if (rc) { CircuitError e = state_err(rc); q_state_.reset(); return e; }
    }
    if (n != q_state_->nr_shots()) return mkerr(Q1T_ERR_INVALID_ARGUMENT, "classical register length must equal the number of shots");
    c_state_.assign(w, w + n);
    has_cstate_ = true;
    return ok();
}

// process-wide cache of lowered gate runs, keyed by content (see do_execute)
namespace {
struct LoweredRun { uint64_t key[2]; size_t n, len; uint64_t stamp; std::vector<LoweredGate> gates; };
std::mutex g_lowered_mu;
std::vector<LoweredRun> g_lowered_runs;
uint64_t g_lowered_clock = 0;
}  // namespace

void Circuit::publish_lowered()
{
    if (lowered_.size() != lowered_len_) return;
    std::lock_guard<std::mutex> lk(g_lowered_mu);
    if (g_lowered_runs.size() >= 8) {
        size_t lru = 0;
        for (size_t i = 1; i < g_lowered_runs.size(); ++i)
            if (g_lowered_runs[i].stamp < g_lowered_runs[lru].stamp) lru = i;
        g_lowered_runs.erase(g_lowered_runs.begin() + lru);
    }
    g_lowered_runs.push_back({ { lowered_key_[0], lowered_key_[1] }, nr_qbits_, lowered_len_, ++g_lowered_clock, lowered_ });
}

// circuit.rs:643-762
CircuitError Circuit::do_execute(q1t_rng rng)
{
    DeviceVectorState &q = *q_state_;
    static const double H[8] = { 0.70710678118654752440, 0, 0.70710678118654752440, 0, 0.70710678118654752440, 0, -0.70710678118654752440, -0.0 };
    static const double S[8] = { 1, 0, 0, 0, 0, 0, 0, 1 };
    static const double SDG[8] = { 1, 0, 0, 0, 0, 0, -0.0, -1 };
    std::vector<std::complex<double>> mat;
    std::vector<uint8_t> apply;
#define TRY(expr) do { const int rc__ = (expr); if (rc__) return state_err(rc__); } while (0)
    const bool prof = host_profile_on();
    double t_kind[2] = { 0.0, 0.0 }, t_eval = 0.0;          // [0] gates (evaluate + lower + queue), [1] everything that observes the state
    // The leading run of plain gates of a circuit whose parameters are all values (no pointer read at execute time,
    // gates/parameter.rs:21-50) lowers to the same list every time it starts from a fresh state: it is lowered once
    // and replayed (execute() of a built circuit: 480 x matrix() + classification per call otherwise).
    size_t skip = 0;
    std::vector<LoweredGate> *recording = nullptr;
    if (fresh_state_) {
        size_t run = 0;
        bool constant = true;
        while (run < ops_.size() && ops_[run].kind == CircuitOp::Gate) {
            for (const Param &pr : ops_[run].gate.params) constant = constant && pr.ptr == nullptr;
            ++run;
        }
        if (constant && run >= 16 && !(lowered_valid_ && lowered_len_ == run)) {
            // a circuit object seen for the first time: the same gate run may have been lowered for another object (a
            // host layer that rebuilds its circuit for every execute()): process-wide cache keyed by the run's content
            uint64_t key = 1469598103934665603ull, key2 = 0x9E3779B97F4A7C15ull;
            auto mix = [&](const void *p, size_t nbytes) {
                const unsigned char *b = static_cast<const unsigned char *>(p);
                for (size_t i = 0; i < nbytes; ++i) {
                    key ^= b[i]; key *= 1099511628211ull;
                    key2 = (key2 ^ b[i]) * 0xFF51AFD7ED558CCDull; key2 ^= key2 >> 29;
                }
            };
            const size_t hdr[2] = { nr_qbits_, run };
            mix(hdr, sizeof hdr);
            for (size_t t = 0; t < run; ++t) {
                const GateSpec &g = ops_[t].gate;
                mix(g.name.data(), g.name.size() + 0);
                const size_t sep[2] = { g.params.size(), g.matrix.size() };
                mix(sep, sizeof sep);
                for (const Param &pr : g.params) mix(&pr.value, sizeof pr.value);
                if (!g.matrix.empty()) mix(g.matrix.data(), sizeof(std::complex<double>) * g.matrix.size());
                mix(ops_[t].bits.data(), sizeof(size_t) * ops_[t].bits.size());
            }
            lowered_key_[0] = key; lowered_key_[1] = key2;
            std::lock_guard<std::mutex> lk(g_lowered_mu);
            for (LoweredRun &lr : g_lowered_runs)
                if (lr.key[0] == key && lr.key[1] == key2 && lr.n == nr_qbits_ && lr.len == run) {
                    lowered_ = lr.gates;
                    lowered_len_ = run;
                    lowered_valid_ = true;
                    lr.stamp = ++g_lowered_clock;
                    break;
                }
        }
        if (constant && run >= 16) {
            if (lowered_valid_ && lowered_len_ == run) {
                TRY(q.apply_lowered(lowered_));
                skip = run;
            } else {
                lowered_.clear();
                lowered_len_ = run;
                recording = &lowered_;
                q.record_lowered(recording);
            }
        }
    }
    fresh_state_ = false;
    size_t op_index = 0;
    for (const CircuitOp &op : ops_) {
        if (recording && op_index == lowered_len_) { q.record_lowered(nullptr); recording = nullptr; lowered_valid_ = true; publish_lowered(); }
        if (op_index++ < skip) continue;
        const double t_op = prof ? now_us() : 0.0;
        struct Tick {
            bool on; double t0; double &acc;
            ~Tick() { if (on) acc += now_us() - t0; }
        } tick{ prof, t_op, t_kind[op.kind == CircuitOp::Gate || op.kind == CircuitOp::ConditionalGate ? 0 : 1] };
        switch (op.kind) {
        case CircuitOp::Gate: {
            const int nb = op.gate.evaluate(mat);
            if (prof) t_eval += now_us() - t_op;
            if (nb < 0) return mkerr(Q1T_ERR_PARSE, "invalid gate");
            TRY(q.apply_gate(reinterpret_cast<const double *>(mat.data()), (size_t)1 << nb, op.bits.data(), op.bits.size(),
                             op.gate.description().c_str()));
            break;
        }
        case CircuitOp::ConditionalGate: {
            // control word bit k = classical bit control[k] (first listed = LSB), circuit.rs:655-665
            apply.assign(c_state_.size(), 0);
            for (size_t s = 0; s < c_state_.size(); ++s) {
                uint64_t w = 0;
                for (size_t idst = 0; idst < op.control.size(); ++idst) w |= ((c_state_[s] >> op.control[idst]) & 1ull) << idst;
                apply[s] = w == op.target;
            }
            const int nb = op.gate.evaluate(mat);
            if (nb < 0) return mkerr(Q1T_ERR_PARSE, "invalid gate");
            TRY(q.apply_conditional_gate(apply.data(), apply.size(), reinterpret_cast<const double *>(mat.data()), (size_t)1 << nb,
                                         op.bits.data(), op.bits.size(), op.gate.description().c_str()));
            break;
        }
        case CircuitOp::Measure:
        case CircuitOp::Peek: {
            const bool collapse = op.kind == CircuitOp::Measure;
            const size_t qb = op.qbit;
            if (op.basis == Basis::X) TRY(q.apply_gate(H, 2, &qb, 1, "H"));
            if (op.basis == Basis::Y) { TRY(q.apply_gate(SDG, 2, &qb, 1, "S\xE2\x80\xA0")); TRY(q.apply_gate(H, 2, &qb, 1, "H")); }
            TRY(q.measure_into(op.qbit, op.cbit, c_state_.data(), c_state_.size(), rng, collapse));
            if (op.basis == Basis::X) TRY(q.apply_gate(H, 2, &qb, 1, "H"));
            if (op.basis == Basis::Y) { TRY(q.apply_gate(H, 2, &qb, 1, "H")); TRY(q.apply_gate(S, 2, &qb, 1, "S")); }
            break;
        }
        case CircuitOp::MeasureAll:
        case CircuitOp::PeekAll: {
            const bool collapse = op.kind == CircuitOp::MeasureAll;
            if (op.basis == Basis::X) TRY(q.apply_unary_gate_all(H, 2, "H"));
            if (op.basis == Basis::Y) { TRY(q.apply_unary_gate_all(SDG, 2, "S\xE2\x80\xA0")); TRY(q.apply_unary_gate_all(H, 2, "H")); }
            TRY(q.measure_all_into(op.bits.data(), op.bits.size(), c_state_.data(), c_state_.size(), rng, collapse));
            if (op.basis == Basis::X) TRY(q.apply_unary_gate_all(H, 2, "H"));
            if (op.basis == Basis::Y) { TRY(q.apply_unary_gate_all(H, 2, "H")); TRY(q.apply_unary_gate_all(S, 2, "S")); }
            break;
        }
        case CircuitOp::Reset:
            TRY(q.reset(op.qbit, rng));
            break;
        case CircuitOp::ResetAll:
            TRY(q.reset_all());
            break;
        case CircuitOp::Barrier:
            break;
        }
    }
#undef TRY
    if (recording) { q.record_lowered(nullptr); lowered_valid_ = true; publish_lowered(); }        // (the circuit is gates only)
    // the reference's execute() returns with the state fully evolved; queued gates after the
    // last measurement are run here so that errors surface now
    const double t_fl = prof ? now_us() : 0.0;
    const int rc = q.flush();
    if (prof)
        std::fprintf(stderr, "q1t host profile: %zu ops: gates %.0f us (matrix() %.0f us), measure/peek/reset %.0f us, final flush %.0f us\n",
                     ops_.size(), t_kind[0], t_eval, t_kind[1], now_us() - t_fl);
    if (rc) return state_err(rc);
    return ok();
}

// ---- histograms (circuit.rs:773-841) ----------------------------------------
std::map<uint64_t, size_t> Circuit::histogram() const
{
    std::map<uint64_t, size_t> h;
    for (uint64_t k : c_state_) h[k] += 1;
    return h;
}

std::map<std::string, size_t> Circuit::histogram_string() const
{
    std::map<std::string, size_t> h;
    for (const auto &kv : histogram()) {
        // format!("{:0width$b}", key, width = nr_cbits): last character = classical bit 0
        std::string s;
        uint64_t k = kv.first;
        while (k) { s.insert(s.begin(), (char)('0' + (k & 1))); k >>= 1; }
        while (s.size() < nr_cbits_) s.insert(s.begin(), '0');
        if (s.empty()) s = "0";
        h[s] += kv.second;
    }
    return h;
}

}  // namespace q1t
