// sharded.cpp -- ShardedVectorState (see sharded.h and DESIGN.md 6).
//
// Rank r of P = 2^g holds the amplitudes whose top g index bits equal r (qubit 0 is the most significant index bit,
// vectorstate.rs:249-250).  where_[q] says where logical qubit q lives: at a rank bit, or at a local engine qubit.
#include "sharded.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace q1t {

typedef std::complex<double> cplx;

static const double kSwap[32] = { 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0 };

// does the k-qubit matrix never couple different values of gate bit j?  (gate bit j is bit k-1-j of the matrix index,
// gates.rs:53-80)
static bool acts_diagonally(const cplx *m, int k, int j)
{
    const int G = 1 << k, mask = 1 << (k - 1 - j);
    for (int r = 0; r < G; ++r)
        for (int c = 0; c < G; ++c)
            if ((r & mask) != (c & mask) && m[r * G + c] != cplx(0, 0)) return false;
    return true;
}

// sub-matrix for fixed values of some gate bits (mask / want on the matrix index), on the remaining bits in order
static std::vector<cplx> select_block(const cplx *m, int k, int mask, int want)
{
    const int G = 1 << k;
    std::vector<int> keep;
    for (int x = 0; x < G; ++x)
        if ((x & mask) == want) keep.push_back(x);
    std::vector<cplx> out(keep.size() * keep.size());
    for (size_t a = 0; a < keep.size(); ++a)
        for (size_t b = 0; b < keep.size(); ++b) out[a * keep.size() + b] = m[keep[a] * G + keep[b]];
    return out;
}

ShardedVectorState::ShardedVectorState(size_t nr_bits, size_t nr_shots, const std::vector<int> &devices, bool dry)
    : n_((int)nr_bits), g_(0), P_((int)devices.size()), nl_(0), shots_(nr_shots), dry_(dry), devices_(devices)
{
    while ((1 << g_) < P_) ++g_;
    nl_ = n_ - g_;
    where_.resize(n_);
    for (int q = 0; q < n_; ++q) where_[q] = canonical(q);
    pin_.assign(g_, 0);
}

ShardedVectorState::~ShardedVectorState()
{
    // every shard closes its peer mappings before any shard frees the buffers the others have mapped
    for (auto &s : shards_)
        if (s) s->group_close();
    shards_.clear();
}

int ShardedVectorState::qubit_at(bool global, int idx) const
{
    for (int q = 0; q < n_; ++q)
        if (where_[q].global == global && where_[q].idx == idx) return q;
    return -1;
}

int ShardedVectorState::open_group()
{
    std::vector<void *> ptrs(3 * (size_t)P_, nullptr);
    unsigned char handles[192];
    for (int r = 0; r < P_; ++r) {
        int rc = shards_[r]->group_export(handles, &ptrs[3 * r]);
        if (rc) return shard_fail(r, rc);
    }
    // shards on different devices of this process reach each other's buffers through peer access
    for (int a = 0; a < P_; ++a)
        for (int b = 0; b < P_; ++b) {
            if (devices_[a] == devices_[b]) continue;
            cudaSetDevice(devices_[a]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices_[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
                return fail(Q1T_ERR_CUDA, std::string("no peer access between devices: ") + cudaGetErrorString(e));
            }
            cudaGetLastError();
        }
    for (int r = 0; r < P_; ++r) {
        int rc = shards_[r]->group_open((size_t)P_, (size_t)r, nullptr, ptrs.data());
        if (rc) return shard_fail(r, rc);
    }
    return Q1T_OK;
}

int ShardedVectorState::init_zero_state()
{
    if (P_ < 2 || (P_ & (P_ - 1)) || P_ > 32) return fail(Q1T_ERR_INVALID_ARGUMENT, "the number of shards must be a power of two between 2 and 32");
    if (nl_ < 10) return fail(Q1T_ERR_INVALID_ARGUMENT, "sharded states need at least 10 local qubits per shard (one canonical leaf)");
    if (nl_ > 34) return fail(Q1T_ERR_INVALID_ARGUMENT, "a shard holds at most 34 qubits");
    for (int q = 0; q < n_; ++q) where_[q] = canonical(q);
    pin_.assign(g_, 0);
    if (dry_) return Q1T_OK;
    // replicated start: every shard is |0..0> on its local qubits while every rank bit is pinned to 0
    for (int r = 0; r < P_; ++r) {
        shards_.emplace_back(new DeviceVectorState((size_t)nl_, shots_, devices_[r]));
        // one host thread drives all shards here: a remap recorded for the next sweep (engine option fused_remap) would launch
        // its barriers from inside one shard's flush while the thread has not reached the next shard yet -- always off
        shards_.back()->set_option("fused_remap", 0);
        int rc = shards_[r]->init_zero_state();
        if (rc) return shard_fail(r, rc);
    }
    return open_group();
}

int ShardedVectorState::reset_all()
{
    for (int q = 0; q < n_; ++q) where_[q] = canonical(q);
    pin_.assign(g_, 0);
    for (int r = 0; r < (int)shards_.size(); ++r) {
        int rc = shards_[r]->reset_all();
        if (rc) return shard_fail(r, rc);
    }
    return Q1T_OK;
}

int ShardedVectorState::set_initial_layout(const std::vector<int> &dest)
{
    // dest[q] = the logical qubit the data labelled q ends as after the run's Swap relabels: label q starts where that
    // qubit belongs, so the run ends in the canonical layout.  Only legal while the state is the fresh |0..0>.
    if ((int)dest.size() != n_) return fail(Q1T_ERR_INVALID_ARGUMENT, "set_initial_layout: one entry per qubit");
    for (int v : pin_)
        if (v != 0) return fail(Q1T_ERR_INVALID_ARGUMENT, "set_initial_layout: the state is not a fresh |0..0>");
    std::vector<char> seen(n_, 0);
    for (int q = 0; q < n_; ++q) {
        if (dest[q] < 0 || dest[q] >= n_ || seen[dest[q]]) return fail(Q1T_ERR_INVALID_ARGUMENT, "set_initial_layout: not a permutation");
        seen[dest[q]] = 1;
    }
    for (int q = 0; q < n_; ++q) where_[q] = canonical(dest[q]);
    return Q1T_OK;
}

int ShardedVectorState::local_swap(int a, int b)
{
    ++local_relabels;
    if (dry_) return Q1T_OK;
    const size_t bits[2] = { (size_t)a, (size_t)b };
    for (int r = 0; r < P_; ++r) {
        int rc = shards_[r]->apply_gate(kSwap, 4, bits, 2, "Swap");
        if (rc) return shard_fail(r, rc);
    }
    return Q1T_OK;
}

int ShardedVectorState::depin(int i)
{
    const int v = pin_[i];
    if (v < 0) return Q1T_OK;
    pin_[i] = -1;
    if (dry_) return Q1T_OK;
    // the ranks on the wrong side of a pin that stops being a known basis value hold nothing
    const uint64_t zero = UINT64_MAX;
    const size_t cnt = shots_;
    for (int r = 0; r < P_; ++r)
        if (rank_bit(r, i) != v) {
            int rc = shards_[r]->replace_columns(1, &zero, &cnt);
            if (rc) return shard_fail(r, rc);
        }
    return Q1T_OK;
}

int ShardedVectorState::depin_all()
{
    for (int i = 0; i < g_; ++i) {
        int rc = depin(i);
        if (rc) return rc;
    }
    return Q1T_OK;
}

int ShardedVectorState::exchange_multi(const std::vector<std::pair<int, int>> &trades)
{
    const int k = (int)trades.size();
    if (k == 0) return Q1T_OK;
    if (k > kMaxRemapBits || nl_ < k + 1) {
        for (const auto &t : trades) {
            int rc = exchange_multi(std::vector<std::pair<int, int>>(1, t));
            if (rc) return rc;
        }
        return Q1T_OK;
    }
    // the traded index bits should be high ones (long contiguous runs on the wire): relabel the victims into the top-k
    // engine qubits first (a zero-byte Swap relabel inside the engine, undone by its next sweep)
    std::vector<int> free_top;
    for (int t = 0; t < k; ++t) {
        bool used = false;
        for (const auto &tr : trades) used = used || where_[tr.second].idx == t;
        if (!used) free_top.push_back(t);
    }
    for (const auto &tr : trades) {
        const int v = tr.second, j = where_[v].idx;
        if (j >= k) {
            const int t = free_top.front();
            free_top.erase(free_top.begin());
            const int qt = qubit_at(false, t);
            int rc = local_swap(t, j);
            if (rc) return rc;
            where_[qt] = Where{ false, j };
            where_[v] = Where{ false, t };
        }
    }
    if (!dry_) {
        int rb[kMaxRemapBits];
        size_t lq[kMaxRemapBits];
        for (int j = 0; j < k; ++j) { rb[j] = trades[j].first; lq[j] = (size_t)where_[trades[j].second].idx; }
        // two passes: whatever may allocate, free or wait first, on every shard; then launches only.  With several shards
        // on one device a cudaFree between two shards' barrier launches would wait for a kernel that waits for the other.
        for (int r = 0; r < P_; ++r) {
            int rc = shards_[r]->group_remap_prepare();
            if (rc) return shard_fail(r, rc);
        }
        for (int r = 0; r < P_; ++r) {                 // enqueued on every shard; the device-side barriers order them
            int rc = shards_[r]->group_remap_issue((size_t)k, rb, lq);
            if (rc) return shard_fail(r, rc);
        }
    }
    ++remaps;
    exchanges += k;
    for (const auto &tr : trades) {
        const int qg = qubit_at(true, tr.first), v = tr.second;
        const Where wv = where_[v];
        where_[v] = Where{ true, tr.first };
        where_[qg] = wv;
    }
    return Q1T_OK;
}

int ShardedVectorState::bring_local(int q, const std::vector<int> &keep)
{
    if (!where_[q].global) return Q1T_OK;
    const int i = where_[q].idx;
    int rc = depin(i);
    if (rc) return rc;
    int victim = -1;
    for (int j = 0; j < nl_ && victim < 0; ++j) {
        const int c = qubit_at(false, j);
        if (c >= 0 && std::find(keep.begin(), keep.end(), c) == keep.end()) victim = c;
    }
    if (victim < 0) return fail(Q1T_ERR_UNSUPPORTED, "no local qubit left to evict");
    return exchange_multi(std::vector<std::pair<int, int>>(1, std::make_pair(i, victim)));
}

int ShardedVectorState::apply_gate(const double *mat, size_t dim, const size_t *bits, size_t k_, const char *desc)
{
    if (!mat || (!bits && k_)) return fail(Q1T_ERR_INVALID_ARGUMENT, "NULL pointer argument");
    size_t gate_bits = 0;
    while (((size_t)1 << gate_bits) < dim) ++gate_bits;
    if (((size_t)1 << gate_bits) != dim || gate_bits != k_) {
        char buf[256];
        std::snprintf(buf, sizeof buf, "Expected %zu bits for \"%s\", got %zu", gate_bits, desc ? desc : "gate", k_);
        return fail(Q1T_ERR_INVALID_NR_BITS, buf);
    }
    const int k = (int)k_, G = 1 << k;
    if (k == 0) return Q1T_OK;
    if (k > 6) return fail(Q1T_ERR_UNSUPPORTED, "gates on more than 6 qubits are not supported on a sharded state");
    for (int j = 0; j < k; ++j) {
        if (bits[j] >= (size_t)n_) {
            char buf[128];
            std::snprintf(buf, sizeof buf, "Invalid index %zu for a quantum bit", bits[j]);
            return fail(Q1T_ERR_INVALID_QBIT, buf);
        }
        for (int i = 0; i < j; ++i)
            if (bits[i] == bits[j]) return fail(Q1T_ERR_INVALID_ARGUMENT, "duplicate qubit index in a gate");
    }
    const cplx *m = reinterpret_cast<const cplx *>(mat);
    if (k == 2 && std::memcmp(mat, kSwap, sizeof kSwap) == 0) {        // swap.rs:78-88 as a relabel
        std::swap(where_[bits[0]], where_[bits[1]]);
        return Q1T_OK;
    }
    std::vector<char> diag(k);
    for (int j = 0; j < k; ++j) diag[j] = acts_diagonally(m, k, j);
    if (k == 1 && where_[bits[0]].global && !diag[0] && pin_[where_[bits[0]].idx] >= 0) {
        // a one-qubit gate on a rank bit pinned to v: column v of its matrix.  Two non-zero entries: every rank takes the
        // entry of its own bit as a scalar and the bit is free; one: the bit stays pinned (X, Y)
        const int i = where_[bits[0]].idx, v = pin_[i];
        const cplx col[2] = { m[0 * 2 + v], m[1 * 2 + v] };
        const bool two = col[0] != cplx(0, 0) && col[1] != cplx(0, 0);
        const int only = col[0] != cplx(0, 0) ? 0 : 1;
        pin_[i] = two ? -1 : only;
        if (dry_) return Q1T_OK;
        for (int r = 0; r < P_; ++r) {
            const cplx s = two ? col[rank_bit(r, i)] : col[only];
            if (s == cplx(1, 0)) continue;
            int rc = shards_[r]->scale_all(s.real(), s.imag());
            if (rc) return shard_fail(r, rc);
        }
        return Q1T_OK;
    }
    std::vector<int> keep(bits, bits + k);
    for (int j = 0; j < k; ++j)
        if (where_[bits[j]].global && !diag[j]) {
            int rc = bring_local((int)bits[j], keep);
            if (rc) return rc;
        }
    if (dry_) return Q1T_OK;
    // every shard applies the block of the matrix that its rank bits (or the pinned values) select
    int mask = 0;
    std::vector<size_t> lbits;
    for (int j = 0; j < k; ++j) {
        if (where_[bits[j]].global) mask |= 1 << (k - 1 - j);
        else lbits.push_back((size_t)where_[bits[j]].idx);
    }
    for (int r = 0; r < P_; ++r) {
        int want = 0;
        for (int j = 0; j < k; ++j)
            if (where_[bits[j]].global) {
                const int i = where_[bits[j]].idx;
                const int v = pin_[i] >= 0 ? pin_[i] : rank_bit(r, i);
                want |= v << (k - 1 - j);
            }
        int rc;
        if (mask == 0) rc = shards_[r]->apply_gate(mat, dim, lbits.data(), lbits.size(), desc);
        else {
            const std::vector<cplx> blk = select_block(m, k, mask, want);
            if (lbits.empty()) {
                const cplx s = blk[0];
                if (s == cplx(1, 0)) continue;
                rc = shards_[r]->scale_all(s.real(), s.imag());       // a rank-dependent scalar
            } else rc = shards_[r]->apply_gate(reinterpret_cast<const double *>(blk.data()), (size_t)1 << lbits.size(), lbits.data(),
                                               lbits.size(), desc);
        }
        if (rc) return shard_fail(r, rc);
    }
    (void)G;
    return Q1T_OK;
}

int ShardedVectorState::apply_unary_gate_all(const double *mat, size_t dim, const char *desc)
{
    for (size_t b = 0; b < (size_t)n_; ++b) {
        int rc = apply_gate(mat, dim, &b, 1, desc);
        if (rc) return rc;
    }
    return Q1T_OK;
}

int ShardedVectorState::canonicalize()
{
    int rc = depin_all();
    if (rc) return rc;
    // the common case -- every misplaced global qubit sits on chip -- is one multi-bit remap
    std::vector<std::pair<int, int>> trades;
    bool simple = true;
    for (int q = 0; q < g_; ++q) {
        const Where want = canonical(q);
        if (where_[q].global && where_[q].idx == want.idx) continue;
        if (where_[q].global) simple = false;
        trades.push_back(std::make_pair(want.idx, q));
    }
    if (!trades.empty() && simple) {
        rc = exchange_multi(trades);
        if (rc) return rc;
    }
    for (int q = 0; q < g_; ++q) {
        const Where want = canonical(q);
        if (where_[q].global && where_[q].idx == want.idx) continue;
        if (where_[q].global) {                      // q sits at another rank bit: bring it on chip first
            rc = bring_local(q, std::vector<int>());
            if (rc) return rc;
        }
        rc = exchange_multi(std::vector<std::pair<int, int>>(1, std::make_pair(want.idx, q)));
        if (rc) return rc;
    }
    for (int q = g_; q < n_; ++q) {                  // local part: qubit q >= g is local engine qubit q - g
        const int want = q - g_, cur = where_[q].idx;
        if (cur == want) continue;
        const int other = qubit_at(false, want);
        rc = local_swap(cur, want);
        if (rc) return rc;
        where_[q] = Where{ false, want };
        where_[other] = Where{ false, cur };
    }
    return Q1T_OK;
}

int ShardedVectorState::column_total(double *out)
{
    int rc = canonicalize();
    if (rc) return rc;
    double run = 0.0;
    if (shards_[0]->nr_leaves() >= kCanonBlock) {
        const size_t nb = shards_[0]->nr_leaves() / kCanonBlock;
        std::vector<double> bt(nb);
        for (int r = 0; r < P_; ++r) { rc = shards_[r]->block_totals_launch((size_t)-1); if (rc) return shard_fail(r, rc); }
        for (int r = 0; r < P_; ++r) {
            rc = shards_[r]->block_totals_fetch(bt.data());
            if (rc) return shard_fail(r, rc);
            for (size_t b = 0; b < nb; ++b) run += bt[b];
        }
    } else {
        // leaves of all shards in rank order, blocks of 1024 leaves chained, block totals chained (as measure_all_into)
        const size_t nl = shards_[0]->nr_leaves();
        std::vector<double> all((size_t)P_ * nl);
        for (int r = 0; r < P_; ++r) {
            rc = shards_[r]->leaf_totals((size_t)-1, all.data() + (size_t)r * nl);
            if (rc) return shard_fail(r, rc);
        }
        for (size_t b0 = 0; b0 < all.size(); b0 += kCanonBlock) {
            double ib = 0.0;
            const size_t e = std::min(all.size(), b0 + (size_t)kCanonBlock);
            for (size_t b = b0; b < e; ++b) ib += all[b];
            run += ib;
        }
    }
    *out = run;
    return Q1T_OK;
}

int ShardedVectorState::read_amplitudes(size_t offset, size_t len, double *out)
{
    int rc = canonicalize();
    if (rc) return rc;
    const size_t per = (size_t)1 << nl_;
    if (offset + len > ((size_t)1 << n_)) return fail(Q1T_ERR_INVALID_ARGUMENT, "amplitude range out of bounds");
    while (len) {
        const size_t r = offset / per, lo = offset % per, take = std::min(len, per - lo);
        rc = shards_[r]->read_amplitudes(0, lo, take, out);
        if (rc) return shard_fail(r, rc);
        out += 2 * take;
        offset += take;
        len -= take;
    }
    return Q1T_OK;
}

q1t_stats ShardedVectorState::shard_stats(size_t r) const
{
    return shards_[r]->stats;
}

// measure_all_into / peek_all_into (vectorstate.rs:106-161) over the shards: canonical leaf and block totals per shard
// on the device, the chain over blocks continued in rank order on the host, one Uniform(0, total) draw
// per shot (vectorstate.rs:120-133), every shard resolves the draws that fall into its prefix range.
int ShardedVectorState::measure_all_into(const size_t *cbits, size_t ncbits, uint64_t *res, size_t res_len, q1t_rng rng, bool collapse)
{
    if (res_len < shots_) {
        char buf[192];
        std::snprintf(buf, sizeof buf, "Not enough space to store %zu measurement results in array of length %zu", shots_, res_len);
        return fail(Q1T_ERR_NOT_ENOUGH_SPACE, buf);
    }
    if (ncbits != (size_t)n_) {
        char buf[128];
        std::snprintf(buf, sizeof buf, "Expected %d measurement bits, but got %zu", n_, ncbits);
        return fail(Q1T_ERR_INVALID_NR_MEASUREMENT_BITS, buf);
    }
    if (!res || !cbits || !rng.next_u64) return fail(Q1T_ERR_INVALID_ARGUMENT, "NULL pointer argument");
    for (size_t j = 0; j < ncbits; ++j)
        if (cbits[j] >= 64) return fail(Q1T_ERR_INVALID_ARGUMENT, "classical bit index must be < 64");
    for (int r = 0; r < P_; ++r)
        if (shards_[r]->nr_columns() != 1)
            return fail(Q1T_ERR_UNSUPPORTED, "a sharded state that has branched into several columns cannot be measured again");
    int rc = canonicalize();
    if (rc) return rc;
    const size_t nleaves = shards_[0]->nr_leaves();
    const bool blocks = nleaves >= kCanonBlock;
    // the draws do not depend on the totals: generated and sorted while the devices work
    std::vector<double> unit(shots_);
    std::vector<std::vector<double>> bp(P_);           // per shard: weight in front of it, then its inclusive block prefixes
    std::vector<std::vector<double>> Pl(P_);           // small shards: inclusive leaf prefixes, global chain
    std::vector<double> ends(P_), base(P_, 0.0);
    if (blocks) {
        const size_t nb = nleaves / kCanonBlock;
        for (int r = 0; r < P_; ++r) { rc = shards_[r]->block_totals_launch((size_t)-1); if (rc) return shard_fail(r, rc); }
        for (size_t j = 0; j < shots_; ++j) unit[j] = uniform_unit(rng);
        if (rng_failed(rng)) return fail(Q1T_ERR_RNG, "the injected random generator ran out of words");
        std::sort(unit.begin(), unit.end());
        double run = 0.0;
        std::vector<double> bt(nb);
        for (int r = 0; r < P_; ++r) {
            rc = shards_[r]->block_totals_fetch(bt.data());
            if (rc) return shard_fail(r, rc);
            bp[r].resize(nb + 1);
            bp[r][0] = run;
            for (size_t b = 0; b < nb; ++b) { run += bt[b]; bp[r][b + 1] = run; }
            ends[r] = run;
        }
    } else {
        for (size_t j = 0; j < shots_; ++j) unit[j] = uniform_unit(rng);
        if (rng_failed(rng)) return fail(Q1T_ERR_RNG, "the injected random generator ran out of words");
        std::sort(unit.begin(), unit.end());
        // leaves of all shards in rank order, blocks of 1024 leaves chained, block totals chained
        std::vector<double> all((size_t)P_ * nleaves);
        for (int r = 0; r < P_; ++r) {
            rc = shards_[r]->leaf_totals((size_t)-1, all.data() + (size_t)r * nleaves);
            if (rc) return shard_fail(r, rc);
        }
        std::vector<double> Pg(all.size());
        double bprefix = 0.0;
        for (size_t b0 = 0; b0 < all.size(); b0 += kCanonBlock) {
            double ib = 0.0;
            const size_t e = std::min(all.size(), b0 + (size_t)kCanonBlock);
            for (size_t b = b0; b < e; ++b) { ib += all[b]; Pg[b] = bprefix + ib; }
            bprefix += ib;
        }
        for (int r = 0; r < P_; ++r) {
            Pl[r].assign(Pg.begin() + (size_t)r * nleaves, Pg.begin() + (size_t)(r + 1) * nleaves);
            base[r] = r ? Pg[(size_t)r * nleaves - 1] : 0.0;
            ends[r] = Pg[(size_t)(r + 1) * nleaves - 1];
        }
    }
    const double total = ends[P_ - 1];
    const UniformF64 u = uniform_new(0.0, total);
    std::vector<double> chosen(shots_);
    for (size_t j = 0; j < shots_; ++j) chosen[j] = uniform_scale(u, unit[j]);     // monotone: still sorted
    // owner of a draw: the number of rank-end prefixes (all but the last) that are <= chosen
    std::vector<uint64_t> idx(shots_);
    size_t at = 0;
    for (int r = 0; r < P_; ++r) {
        size_t end = at;
        if (r + 1 == P_) end = shots_;
        else while (end < shots_ && !(ends[r] <= chosen[end])) ++end;
        const size_t cnt = end - at;
        if (cnt) {
            if (blocks) rc = shards_[r]->resolve_draws_blocks(0, bp[r].data(), chosen.data() + at, cnt, idx.data() + at);
            else rc = shards_[r]->resolve_draws(0, Pl[r].data(), base[r], chosen.data() + at, cnt, idx.data() + at);
            if (rc) return shard_fail(r, rc);
            for (size_t j = at; j < end; ++j) idx[j] |= (uint64_t)r << nl_;
        }
        at = end;
    }
    // idx is ascending (sorted draws, monotone prefixes): group equal outcomes, route qubit q -> classical bit cbits[q]
    uint64_t mask = 0;
    for (size_t j = 0; j < ncbits; ++j) mask |= 1ull << cbits[j];
    std::vector<uint64_t> vals;
    std::vector<size_t> mult;
    for (size_t j = 0; j < shots_; ++j) {
        if (!vals.empty() && vals.back() == idx[j]) { ++mult.back(); continue; }
        vals.push_back(idx[j]);
        mult.push_back(1);
    }
    size_t s = 0;
    for (size_t v = 0; v < vals.size(); ++v) {
        uint64_t word = 0;
        for (int q = 0; q < n_; ++q) word |= ((vals[v] >> (n_ - 1 - q)) & 1ull) << cbits[q];
        for (size_t c = 0; c < mult[v]; ++c, ++s) res[s] = (res[s] & ~mask) | word;
    }
    if (collapse) {
        // every column is a basis state held by exactly one shard (vectorstate.rs:150-158 without the dense matrix)
        const uint64_t lm = ((uint64_t)1 << nl_) - 1;
        std::vector<uint64_t> local(vals.size());
        for (int r = 0; r < P_; ++r) {
            for (size_t v = 0; v < vals.size(); ++v) local[v] = (vals[v] >> nl_) == (uint64_t)r ? (vals[v] & lm) : UINT64_MAX;
            rc = shards_[r]->replace_columns(vals.size(), local.data(), mult.data());
            if (rc) return shard_fail(r, rc);
        }
    }
    return Q1T_OK;
}

}  // namespace q1t
