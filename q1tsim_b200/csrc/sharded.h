// sharded.h -- ShardedVectorState: a state too large for one GPU, sharded by its top log2(P) index bits over P
// devices of ONE process.  The C++ form of q1tsim_b200/sharded.py for what `Circuit::execute` needs of a
// state beyond one GPU: gates, measure_all / peek_all, reset_all, read-out.  Same mechanisms: replicated start (rank
// bits pinned to a basis value cost no exchange), blocks of gate matrices selected by rank bits, multi-bit qubit remaps
// over peer memory (q1t_group_*), canonical reductions chained in rank order.  Every shard is a DeviceVectorState; the
// devices may repeat (several shards on one GPU: how the path is tested on a one-GPU box).
#pragma once
#include <complex>
#include <memory>
#include <string>
#include <vector>

#include "engine.h"

namespace q1t {

class ShardedVectorState {
public:
    // devices.size() must be a power of two >= 2; nr_bits - log2(P) >= 10.  `dry` builds no shards: the layout
    // bookkeeping alone, to count what an op list would cost (remaps) from a given initial layout.
    ShardedVectorState(size_t nr_bits, size_t nr_shots, const std::vector<int> &devices, bool dry = false);
    ~ShardedVectorState();
    int init_zero_state();                          // vectorstate.rs:41-53
    // a run from |0..0> may start from any layout: where0[q] = {global?, rank bit | local engine qubit} (see sharded.py)
    int set_initial_layout(const std::vector<int> &dest);
    int apply_gate(const double *mat, size_t dim, const size_t *bits, size_t k, const char *desc);      // vectorstate.rs:166-178
    int apply_unary_gate_all(const double *mat, size_t dim, const char *desc);                          // vectorstate.rs:180-189
    int measure_all_into(const size_t *cbits, size_t ncbits, uint64_t *res, size_t res_len, q1t_rng rng, bool collapse);   // :106-161
    int reset_all();                                // vectorstate.rs:410-415
    int canonicalize();
    int read_amplitudes(size_t offset, size_t len, double *out);      // canonical index order
    int column_total(double *out);
    size_t nr_bits() const { return n_; }
    size_t nr_shots() const { return shots_; }
    size_t nr_shards() const { return P_; }
    const char *last_error() const { return err_.c_str(); }
    uint64_t remaps = 0, exchanges = 0, local_relabels = 0;
    q1t_stats shard_stats(size_t r) const;

private:
    struct Where { bool global; int idx; };         // rank bit, or local engine qubit (0 = top index bit of the shard)
    int n_, g_, P_, nl_;
    size_t shots_;
    bool dry_;
    std::vector<int> devices_;
    std::vector<std::unique_ptr<DeviceVectorState>> shards_;
    std::vector<Where> where_;
    std::vector<int> pin_;                          // per rank bit: -1 free, else the basis value it is pinned to
    std::string err_;

    int fail(int code, const std::string &m) { err_ = m; return code; }
    int shard_fail(size_t r, int rc) { err_ = shards_[r]->last_error(); return rc; }
    Where canonical(int q) const { return q < g_ ? Where{ true, g_ - 1 - q } : Where{ false, q - g_ }; }
    int qubit_at(bool global, int idx) const;
    int rank_bit(int r, int i) const { return (r >> i) & 1; }
    int depin(int i);
    int depin_all();
    int bring_local(int q, const std::vector<int> &keep);
    int exchange_multi(const std::vector<std::pair<int, int>> &trades);       // (rank bit, logical qubit that is local)
    int local_swap(int a, int b);                                             // engine-level relabel on every shard
    int open_group();
};

}  // namespace q1t
