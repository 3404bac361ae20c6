// engine.h -- DeviceVectorState: the HBM-resident replacement of the reference's
// `VectorState` (vectorstate.rs:25-415), one method per `QuState` trait method.
#pragma once
#include <cuda_runtime.h>

#include <complex>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/q1t_engine.h"
#include "kernels.h"
#include "planner.h"

namespace q1t {

// one branch column (vectorstate.rs:31-34: `counts[c]` shots share `states[.., c]`)
struct Column {
    double2 *buf = nullptr;     // 2^n amplitudes in HBM, or null for a lazy basis state
    bool basis = false;         // unit vector |basis_idx> kept as an index (measure_all collapse)
    uint64_t basis_idx = 0;
    size_t count = 0;
};

class DeviceVectorState {
public:
    DeviceVectorState(size_t nr_bits, size_t nr_shots, int device);
    ~DeviceVectorState();
    int init_zero_state();
    int init_from_qubit_coefs(const double *coefs);
    int set_product_state(const double *coefs);

    int apply_gate(const double *mat, size_t dim, const size_t *bits, size_t k, const char *desc);
    int apply_unary_gate_all(const double *mat, size_t dim, const char *desc);
    int apply_lowered(const std::vector<LoweredGate> &lgs);
    void record_lowered(std::vector<LoweredGate> *rec) { lowered_rec_ = rec; }
    bool layout_is_identity() const
    {
        for (int l = 0; l < n_; ++l)
            if (perm_[l] != l) return false;
        return queue_cols_.empty();
    }
    int apply_conditional_gate(const uint8_t *control, size_t ncontrol, const double *mat, size_t dim,
                               const size_t *bits, size_t k, const char *desc);
    int measure_into(size_t qbit, size_t cbit, uint64_t *res, size_t res_len, q1t_rng rng, bool collapse);
    int measure_all_into(const size_t *cbits, size_t ncbits, uint64_t *res, size_t res_len, q1t_rng rng, bool collapse);
    int reset(size_t bit, q1t_rng rng);
    int reset_all();

    // shard primitives (multi-GPU composition)
    int init_empty();
    size_t nr_leaves() const;
    int leaf_totals(size_t qbit, double *out);
    int resolve_draws(size_t col, const double *P, double base, const double *chosen, size_t nd, uint64_t *idx);
    int block_totals(size_t qbit, double *out);
    int block_totals_launch(size_t qbit);
    int block_totals_fetch(double *out);
    int device() const { return device_; }
    int resolve_draws_blocks(size_t col, const double *bp, const double *chosen, size_t nd, uint64_t *idx);
    int scale_split_columns(const double *f0, const double *f1, const size_t *n0s);
    int collapse_columns(size_t qbit, const double *w0s, const size_t *n0s);
    int replace_columns(size_t ncols, const uint64_t *idx, const size_t *counts);
    int column_ptr(size_t col, void **ptr);
    int ipc_export(size_t col, unsigned char *handle64);
    int peer_swap(size_t col, const unsigned char *peer_handle64, size_t local_qubit, int my_bit);

    // peer group (one shard per GPU): two registered shard buffers and a mailbox per rank, mapped by every peer
    // once; barrier and multi-bit remap are stream-ordered device work (kernels.cu group_*_kernel)
    int scale_all(double re, double im);          // every column times a scalar (a rank's share of a gate on a pinned rank bit)
    int group_export(unsigned char *handles3x64, void **ptrs3);
    int group_open(size_t P, size_t rank, const unsigned char *all_handles, void *const *all_ptrs);
    int group_barrier();
    int group_remap(size_t k, const int *rank_bits, const size_t *local_qubits);
    int group_remap_prepare();
    int group_remap_issue(size_t k, const int *rank_bits, const size_t *local_qubits);
    int group_close();

    int counts(size_t *out);
    size_t nr_columns() { return cols_.size(); }
    int read_amplitudes(size_t col, size_t offset, size_t len, double *out);
    int write_amplitudes(size_t col, size_t offset, size_t len, const double *in);
    int marginal0(size_t qbit, double *w0_out);
    int column_totals(double *out);
    int flush();                 // run queued gates, restore canonical layout, synchronise
    int flush_async();           // the same, enqueued only
    int set_option(const char *key, long value);

    size_t nr_bits() const { return n_; }
    size_t nr_shots() const { return shots_; }
    const char *last_error() const { return err_.c_str(); }
    q1t_stats stats{};
    bool timing = false;

private:
    int n_;
    size_t shots_;
    int device_;
    cudaStream_t stream_ = nullptr;
    std::vector<Column> cols_;
    std::vector<int> perm_;                  // logical index bit -> physical position
    std::vector<LoweredGate> queue_;         // gates waiting for the next fused flush
    std::vector<int> queue_cols_;            // column subset of the queue (empty = all)
    std::vector<double2 *> free_bufs_;
    std::string err_;
    long tile_bits_ = 12;
    long coalesce_bits_ = 3;                 // low index bits kept contiguous per tile (3 = 128 B, 2 = 64 B)
    bool sparse_c2_ = true;                  // batches that start from basis states: 64-byte tiles (10 free bits per sweep)
    bool fuse_leaf_totals_ = true;           // measure_all: leaf totals produced by the last sweep's store pass
    bool want_leaf_fusion_ = false, leaf_fused_ = false;   // handshake between measure_all_into and run_sweeps
    bool track_support_ = true;              // skip what is known to be zero while a batch starts from basis states
    long balance_ = -1;                      // sweep packing: -1 try both, 0 greedy, 1 balanced
    long prefetch_ahead_ = 0;
    bool direct_ = true;
    bool graphs_ = true;                     // launch-bound sweep batches are captured into CUDA graphs and replayed
    uint64_t cur_plan_key_ = 0;              // plan-cache key of the batch run_queue is handing to run_sweeps (0: none)
    bool tma_ = false;                       // dense ladder sweeps load their tiles by TMA (cp.async.bulk.tensor): measured slower than cp.async (profiles/r2_ladder_tma.md), off
    bool fuse_ = true;
    bool no_relabel_ = false;                // conditional gates: Swap must move data, not relabel
    long inplace_relabel_ = 0;               // -1 never, 0 when a second column buffer cannot fit, 1 always
    bool want_inplace_relabel();
    double pending_scale_ = 1.0;             // real factor owed to every column: rides on the next sweep batch's deferred scale
    int apply_pending_scale();

    // device scratch
    double2 **d_colptrs_ = nullptr; size_t colptrs_cap_ = 0;
    PhaseTab *d_ptabs_ = nullptr;
    double *d_leaf_ = nullptr; size_t leaf_cap_ = 0;
    double *d_block_ = nullptr; size_t block_cap_ = 0;
    double *d_totals_ = nullptr; size_t totals_cap_ = 0;
    double *d_chosen_ = nullptr; uint64_t *d_idx_ = nullptr; size_t draws_cap_ = 0;
    double2 *d_mat_ = nullptr;
    size_t mat_cap_ = 0;                     // complex entries d_mat_ holds (grows for dense blocks on more than 6 targets)
    double2 **d_pair_ = nullptr;
    unsigned long long *d_gen_ = nullptr; size_t gen_cap_ = 0;
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    struct PeerGroup {
        bool exported = false, open = false;
        int P = 0, rank = 0;
        double2 *bufs[2] = { nullptr, nullptr };        // this rank's registered shard buffers (never freed while exported)
        unsigned long long *mail = nullptr;             // this rank's mailbox
        std::vector<void *> ipc_opened;                 // mappings to close
        void **d_peer_buf = nullptr;                    // [P][2]
        unsigned long long **d_peer_mail = nullptr;     // [P]
        unsigned long long epoch = 0;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        bool timing_pending = false;
        bool pending = false;                           // option fused_remap: a recorded trade the next sweep will read through
        GroupRemapArgs pend;
    } grp_;
    int group_remap_run(const GroupRemapArgs &a);       // barrier, swap pass, barrier
    int resolve_pending_remap();                        // a recorded trade nobody could fuse: run it as a swap pass now
    bool fused_remap_ = false;
    long mid_relabel_ = 1;                   // dense batches: relabelling stores in the middle of a plan (planner.h); 0 off, 1: states of >= 2^24 (measured: dense QFT-30 sweeps 7.23 -> 6.84 ms), 2: any size
    void group_collect_timing();

    int fail(int code, const std::string &msg) { err_ = msg; return code; }
    int cuda_fail(cudaError_t e, const char *what);
    int ensure_device();
    int alloc_column(double2 **out);
    void release_column(double2 *p);
    int materialize(Column &c);
    int upload_colptrs(const std::vector<int> &which);
    int reserve_colptrs(size_t need);
    int issue_sweeps(std::vector<PlannedSweep> &sweeps, const std::vector<int> &which, bool final_relabel, bool generate,
                     const std::vector<unsigned long long> &gen, bool ident);
    int run_queue(bool final_relabel = false);   // queue -> planner -> sweep launches
    int run_sweeps(std::vector<PlannedSweep> &sweeps, const std::vector<int> &which, bool final_relabel);
    int run_generic(const LoweredGate &g, const std::vector<int> &which);
    int canonicalize();                      // undo swap relabelling (perm_ -> identity)
    int reduce_launch(uint64_t mask, uint64_t want, std::vector<int> &dev_cols, bool leaf_totals_ready);
    int reduce_fetch(std::vector<double> &totals, size_t ndev);
    int reduce_columns(uint64_t mask, uint64_t want, std::vector<double> &totals, std::vector<int> &dev_cols,
                       bool leaf_totals_ready = false);
    int ensure_scratch(size_t ncols);
    int lower_and_queue(const double *mat, size_t dim, const size_t *bits, size_t k, const char *desc);
    int enqueue_lowered(const LoweredGate &lg);
    std::vector<LoweredGate> *lowered_rec_ = nullptr;
    void time_begin();
    void time_end(double &acc);
};

// host sampling helpers (sampling.cpp): rand 0.7 Uniform / rand_distr 0.2 Binomial restated
struct UniformF64 { double low, scale; };
UniformF64 uniform_new(double low, double high);
double uniform_sample(const UniformF64 &u, q1t_rng rng);
double uniform_unit(q1t_rng rng);                          // uniform_sample == uniform_scale(u, uniform_unit(rng))
double uniform_scale(const UniformF64 &u, double v01);
uint64_t binomial_sample(q1t_rng rng, uint64_t n, double p);
bool rng_failed(q1t_rng rng);    // true if a built-in injected generator ran dry

// gate table (gates.cpp)
int builtin_gate_matrix(const char *name, const double *params, size_t nparams, std::complex<double> *out);

}  // namespace q1t
