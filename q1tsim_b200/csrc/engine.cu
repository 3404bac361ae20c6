// engine.cu -- DeviceVectorState: HBM-resident multi-column run state.
//
// Mirrors `impl QuState for VectorState` (vectorstate.rs:164-416).  Storage is
// one contiguous buffer of 2^n complex128 per branch column (cmatrix.rs moved
// into HBM; the reference interleaves columns in an (N, C) row-major matrix,
// which would re-stride every amplitude whenever a measurement splits a
// column).  Gates are queued and executed by the fused sweep kernel at the next
// point where the state is observed.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace q1t {

#define CK(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// every host<->device copy of the engine goes through here: q1t_stats::h2d_bytes / d2h_bytes are what bench.py reports
// as the end-to-end transfer volume
// While a sweep batch is being captured into a CUDA graph, the host sources of its copies are moved into pinned
// memory owned by the graph: a replay re-reads them from there, so they must neither move nor change.
struct GraphArena {
    std::vector<void *> chunks;
    void *stage(const void *src, size_t nbytes)
    {
        void *p = nullptr;
        if (cudaMallocHost(&p, nbytes ? nbytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        std::memcpy(p, src, nbytes);
        chunks.push_back(p);
        return p;
    }
    // programs and phase tables of the captured sweeps live in ONE device blob that stays resident as long as the
    // graph does (uploaded once, after the capture): the kernels read them from there, a replay uploads nothing
    char *d_blob = nullptr, *h_blob = nullptr;
    size_t blob_cap = 0, blob_used = 0;
    bool reserve_blob(size_t nbytes)
    {
        if (cudaMalloc(&d_blob, nbytes) != cudaSuccess) { cudaGetLastError(); d_blob = nullptr; return false; }
        if (cudaMallocHost(&h_blob, nbytes) != cudaSuccess) { cudaGetLastError(); cudaFree(d_blob); d_blob = nullptr; h_blob = nullptr; return false; }
        blob_cap = nbytes;
        blob_used = 0;
        return true;
    }
    // returns the DEVICE address the bytes will have once the blob is uploaded
    const void *put(const void *src, size_t nbytes)
    {
        const size_t at = (blob_used + 255) & ~(size_t)255;
        if (!d_blob || at + nbytes > blob_cap) return nullptr;
        std::memcpy(h_blob + at, src, nbytes);
        blob_used = at + nbytes;
        return d_blob + at;
    }
    void release()
    {
        for (void *p : chunks) cudaFreeHost(p);
        chunks.clear();
        if (h_blob) cudaFreeHost(h_blob);
        if (d_blob) cudaFree(d_blob);
        h_blob = d_blob = nullptr;
    }
};
static thread_local GraphArena *g_capture_arena = nullptr;      // non-null while this thread captures a sweep batch

#define cudaMemcpyAsync(dst, src, nbytes, kind, stream) counted_copy(stats, dst, src, nbytes, kind, stream)
static inline cudaError_t counted_copy(q1t_stats &st, void *dst, const void *src, size_t nbytes, cudaMemcpyKind kind, cudaStream_t stream)
{
    if (kind == cudaMemcpyHostToDevice) st.h2d_bytes += nbytes;
    else if (kind == cudaMemcpyDeviceToHost) st.d2h_bytes += nbytes;
    if (g_capture_arena && kind == cudaMemcpyHostToDevice) {
        src = g_capture_arena->stage(src, nbytes);
        if (!src) return cudaErrorMemoryAllocation;
    }
    return (cudaMemcpyAsync)(dst, src, nbytes, kind, stream);
}

// ---------------------------------------------------------------------------
// process-wide cache of column buffers: execute() builds a fresh state every
// call (circuit.rs:594-600), so without a cache every call would pay
// cudaMalloc + cudaFree of 16 B * 2^n
// ---------------------------------------------------------------------------
namespace {
struct PoolEntry { int device; size_t bytes; void *ptr; };
std::mutex g_pool_mu;
std::vector<PoolEntry> g_pool;
const size_t kPoolMaxPerClass = 4;

void *pool_take(int device, size_t bytes)
{
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (size_t i = 0; i < g_pool.size(); ++i)
        if (g_pool[i].device == device && g_pool[i].bytes == bytes) {
            void *p = g_pool[i].ptr;
            g_pool.erase(g_pool.begin() + i);
            return p;
        }
    return nullptr;
}
// returns false if the caller must cudaFree the buffer itself.  Budget: at most kPoolMaxPerClass
// buffers per size, and never more than 72 % of the device memory parked in the cache.
bool pool_give(int device, size_t bytes, void *ptr)
{
    std::lock_guard<std::mutex> lk(g_pool_mu);
    size_t same = 0, cached = 0;
    for (const PoolEntry &e : g_pool)
        if (e.device == device) { cached += e.bytes; if (e.bytes == bytes) ++same; }
    if (same >= kPoolMaxPerClass) return false;
    static size_t total_of[64] = { 0 };      // cudaMemGetInfo costs up to a millisecond: ask once per device
    size_t total_b = (device >= 0 && device < 64) ? total_of[device] : 0;
    if (!total_b) {
        size_t free_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); total_b = 0; }
        if (device >= 0 && device < 64) total_of[device] = total_b;
    }
    if (total_b && cached + bytes > total_b / 100 * 72) return false;
    g_pool.push_back({ device, bytes, ptr });
    return true;
}
// free every cached buffer of a device (called before reporting out-of-memory)
void pool_flush(int device)
{
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (size_t i = 0; i < g_pool.size();) {
        if (g_pool[i].device == device) { cudaFree(g_pool[i].ptr); g_pool.erase(g_pool.begin() + i); }
        else ++i;
    }
}
// small scratch allocations go through the same cache (exact-size classes)
cudaError_t scratch_alloc(int device, void **out, size_t bytes)
{
    if (void *p = pool_take(device, bytes)) { *out = p; return cudaSuccess; }
    return cudaMalloc(out, bytes);
}
void scratch_free(int device, void *p, size_t bytes)
{
    if (!p) return;
    if (!pool_give(device, bytes, p)) cudaFree(p);
}
}  // namespace

namespace {
struct PlanCacheEntry { uint64_t key, key2; size_t ngates; uint64_t stamp; std::vector<PlannedSweep> sweeps; };
std::mutex g_plan_mu;
std::vector<PlanCacheEntry> g_plan_cache;
uint64_t g_plan_clock = 0;
const size_t kPlanCacheMax = 16;
}  // namespace

// process-wide cache of captured sweep batches (launch-bound circuits: a hundred 10-microsecond sweeps per execute())
namespace {
struct GraphEntry { uint64_t key; cudaGraphExec_t exec; GraphArena arena; uint64_t stamp; uint64_t launches, sweeps, col_passes, bytes, h2d; };
std::mutex g_graph_mu;
std::vector<GraphEntry> g_graphs;
uint64_t g_graph_clock = 0;
const size_t kGraphCacheMax = 12;
}  // namespace

int DeviceVectorState::cuda_fail(cudaError_t e, const char *what)
{
    char buf[512];
    std::snprintf(buf, sizeof buf, "CUDA error: %s (%s)", cudaGetErrorString(e), what);
    cudaGetLastError();
    return fail(Q1T_ERR_CUDA, buf);
}

DeviceVectorState::DeviceVectorState(size_t nr_bits, size_t nr_shots, int device)
    : n_((int)nr_bits), shots_(nr_shots), device_(device)
{
    perm_.resize(n_);
    for (int i = 0; i < n_; ++i) perm_[i] = i;
    // A/B switches for whole-process experiments (bench.py, tools/): same as set_option()
    if (const char *e = std::getenv("Q1T_COALESCE_BITS")) { const long v = std::atol(e); if (v == 2 || v == 3) coalesce_bits_ = v; }
    if (const char *e = std::getenv("Q1T_BALANCE")) balance_ = std::atol(e);
    if (const char *e = std::getenv("Q1T_TRACK_SUPPORT")) track_support_ = std::atol(e) != 0;
    if (const char *e = std::getenv("Q1T_SPARSE_C2")) sparse_c2_ = std::atol(e) != 0;
    if (const char *e = std::getenv("Q1T_FUSE_LEAF")) fuse_leaf_totals_ = std::atol(e) != 0;
    if (const char *e = std::getenv("Q1T_INPLACE_RELABEL")) inplace_relabel_ = std::atol(e);
    if (const char *e = std::getenv("Q1T_TMA")) tma_ = std::atol(e) != 0;
    if (const char *e = std::getenv("Q1T_GRAPHS")) graphs_ = std::atol(e) != 0;
    if (const char *e = std::getenv("Q1T_PREFETCH_AHEAD")) prefetch_ahead_ = std::atol(e);
    if (const char *e = std::getenv("Q1T_TILE_BITS")) { const long v = std::atol(e); if (v >= 8 && v <= kMaxTileBits) tile_bits_ = v; }
    if (const char *e = std::getenv("Q1T_FUSED_REMAP")) fused_remap_ = std::atol(e) != 0;
    if (const char *e = std::getenv("Q1T_MID_RELABEL")) mid_relabel_ = std::atol(e);
}

DeviceVectorState::~DeviceVectorState()
{
    if (stream_) {
        cudaSetDevice(device_);
        cudaStreamSynchronize(stream_);
        const size_t bytes = sizeof(double2) << n_;
        group_close();
        // buffers that were exported to peers are freed, never pooled: a pooled buffer could be handed to another
        // state while a peer still holds a mapping of it
        auto exported = [&](double2 *p) { return grp_.exported && (p == grp_.bufs[0] || p == grp_.bufs[1]); };
        for (Column &c : cols_)
            if (c.buf && (exported(c.buf) || !pool_give(device_, bytes, c.buf))) cudaFree(c.buf);
        for (double2 *p : free_bufs_)
            if (exported(p) || !pool_give(device_, bytes, p)) cudaFree(p);
        if (grp_.mail) cudaFree(grp_.mail);
        if (grp_.ev0) cudaEventDestroy(grp_.ev0);
        if (grp_.ev1) cudaEventDestroy(grp_.ev1);
        scratch_free(device_, d_colptrs_, sizeof(double2 *) * colptrs_cap_);
        scratch_free(device_, d_ptabs_, sizeof(PhaseTab) * kMaxPhase);
        scratch_free(device_, d_leaf_, sizeof(double) * leaf_cap_);
        scratch_free(device_, d_block_, sizeof(double) * block_cap_);
        scratch_free(device_, d_totals_, sizeof(double) * totals_cap_);
        scratch_free(device_, d_chosen_, sizeof(double) * draws_cap_);
        scratch_free(device_, d_idx_, sizeof(uint64_t) * draws_cap_);
        scratch_free(device_, d_mat_, sizeof(double2) * mat_cap_);
        scratch_free(device_, d_pair_, sizeof(double2 *) * 2);
        scratch_free(device_, d_gen_, sizeof(unsigned long long) * gen_cap_);
        if (ev0_) cudaEventDestroy(ev0_);
        if (ev1_) cudaEventDestroy(ev1_);
        cudaStreamDestroy(stream_);
        cprog_stream_unregister(device_);
    }
}

int DeviceVectorState::ensure_device()
{
    if (stream_) { CK(cudaSetDevice(device_)); return Q1T_OK; }
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        return fail(Q1T_ERR_CUDA, "no CUDA device available: the q1tsim B200 engine has no CPU fallback");
    }
    if (device_ < 0 || device_ >= cnt) return fail(Q1T_ERR_CUDA, "invalid CUDA device ordinal");
    CK(cudaSetDevice(device_));
    CK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    cprog_stream_register(device_);
    CK(cudaEventCreate(&ev0_));
    CK(cudaEventCreate(&ev1_));
    CK(scratch_alloc(device_, (void **)&d_ptabs_, sizeof(PhaseTab) * kMaxPhase));
    CK(scratch_alloc(device_, (void **)&d_mat_, sizeof(double2) << (2 * kMaxGenericBits)));
    mat_cap_ = (size_t)1 << (2 * kMaxGenericBits);
    CK(scratch_alloc(device_, (void **)&d_pair_, sizeof(double2 *) * 2));
    return Q1T_OK;
}

int DeviceVectorState::alloc_column(double2 **out)
{
    if (!free_bufs_.empty()) { *out = free_bufs_.back(); free_bufs_.pop_back(); return Q1T_OK; }
    const size_t bytes = sizeof(double2) << n_;
    if (void *p = pool_take(device_, bytes)) { *out = static_cast<double2 *>(p); return Q1T_OK; }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaStreamSynchronize(stream_);
        pool_flush(device_);
        e = cudaMalloc(out, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        char buf[256];
        std::snprintf(buf, sizeof buf, "out of device memory allocating a branch column of %zu bytes (%zu columns live)",
                      sizeof(double2) << n_, cols_.size());
        return fail(Q1T_ERR_CUDA, buf);
    }
    return Q1T_OK;
}

void DeviceVectorState::release_column(double2 *p)
{
    if (!p) return;
    // keep spare buffers (relabel scratch + one split target) as long as live + spare columns stay
    // within ~80 % of the device memory; large shards keep exactly one spare
    size_t live = 0;
    for (const Column &c : cols_)
        if (c.buf) ++live;
    const size_t bytes = sizeof(double2) << n_;
    static size_t total_of[64] = { 0 };      // cudaMemGetInfo costs up to a millisecond: ask once per device
    size_t total_b = (device_ >= 0 && device_ < 64) ? total_of[device_] : 0;
    if (!total_b) {
        size_t free_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); total_b = 0; }
        if (device_ >= 0 && device_ < 64) total_of[device_] = total_b;
    }
    const bool fits = total_b == 0 || (live + free_bufs_.size() + 1) * bytes <= total_b / 100 * 80;
    const bool registered = grp_.exported && (p == grp_.bufs[0] || p == grp_.bufs[1]);    // peers have it mapped: keep
    if (registered || (free_bufs_.size() < 2 && fits)) free_bufs_.push_back(p);
    else { cudaStreamSynchronize(stream_); cudaFree(p); }
}

int DeviceVectorState::init_zero_state()
{
    { const int rcp = resolve_pending_remap(); if (rcp) return rcp; }      // a recorded qubit remap: the peers still read this shard through it
    int rc = ensure_device();
    if (rc) return rc;
    // |0..0> is kept as a lazy basis column: the first fused sweep synthesises it on the fly
    // instead of paying a memset (one full write) plus a read of 2^n zeros
    Column c;
    c.count = shots_;
    c.basis = true;
    c.basis_idx = 0;
    cols_.push_back(c);
    return Q1T_OK;
}

// vectorstate.rs:62-83
int DeviceVectorState::init_from_qubit_coefs(const double *coefs)
{
    { const int rcp = resolve_pending_remap(); if (rcp) return rcp; }      // a recorded qubit remap: the peers still read this shard through it
    int rc = ensure_device();
    if (rc) return rc;
    std::vector<double2> nc(2 * (size_t)n_);
    for (int q = 0; q < n_; ++q) {
        const double *c = coefs + 4 * q;
        const double norm = std::sqrt((c[0] * c[0] + c[1] * c[1]) + (c[2] * c[2] + c[3] * c[3]));
        nc[2 * q] = make_double2(c[0] / norm, c[1] / norm);
        nc[2 * q + 1] = make_double2(c[2] / norm, c[3] / norm);
    }
    Column c;
    rc = alloc_column(&c.buf);
    if (rc) return rc;
    c.count = shots_;
    double2 *d_coefs = nullptr;
    CK(cudaMalloc(&d_coefs, sizeof(double2) * nc.size() + 16));
    CK(cudaMemcpyAsync(d_coefs, nc.data(), sizeof(double2) * nc.size(), cudaMemcpyHostToDevice, stream_));
    CK(launch_product_state(c.buf, n_, d_coefs, stream_));
    stats.kernel_launches++;
    CK(cudaStreamSynchronize(stream_));
    cudaFree(d_coefs);
    cols_.push_back(c);
    return Q1T_OK;
}

static unsigned long long physical_index(uint64_t logical, const std::vector<int> &perm);

// A lazy basis column is kept by its LOGICAL index; the buffer is laid out in the current physical order
// (a pending zero-byte Swap relabelling makes the two differ), so the 1 goes to the physical position.
// the state becomes the product state of `coefs` in place (its column buffers stay the ones it has)
int DeviceVectorState::set_product_state(const double *coefs)
{
    { const int rcp = resolve_pending_remap(); if (rcp) return rcp; }      // a recorded qubit remap: the peers still read this shard through it
    int rc = ensure_device();
    if (rc) return rc;
    queue_.clear();
    queue_cols_.clear();
    pending_scale_ = 1.0;
    CK(cudaStreamSynchronize(stream_));
    for (Column &c : cols_)
        if (c.buf) release_column(c.buf);
    cols_.clear();
    for (int l = 0; l < n_; ++l) perm_[l] = l;
    return init_from_qubit_coefs(coefs);
}

int DeviceVectorState::materialize(Column &c)
{
    if (!c.basis) return Q1T_OK;
    int rc = alloc_column(&c.buf);
    if (rc) return rc;
    CK(cudaMemsetAsync(c.buf, 0, sizeof(double2) << n_, stream_));
    if (c.basis_idx != UINT64_MAX) {        // UINT64_MAX: a lazy all-zero column (shard that does not hold the basis state)
        CK(launch_set_basis(c.buf, physical_index(c.basis_idx, perm_), stream_));
        stats.kernel_launches++;
    }
    c.basis = false;
    return Q1T_OK;
}

int DeviceVectorState::reserve_colptrs(size_t need)
{
    if (need > colptrs_cap_) {
        if (d_colptrs_) { CK(cudaStreamSynchronize(stream_)); scratch_free(device_, d_colptrs_, sizeof(double2 *) * colptrs_cap_); }
        colptrs_cap_ = std::max<size_t>(need * 2, 16);
        CK(scratch_alloc(device_, (void **)&d_colptrs_, sizeof(double2 *) * colptrs_cap_));
    }
    return Q1T_OK;
}

int DeviceVectorState::upload_colptrs(const std::vector<int> &which)
{
    const size_t need = which.size();
    int rc0 = reserve_colptrs(need);
    if (rc0) return rc0;
    std::vector<double2 *> h(need);
    for (size_t i = 0; i < need; ++i) h[i] = cols_[which[i]].buf;
    CK(cudaMemcpyAsync(d_colptrs_, h.data(), sizeof(double2 *) * need, cudaMemcpyHostToDevice, stream_));
    return Q1T_OK;
}

void DeviceVectorState::time_begin()
{
    if (timing) cudaEventRecord(ev0_, stream_);
}
void DeviceVectorState::time_end(double &acc)
{
    if (!timing) return;
    cudaEventRecord(ev1_, stream_);
    cudaEventSynchronize(ev1_);
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0_, ev1_);
    acc += ms;
    static const bool log_each = std::getenv("Q1T_SWEEP_LOG") != nullptr;      // tools/: per-launch times on stderr
    if (log_each) std::fprintf(stderr, "q1t launch %.4f ms\n", ms);
}

// ---------------------------------------------------------------------------
// gate queue
// ---------------------------------------------------------------------------
int DeviceVectorState::lower_and_queue(const double *mat, size_t dim, const size_t *bits, size_t k, const char *desc)
{
    if (!mat || (!bits && k)) return fail(Q1T_ERR_INVALID_ARGUMENT, "NULL pointer argument");
    // vectorstate.rs:169-174: gate.nr_affected_bits() != bits.len()
    size_t gate_bits = 0;
    while (((size_t)1 << gate_bits) < dim) ++gate_bits;
    if (((size_t)1 << gate_bits) != dim || gate_bits != k) {
        char buf[256];
        std::snprintf(buf, sizeof buf, "Expected %zu bits for \"%s\", got %zu", gate_bits, desc ? desc : "gate", k);
        return fail(Q1T_ERR_INVALID_NR_BITS, buf);
    }
    if (k == 0) return Q1T_OK;
    if (k > 12) return fail(Q1T_ERR_UNSUPPORTED, "gates on more than 12 qubits are not supported");
    int phys[16];
    for (size_t j = 0; j < k; ++j) {
        if (bits[j] >= (size_t)n_) {
            char buf[128];
            std::snprintf(buf, sizeof buf, "Invalid index %zu for a quantum bit", bits[j]);
            return fail(Q1T_ERR_INVALID_QBIT, buf);
        }
        phys[j] = perm_[n_ - 1 - (int)bits[j]];
    }
    LoweredGate lg;
    std::string e;
    if (!lower_gate(reinterpret_cast<const cplx *>(mat), (int)k, phys, lg, e))
        return fail(e.find("duplicate") != std::string::npos ? Q1T_ERR_INVALID_ARGUMENT : Q1T_ERR_UNSUPPORTED, e);
    if (lowered_rec_) lowered_rec_->push_back(lg);
    return enqueue_lowered(lg);
}

// a gate already lowered onto physical positions: relabel (Swap), drop (identity) or queue
int DeviceVectorState::enqueue_lowered(const LoweredGate &lg)
{
    stats.gates_queued++;
    if (lg.kind == LoweredGate::POLY && lg.nb == 0) return Q1T_OK;          // identity
    if (lg.kind == LoweredGate::SWAP && n_ >= 5 && fuse_ && !no_relabel_) {
        // zero-byte relabel: exchange the physical homes of the two logical bits
        int la = -1, lb = -1;
        for (int l = 0; l < n_; ++l) {
            if (perm_[l] == lg.b[0]) { perm_[l] = lg.b[1]; la = l; }
            else if (perm_[l] == lg.b[1]) { perm_[l] = lg.b[0]; lb = l; }
        }
        // a lazy basis column is kept as a LOGICAL index: Swap|idx> exchanges the two bits there,
        // which leaves its physical position (what the queued gates will act on) unchanged
        if (la >= 0 && lb >= 0)
            for (Column &c : cols_)
                if (c.basis && c.basis_idx != UINT64_MAX && (((c.basis_idx >> la) ^ (c.basis_idx >> lb)) & 1ull))
                    c.basis_idx ^= (1ull << la) | (1ull << lb);
        return Q1T_OK;
    }
    queue_.push_back(lg);
    return Q1T_OK;
}

static void to_generic(const LoweredGate &g, LoweredGate &out)
{
    out = g;
    if (g.kind == LoweredGate::GENERIC) return;
    out.kind = LoweredGate::GENERIC;
    out.pos.clear();
    out.mat.clear();
    if (g.kind == LoweredGate::G1) {
        out.pos.push_back(g.target);
        for (int e = 0; e < 4; ++e) out.mat.push_back(cplx(g.m[2 * e], g.m[2 * e + 1]));
    } else if (g.kind == LoweredGate::SWAP) {
        out.pos.push_back(g.b[0]); out.pos.push_back(g.b[1]);
        static const double sw[16] = { 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1 };
        for (int e = 0; e < 16; ++e) out.mat.push_back(cplx(sw[e], 0));
        out.cmask = 0;
    } else {   // POLY -> diagonal matrix
        const int k = g.nb, G = 1 << k;
        out.cmask = 0;
        for (int j = 0; j < k; ++j) out.pos.push_back(g.b[j]);
        out.mat.assign((size_t)G * G, cplx(0, 0));
        for (int x = 0; x < G; ++x) {
            double a = g.c0;
            if (k == 1) a += x ? g.lin[0] : 0.0;
            else {
                const int x0 = (x >> 1) & 1, x1 = x & 1;
                a += x0 * g.lin[0] + x1 * g.lin[1] + (x0 & x1) * g.quad;
            }
            out.mat[(size_t)x * G + x] = cplx(std::cos(M_PI * a), std::sin(M_PI * a));
        }
    }
}

int DeviceVectorState::run_generic(const LoweredGate &g, const std::vector<int> &which)
{
    const int k = (int)g.pos.size();
    if (k > kMaxGenericBits - 1) {
        // 6..10 targets: groups staged in shared memory, outputs accumulated from the transposed matrix (kernels.cu)
        if (k > kMaxBigGenericBits || k > n_) return fail(Q1T_ERR_UNSUPPORTED, "dense gate blocks on more than 10 target qubits are not supported");
        GenericBigArgs b;
        std::memset(&b, 0, sizeof b);
        b.n = n_; b.k = k; b.cmask = g.cmask;
        std::vector<int> sorted = g.pos;
        std::sort(sorted.begin(), sorted.end());
        for (int j = 0; j < k; ++j) { b.pos[j] = g.pos[j]; b.sorted_pos[j] = sorted[j]; }
        const size_t G = (size_t)1 << k;
        if (G * G > mat_cap_) {
            CK(cudaStreamSynchronize(stream_));
            if (d_mat_) scratch_free(device_, d_mat_, sizeof(double2) * mat_cap_);
            d_mat_ = nullptr;
            CK(scratch_alloc(device_, (void **)&d_mat_, sizeof(double2) * G * G));
            mat_cap_ = G * G;
        }
        std::vector<cplx> mt(G * G);
        for (size_t i = 0; i < G; ++i)
            for (size_t h = 0; h < G; ++h) mt[h * G + i] = g.mat[i * G + h];
        CK(cudaMemcpyAsync(d_mat_, mt.data(), sizeof(double2) * G * G, cudaMemcpyHostToDevice, stream_));
        CK(cudaStreamSynchronize(stream_));                 // (the transposed copy is a local)
        time_begin();
        CK(launch_generic_gate_big(d_colptrs_, (int)which.size(), b, d_mat_, stream_));
        time_end(stats.sweep_ms);
        stats.kernel_launches++;
        stats.sweeps++;
        stats.fallback_sweeps++;
        stats.sweep_column_passes += which.size();
        stats.sweep_bytes += (uint64_t)which.size() * (32ull << n_);
        return Q1T_OK;
    }
    GenericGateArgs a;
    std::memset(&a, 0, sizeof a);
    std::vector<int> sorted = g.pos;
    std::sort(sorted.begin(), sorted.end());
    for (int j = 0; j < k; ++j) a.sorted_pos[j] = sorted[j];
    for (int h = 0; h < (1 << k); ++h) {
        unsigned long long o = 0;
        for (int j = 0; j < k; ++j)
            if ((h >> (k - 1 - j)) & 1) o |= 1ull << g.pos[j];
        a.offs[h] = o;
    }
    a.cmask = g.cmask;
    CK(cudaMemcpyAsync(d_mat_, g.mat.data(), sizeof(double2) * g.mat.size(), cudaMemcpyHostToDevice, stream_));
    time_begin();
    CK(launch_generic_gate(d_colptrs_, (int)which.size(), n_, k, a, d_mat_, stream_));
    time_end(stats.sweep_ms);
    stats.kernel_launches++;
    stats.sweeps++;
    stats.fallback_sweeps++;
    stats.sweep_column_passes += which.size();
    stats.sweep_bytes += (uint64_t)which.size() * (32ull << n_);
    return Q1T_OK;
}

// physical (current layout) index of a logical basis index
static unsigned long long physical_index(uint64_t logical, const std::vector<int> &perm)
{
    unsigned long long p = 0;
    for (size_t l = 0; l < perm.size(); ++l)
        if ((logical >> l) & 1ull) p |= 1ull << perm[l];
    return p;
}

// Launch the planned sweeps on the columns `which`.
//  - if every column is still a lazy basis state, the first sweep generates its input;
//  - if `final_relabel` is set and the qubit relabelling is not the identity, the last
//    sweep also restores the canonical layout on its way out (out-of-place), when its
//    tile allows coalesced stores; otherwise canonicalize() runs a separate sweep later.
int DeviceVectorState::run_sweeps(std::vector<PlannedSweep> &sweeps, const std::vector<int> &which, bool final_relabel)
{
    if (sweeps.empty()) return Q1T_OK;
    bool all_basis = true, any_basis = false;
    for (int c : which) { all_basis = all_basis && cols_[c].basis; any_basis = any_basis || cols_[c].basis; }
    std::vector<unsigned long long> gen(which.size());
    const bool generate = all_basis;
    if (generate) {
        for (size_t i = 0; i < which.size(); ++i) {
            Column &c = cols_[which[i]];
            gen[i] = c.basis_idx == UINT64_MAX ? ~0ull : physical_index(c.basis_idx, perm_);
            int rc = alloc_column(&c.buf);
            if (rc) return rc;
            c.basis = false;
        }
        if (which.size() > gen_cap_) {
            if (d_gen_) { CK(cudaStreamSynchronize(stream_)); scratch_free(device_, d_gen_, sizeof(unsigned long long) * gen_cap_); }
            gen_cap_ = std::max<size_t>(which.size() * 2, 16);
            CK(scratch_alloc(device_, (void **)&d_gen_, sizeof(unsigned long long) * gen_cap_));
        }
    } else {
        for (int c : which) {
            int rc = materialize(cols_[c]);
            if (rc) return rc;
        }
    }
    int rc = reserve_colptrs(which.size());
    if (rc) return rc;
    bool ident = true;
    for (int l = 0; l < n_; ++l)
        if (perm_[l] != l) ident = false;
    // Launch-bound batches (small states, many sweeps) are captured into a CUDA graph once and replayed: a sweep of a
    // 2^20 state runs for ~10 us, its launch plus the upload of its program costs ~40 us of host time.  A batch qualifies
    // if it comes from the plan cache (the key names the gate list), touches no lazy column half-way, and restores no
    // layout (the relabelling path allocates).  The graph is keyed by everything the issued work depends on.
    const bool may_relabel = final_relabel && (!ident || (want_leaf_fusion_ && n_ >= 12 && sweep_uses_ladder_kernel(sweeps.back().prog)));
    bool has_mid = false;
    for (size_t si = 0; si + 1 < sweeps.size(); ++si) has_mid = has_mid || !sweeps[si].mid_dstpos.empty();
    const bool graph_try = graphs_ && !grp_.pending && !has_mid && !timing && cur_plan_key_ != 0 && sweeps.size() >= 4 && n_ <= 26 && !tma_ && !cprog_device_shared(device_) &&
                           (generate || !any_basis) && !may_relabel;
    uint64_t gkey = 0;
    if (graph_try) {
        gkey = 1469598103934665603ull;
        auto mix = [&](const void *p, size_t nbytes) {
            const unsigned char *b = static_cast<const unsigned char *>(p);
            for (size_t i = 0; i < nbytes; ++i) { gkey ^= b[i]; gkey *= 1099511628211ull; }
        };
        const uint64_t hdr[8] = { cur_plan_key_, (uint64_t)sweeps.size(), (uint64_t)generate, (uint64_t)track_support_, (uint64_t)which.size(),
                                  (uint64_t)n_, (uint64_t)device_, (uint64_t)direct_ };
        mix(hdr, sizeof hdr);
        const void *ptrs[3] = { d_ptabs_, d_colptrs_, d_gen_ };
        mix(ptrs, sizeof ptrs);
        for (int c : which) { const void *b = cols_[c].buf; mix(&b, sizeof b); }
        if (generate) mix(gen.data(), sizeof(unsigned long long) * gen.size());
        const double ps = which.size() == cols_.size() ? pending_scale_ : 1.0;
        mix(&ps, sizeof ps);
        cudaGraphExec_t exec = nullptr;
        {
            std::lock_guard<std::mutex> lk(g_graph_mu);
            for (GraphEntry &ge : g_graphs)
                if (ge.key == gkey) {
                    exec = ge.exec;
                    ge.stamp = ++g_graph_clock;
                    stats.kernel_launches += ge.launches;
                    stats.sweeps += ge.sweeps;
                    stats.sweep_column_passes += ge.col_passes;
                    stats.sweep_bytes += ge.bytes;
                    stats.h2d_bytes += ge.h2d;
                }
        }
        if (exec) {
            if (which.size() == cols_.size()) pending_scale_ = 1.0;
            CK(cudaGraphLaunch(exec, stream_));
            stats.graph_replays++;
            sweeps.clear();
            return Q1T_OK;
        }
    }
    if (!graph_try) return issue_sweeps(sweeps, which, final_relabel, generate, gen, ident);
    // capture, instantiate, launch, remember
    GraphEntry ge;
    ge.key = gkey;
    ge.exec = nullptr;
    const q1t_stats before = stats;
    ge.arena.reserve_blob(sweeps.size() * (sizeof(SweepProgram) + 512 + (sizeof(PhaseTab) + 16) * kMaxPhase));    // (without it: uploads stay in the graph)
    CK(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeRelaxed));   // relaxed: the pinned staging chunks are allocated while capturing
    g_capture_arena = &ge.arena;
    rc = issue_sweeps(sweeps, which, final_relabel, generate, gen, ident);
    g_capture_arena = nullptr;
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(stream_, &graph);
    if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        ge.arena.release();
        if (rc) return rc;
        return cuda_fail(ce, "cudaStreamEndCapture");
    }
    ce = cudaGraphInstantiate(&ge.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { ge.arena.release(); return cuda_fail(ce, "cudaGraphInstantiate"); }
    if (ge.arena.blob_used) {
        // the one upload of the batch's programs and phase tables (not part of the graph)
        stats.h2d_bytes += ge.arena.blob_used;
        CK((cudaMemcpyAsync)(ge.arena.d_blob, ge.arena.h_blob, ge.arena.blob_used, cudaMemcpyHostToDevice, stream_));
    }
    CK(cudaGraphLaunch(ge.exec, stream_));
    if (ge.arena.h_blob) {
        CK(cudaStreamSynchronize(stream_));
        cudaFreeHost(ge.arena.h_blob);
        ge.arena.h_blob = nullptr;
    }
    stats.graph_captures++;
    ge.launches = stats.kernel_launches - before.kernel_launches;
    ge.sweeps = stats.sweeps - before.sweeps;
    ge.col_passes = stats.sweep_column_passes - before.sweep_column_passes;
    ge.bytes = stats.sweep_bytes - before.sweep_bytes;
    ge.h2d = stats.h2d_bytes - before.h2d_bytes - ge.arena.blob_used;      // a replay does not upload the resident blob again
    {
        std::lock_guard<std::mutex> lk(g_graph_mu);
        if (g_graphs.size() >= kGraphCacheMax) {
            size_t lru = 0;
            for (size_t i = 1; i < g_graphs.size(); ++i)
                if (g_graphs[i].stamp < g_graphs[lru].stamp) lru = i;
            // (an evicted graph may still be running on some stream: let everything drain before it goes)
            cudaDeviceSynchronize();
            cudaGraphExecDestroy(g_graphs[lru].exec);
            g_graphs[lru].arena.release();
            g_graphs.erase(g_graphs.begin() + lru);
        }
        ge.stamp = ++g_graph_clock;
        g_graphs.push_back(ge);
    }
    return Q1T_OK;
}

// the stream work of a sweep batch (immediately, or into the graph being captured)
int DeviceVectorState::issue_sweeps(std::vector<PlannedSweep> &sweeps, const std::vector<int> &which, bool final_relabel, bool generate,
                                    const std::vector<unsigned long long> &gen, bool ident)
{
    if (generate) CK(cudaMemcpyAsync(d_gen_, gen.data(), sizeof(unsigned long long) * gen.size(), cudaMemcpyHostToDevice, stream_));
    int rc = upload_colptrs(which);
    if (rc) return rc;
    // the deferred Hadamard normalisations of the whole batch are applied once: for free in the
    // generated basis element, otherwise at the store of the last sweep
    {
        double total = 1.0;
        for (PlannedSweep &ps : sweeps) { total *= ps.prog.scale; ps.prog.scale = 1.0; ps.prog.gen_scale = 1.0; }
        if (which.size() == cols_.size()) { total *= pending_scale_; pending_scale_ = 1.0; }
        if (generate) sweeps.front().prog.gen_scale = total;
        else sweeps.back().prog.scale = total;
    }
    for (PlannedSweep &ps : sweeps) {
        ps.prog.prefetch_ahead = (int)prefetch_ahead_;
        if (!direct_) {                       // A/B switch: always stage through shared memory, CTA-wide barriers
            ps.prog.direct_load = ps.prog.direct_store = 0;
            for (int r = 0; r < ps.prog.nrounds; ++r) ps.prog.rounds[r].sync_before = 2;
        }
    }
    // Support tracking.  A batch that starts from basis states |idx> (every circuit does: VectorState::new,
    // vectorstate.rs:41-53) knows that an amplitude is zero unless its index agrees with idx in every bit
    // no non-diagonal gate has touched yet.  Sweeps read, compute and write only what lies inside that
    // support; the last sweep of the batch writes the remaining tiles as zeros, so the columns leave dense.
    bool track = generate && track_support_;
    for (unsigned long long g : gen) track = track && g != ~0ull;
    for (const PlannedSweep &ps : sweeps) track = track && sweep_uses_ladder_kernel(ps.prog);
    uint64_t pinned = n_ >= 64 ? ~0ull : ((1ull << n_) - 1ull);
    std::vector<uint64_t> bytes_moved(sweeps.size());
    for (size_t si = 0; si < sweeps.size(); ++si) {
        SweepProgram &P = sweeps[si].prog;
        const bool last = si + 1 == sweeps.size();
        const bool gen_here = generate && si == 0;
        P.sup_mask = 0;
        P.sup_mode = 0;
        for (int r = 0; r < P.nrounds; ++r) P.rounds[r].zmask = P.rounds[r].smask = 0;
        bytes_moved[si] = (gen_here ? 16ull : 32ull) << n_;
        if (!track || pinned == 0) continue;
        P.sup_mask = pinned;
        P.sup_mode = last ? 2 : 1;
        uint32_t pinned_tl = 0;                         // pinned tile bits, tile-local
        for (int i = 0; i < P.T; ++i)
            if ((pinned >> P.tsrc[i]) & 1ull) pinned_tl |= 1u << i;
        for (int r = 0; r < P.nrounds; ++r) {
            RoundDesc &R = P.rounds[r];
            uint32_t regs = 0;
            for (int j = 0; j < kRegBits; ++j) regs |= 1u << R.reg_tb[j];
            R.zmask = pinned_tl & ~regs;
            R.smask = 0;
            for (int j = 0; j < kRegBits; ++j) R.smask |= ((pinned_tl >> R.reg_tb[j]) & 1u) << j;
            for (int j = kRegBits - R.nsteps; j < kRegBits; ++j) pinned_tl &= ~(1u << R.reg_tb[j]);   // this round's targets
        }
        int pinned_outer = 0, pinned_all = 0;
        for (int i = 0; i < P.n_outer; ++i) pinned_outer += (int)((pinned >> P.osrc[i]) & 1ull);
        for (int b = 0; b < n_; ++b) pinned_all += (int)((pinned >> b) & 1ull);
        const uint64_t tiles_written = last ? (1ull << P.n_outer) : (1ull << (P.n_outer - pinned_outer));
        bytes_moved[si] = ((tiles_written << P.T) + (gen_here ? 0ull : (1ull << (n_ - pinned_all)))) * 16ull;
        pinned &= ~sweeps[si].touched;
    }
    // A recorded qubit remap (option fused_remap, group_remap_issue): the first sweep of the batch reads its tiles through
    // the trade -- from the peers' shards over NVLink, local tiles in between -- instead of waiting for a swap pass.  Needs a
    // dense ladder sweep, the column in one registered buffer and the other registered buffer free to write into.
    double2 *gather_dst = nullptr;
    if (grp_.pending) {
        bool ok = !generate && which.size() == 1 && cols_.size() == 1 && !g_capture_arena && !tma_ && grp_.open && grp_.bufs[0] && grp_.bufs[1] &&
                  sweep_uses_ladder_kernel(sweeps[0].prog) && sweeps[0].prog.nrounds <= 3;
        if (ok) {
            const double2 *cur = cols_[which[0]].buf;
            gather_dst = cur == grp_.bufs[0] ? grp_.bufs[1] : cur == grp_.bufs[1] ? grp_.bufs[0] : nullptr;
            if (std::find(free_bufs_.begin(), free_bufs_.end(), gather_dst) == free_bufs_.end()) gather_dst = nullptr;
        }
        if (!gather_dst) {
            rc = resolve_pending_remap();
            if (rc) return rc;
            rc = upload_colptrs(which);                               // the gather pass moved the column into the other buffer
            if (rc) return rc;
        }
    }
    for (size_t si = 0; si < sweeps.size(); ++si) {
        PlannedSweep &ps = sweeps[si];
        const PhaseTab *d_ptabs_here = d_ptabs_;
        bool ptabs_resident = false;
        if (!ps.ptabs.empty()) {
            if (g_capture_arena && g_capture_arena->d_blob) {
                if (const void *dp = g_capture_arena->put(ps.ptabs.data(), sizeof(PhaseTab) * ps.ptabs.size())) {
                    d_ptabs_here = static_cast<const PhaseTab *>(dp);
                    ptabs_resident = true;
                }
            }
            if (!ptabs_resident)
                CK(cudaMemcpyAsync(d_ptabs_, ps.ptabs.data(), sizeof(PhaseTab) * ps.ptabs.size(), cudaMemcpyHostToDevice, stream_));
        }
        ps.prog.generate = (generate && si == 0) ? 1 : 0;
        const bool last = si + 1 == sweeps.size();
        // a relabelling store planned for the middle of the batch (Planner mid_relabel): out of place, and the qubit map is
        // composed with it afterwards -- every later sweep of the plan is expressed in the layout it leaves
        const bool mid = !last && !ps.mid_dstpos.empty();
        if (mid) set_relabel(ps.prog, ps.mid_dstpos, false);
        bool relabel = false;
        const bool try_leaf = last && final_relabel && want_leaf_fusion_ && n_ >= 12 && sweep_uses_ladder_kernel(ps.prog);
        // with the layout already canonical the last sweep is still sent through the (out-of-place) staged store when a
        // measurement follows: its store pass then delivers the canonical leaf totals, which saves the read pass
        bool leaf_inplace = false;
        if (last && final_relabel && ident && try_leaf && which.size() == cols_.size() && want_inplace_relabel()) {
            // no room for a second buffer (128 GiB shards): the same fused store pass, launched in place -- a tile is read
            // completely before its CTA stores it, and with the layout unchanged it stores exactly what it has read
            std::vector<int> dstpos(n_);
            for (int l = 0; l < n_; ++l) dstpos[l] = l;
            const SweepProgram before = ps.prog;
            set_relabel(ps.prog, dstpos, true);
            if (ps.prog.leaf_fuse && !ensure_scratch(which.size())) { ps.prog.direct_store = 0; leaf_inplace = true; }
            else ps.prog = before;
        }
        if (last && final_relabel && (!ident || try_leaf) && which.size() == cols_.size() && !want_inplace_relabel()) {
            std::vector<int> dstpos(n_);
            for (int l = 0; l < n_; ++l) dstpos[perm_[l]] = l;
            if (can_fuse_relabel(ps.prog, dstpos)) {
                const SweepProgram before = ps.prog;
                set_relabel(ps.prog, dstpos, try_leaf);
                relabel = true;
                if (ps.prog.leaf_fuse) {
                    if (ensure_scratch(which.size())) ps.prog.leaf_fuse = 0;
                    else ps.prog.direct_store = 0;                    // the totals come out of the staged store pass
                }
                if (ident && !ps.prog.leaf_fuse) { ps.prog = before; relabel = false; }     // nothing gained: run in place
            }
        }
        // a tracked sweep whose steps all act on bits still pinned to 0 is a pure broadcast (kernels.cu
        // ladder_broadcast_tiles), which lives in the staged store pass: give up the direct store for it
        if (!relabel && ps.prog.sup_mode && !ps.prog.generate && ps.prog.direct_store && ps.prog.nrounds > 0) {
            bool ok = true;
            uint32_t tgt = 0;
            for (int r = 0; r < ps.prog.nrounds && ok; ++r) {
                const RoundDesc &R = ps.prog.rounds[r];
                for (int j = kRegBits - R.nsteps; j < kRegBits; ++j) {
                    const unsigned tb = R.reg_tb[j];
                    if (!((R.smask >> j) & 1u) || ((tgt >> tb) & 1u)) ok = false;
                    for (unsigned long long g : gen)
                        if ((g >> ps.prog.tsrc[tb]) & 1ull) ok = false;
                    tgt |= 1u << tb;
                }
            }
            if (ok && __builtin_popcount(tgt) >= 2) ps.prog.direct_store = 0;
        }
        if (si == 0 && gather_dst) {
            Column &col = cols_[which[0]];
            free_bufs_.erase(std::find(free_bufs_.begin(), free_bufs_.end(), gather_dst));
            RemoteGather rg;
            std::memset(&rg, 0, sizeof rg);
            rg.k = grp_.pend.k; rg.rank = grp_.rank; rg.P = grp_.P;
            for (int j = 0; j < rg.k; ++j) { rg.gb[j] = grp_.pend.gb[j]; rg.lp[j] = grp_.pend.lp[j]; }
            rg.peer_buf = grp_.d_peer_buf;
            rg.my_mail = nullptr;      // set after the barrier: the slot of its epoch
            grp_.pending = false;
            rc = group_barrier();                                     // every rank has finished what precedes, and published its buffer
            if (rc) return rc;
            rg.my_mail = grp_.mail + 64 * (grp_.epoch & 1ull);
            double2 *h[2] = { col.buf, gather_dst };
            CK(cudaMemcpyAsync(d_pair_, h, sizeof h, cudaMemcpyHostToDevice, stream_));
            fill_byte_tables(ps.prog);
            time_begin();
            stats.h2d_bytes += sizeof(SweepProgram);
            CK(launch_sweep(ps.prog, d_pair_, d_pair_ + 1, 1, d_ptabs_here, nullptr, stream_, ps.prog.leaf_fuse && relabel ? d_leaf_ : nullptr, nullptr,
                            nullptr, &rg));
            time_end(stats.sweep_ms);
            double2 *old = col.buf;
            col.buf = gather_dst;
            release_column(old);
            rc = group_barrier();                                     // nobody overwrites a shard its peers may still be reading
            if (rc) return rc;
            stats.kernel_launches++;
            stats.sweeps++;
            stats.fused_remaps++;
            stats.sweep_column_passes += 1;
            stats.sweep_bytes += bytes_moved[si];
            stats.peer_swap_bytes += (((1ull << rg.k) - 1ull) << (n_ - rg.k)) * 16ull;     // remote reads of this rank
            if (relabel) {
                stats.fused_relabels++;
                if (ps.prog.leaf_fuse) leaf_fused_ = true;
                for (int l = 0; l < n_; ++l) perm_[l] = l;
            } else {
                if (mid) {
                    stats.fused_relabels++;
                    for (int l = 0; l < n_; ++l) perm_[l] = ps.mid_dstpos[perm_[l]];
                    ident = true;
                    for (int l = 0; l < n_; ++l) ident = ident && perm_[l] == l;
                }
                rc = upload_colptrs(which);                           // the sweeps that follow run in place in the new buffer
                if (rc) return rc;
            }
            continue;
        }
        // dense ladder sweeps take their tiles by TMA (planner.cpp apply_tma_layout, kernels.cu ladder_kernel)
        const bool tma_ok = tma_ && !ps.prog.generate && ps.prog.sup_mode == 0 && sweep_uses_ladder_kernel(ps.prog) && tma_available();
        if (!relabel && !mid) {
            std::vector<const double2 *> hcols;
            if (tma_ok && which.size() <= (size_t)kMaxTmaCols) {
                SweepProgram q = ps.prog;
                for (int c : which) hcols.push_back(cols_[c].buf);
                if (apply_tma_layout(q) && tma_can_encode(q, hcols.data(), (int)hcols.size())) {
                    ps.prog = q;
                    stats.tma_sweeps++;
                } else hcols.clear();
            }
            fill_byte_tables(ps.prog);
            time_begin();
            // (captured into a graph, the program upload re-reads its source at every replay: pinned copy owned by the graph)
            const SweepProgram *prog_src = &ps.prog, *d_prog = nullptr;
            if (g_capture_arena) {
                if (!sweep_uses_ladder_kernel(ps.prog) && g_capture_arena->d_blob)
                    d_prog = static_cast<const SweepProgram *>(g_capture_arena->put(&ps.prog, sizeof(SweepProgram)));     // device-resident
                if (!d_prog) {
                    prog_src = static_cast<const SweepProgram *>(g_capture_arena->stage(&ps.prog, sizeof(SweepProgram)));
                    if (!prog_src) return fail(Q1T_ERR_CUDA, "out of pinned host memory while capturing a sweep batch");
                }
            }
            if (!d_prog) stats.h2d_bytes += sizeof(SweepProgram);
            CK(launch_sweep(*prog_src, d_colptrs_, d_colptrs_, (int)which.size(), d_ptabs_here, d_gen_, stream_, leaf_inplace ? d_leaf_ : nullptr,
                            hcols.empty() ? nullptr : hcols.data(), d_prog));
            if (leaf_inplace) leaf_fused_ = true;
            time_end(stats.sweep_ms);
            stats.kernel_launches++;
            stats.sweeps++;
            stats.sweep_column_passes += which.size();
            stats.sweep_bytes += (uint64_t)which.size() * bytes_moved[si];
            continue;
        }
        // out-of-place: one column at a time through a scratch buffer
        bool tma_here = false;
        if (tma_ok && !which.empty()) {
            SweepProgram q = ps.prog;
            const double2 *h0 = cols_[which[0]].buf;
            if (apply_tma_layout(q) && tma_can_encode(q, &h0, 1)) {
                ps.prog = q;
                tma_here = true;
                stats.tma_sweeps++;
            }
        }
        fill_byte_tables(ps.prog);
        for (size_t i = 0; i < which.size(); ++i) {
            Column &col = cols_[which[i]];
            double2 *scratch = nullptr;
            rc = alloc_column(&scratch);
            if (rc) return rc;
            double2 *h[2] = { col.buf, scratch };
            CK(cudaMemcpyAsync(d_pair_, h, sizeof h, cudaMemcpyHostToDevice, stream_));
            time_begin();
            const double2 *hsrc = col.buf;
            stats.h2d_bytes += sizeof(SweepProgram); CK(launch_sweep(ps.prog, d_pair_, d_pair_ + 1, 1, d_ptabs_, d_gen_ ? d_gen_ + i : nullptr, stream_,
                            ps.prog.leaf_fuse ? d_leaf_ + (i << (n_ - kCanonLeafBits)) : nullptr, tma_here ? &hsrc : nullptr));
            time_end(stats.sweep_ms);
            stats.kernel_launches++;
            double2 *old = col.buf;
            col.buf = scratch;
            release_column(old);
        }
        stats.sweeps++;
        stats.fused_relabels++;
        if (ps.prog.leaf_fuse) leaf_fused_ = true;
        stats.sweep_column_passes += which.size();
        stats.sweep_bytes += (uint64_t)which.size() * bytes_moved[si];
        if (mid) {
            for (int l = 0; l < n_; ++l) perm_[l] = ps.mid_dstpos[perm_[l]];
            ident = true;
            for (int l = 0; l < n_; ++l) ident = ident && perm_[l] == l;
            rc = upload_colptrs(which);                               // the sweeps that follow run in place in the new buffers
            if (rc) return rc;
        } else {
            for (int l = 0; l < n_; ++l) perm_[l] = l;
        }
    }
    sweeps.clear();
    return Q1T_OK;
}

int DeviceVectorState::run_queue(bool final_relabel)
{
    if (queue_.empty()) return resolve_pending_remap();      // (nothing to fuse a recorded remap into)
    int rc = ensure_device();
    if (rc) return rc;
    std::vector<int> which = queue_cols_;
    if (which.empty())
        for (size_t c = 0; c < cols_.size(); ++c) which.push_back((int)c);
    std::vector<LoweredGate> q;
    q.swap(queue_);
    queue_cols_.clear();
    if (grp_.pending) {
        // a recorded qubit remap (option fused_remap): only a batch that is planned as fused sweeps can read through it
        bool fusable = n_ >= 5 && fuse_ && balance_ < 0 && cols_.size() == 1;
        for (const LoweredGate &g : q) fusable = fusable && (g.kind == LoweredGate::POLY || g.kind == LoweredGate::G1);
        if (!fusable) {
            rc = resolve_pending_remap();
            if (rc) return rc;
        }
    }
    auto materialize_all = [&]() -> int {
        for (int c : which) {
            int r = materialize(cols_[c]);
            if (r) return r;
        }
        return upload_colptrs(which);
    };
    if (n_ < 5 || !fuse_) {
        rc = materialize_all();
        if (rc) return rc;
        for (const LoweredGate &g : q) {
            LoweredGate gg;
            to_generic(g, gg);
            rc = run_generic(gg, which);
            if (rc) return rc;
        }
        return Q1T_OK;
    }
    // two packings are tried on the host (planning is cheap): greedy (fill every tile) and balanced
    // (never open a round the tile cannot fill); the one with fewer sweeps, then fewer rounds, runs
    bool balance = balance_ > 0;
    if (balance_ < 0) {
        bool fusable = true;
        for (const LoweredGate &g : q) fusable = fusable && (g.kind == LoweredGate::POLY || g.kind == LoweredGate::G1);
        if (fusable) {
            // plan cache: execute() of the same circuit lowers to the same gate list every time
            // (circuit.rs:594-600 builds a fresh state per call); planning it again is pure host time
            // (two independent 64-bit hashes of the lowered list: a hit needs both, and the gate count, to agree)
            uint64_t key = 1469598103934665603ull, key2 = 0x9E3779B97F4A7C15ull;
            auto mix = [&](const void *p, size_t nbytes) {
                const unsigned char *b = static_cast<const unsigned char *>(p);
                for (size_t i = 0; i < nbytes; ++i) {
                    key ^= b[i]; key *= 1099511628211ull;
                    key2 = (key2 ^ b[i]) * 0xFF51AFD7ED558CCDull; key2 ^= key2 >> 29;
                }
            };
            // batches that start from basis states run with support tracking: their loads are negligible,
            // so 64-byte tiles cost nothing and leave 10 (not 9) free bits per sweep -- provided every sweep
            // of the plan qualifies for the ladder kernel (otherwise the dense default is planned)
            bool sparse_start = sparse_c2_ && track_support_ && coalesce_bits_ == 3 && tile_bits_ == 12;
            for (int c : which) sparse_start = sparse_start && cols_[c].basis && cols_[c].basis_idx != UINT64_MAX;
            // dense batches over the whole state with room for a second buffer may store relabelled in the middle of the
            // plan (Planner mid_relabel: a sweep with strided tiles writes whole contiguous tiles instead)
            bool mid_ok = mid_relabel_ > 0 && (mid_relabel_ > 1 || n_ >= 24) && !sparse_start && which.size() == cols_.size() && !want_inplace_relabel();
            for (int c : which) mid_ok = mid_ok && !cols_[c].basis;
            const long cfg[6] = { (long)n_, tile_bits_, coalesce_bits_, (long)kRegBits, (long)sparse_start, (long)mid_ok };
            mix(cfg, sizeof cfg);
            for (const LoweredGate &g : q) {
                const int hdr[5] = { (int)g.kind, g.nb, g.b[0], g.b[1], g.target };
                mix(hdr, sizeof hdr);
                mix(&g.cmask, sizeof g.cmask);
                if (g.kind == LoweredGate::POLY) { mix(&g.c0, sizeof g.c0); mix(g.lin, sizeof g.lin); mix(&g.quad, sizeof g.quad); }
                else mix(g.m, sizeof g.m);
            }
            {
                std::lock_guard<std::mutex> lk(g_plan_mu);
                for (PlanCacheEntry &pc : g_plan_cache)
                    if (pc.key == key && pc.key2 == key2 && pc.ngates == q.size()) {
                        std::vector<PlannedSweep> plan = pc.sweeps;
                        pc.stamp = ++g_plan_clock;
                        stats.plan_cache_hits++;
                        cur_plan_key_ = key ? key : 1;
                        const int rcs = run_sweeps(plan, which, final_relabel);
                        cur_plan_key_ = 0;
                        return rcs;
                    }
            }
            // Candidates: {64-byte, 128-byte tiles} x {greedy, balanced packing}.  A dense batch is judged by sweeps, then
            // rounds.  A batch that starts from basis states and runs tracked (every sweep in the ladder kernel) moves only
            // the support of the state, so it is judged by the BYTES its sweeps will move: the packing decides how much of
            // the growth of the support falls into the last sweeps (a plan that leaves three low bits for a last small
            // sweep reads and writes the whole state once more).
            std::vector<PlannedSweep> plan[2];
            std::vector<PlannedSweep> best;
            uint64_t best_cost[3] = { ~0ull, ~0ull, ~0ull };
            auto tracked_bytes = [&](const std::vector<PlannedSweep> &pl) {
                uint64_t pinned = n_ >= 64 ? ~0ull : ((1ull << n_) - 1ull), total = 0;
                for (size_t si = 0; si < pl.size(); ++si) {
                    const SweepProgram &P = pl[si].prog;
                    int pinned_outer = 0, pinned_all = 0;
                    for (int i = 0; i < P.n_outer; ++i) pinned_outer += (int)((pinned >> P.osrc[i]) & 1ull);
                    for (int b = 0; b < n_; ++b) pinned_all += (int)((pinned >> b) & 1ull);
                    const bool last = si + 1 == pl.size();
                    const uint64_t tiles_written = last ? (1ull << P.n_outer) : (1ull << (P.n_outer - pinned_outer));
                    // a live element whose lowest index bits are still pinned sits alone in its 32-byte sector / 64-byte burst
                    const uint64_t per_read = (pinned & 1ull) ? ((pinned & 2ull) ? 64 : 32) : 16;
                    total += (tiles_written << P.T) * 16ull + (si == 0 ? 0ull : (1ull << (n_ - pinned_all)) * per_read);
                    pinned &= ~pl[si].touched;
                }
                return total;
            };
            for (int attempt = sparse_start ? 0 : 1; attempt < 2; ++attempt) {
                const int cbits = attempt == 0 ? 2 : (int)coalesce_bits_;
                for (int b = 0; b < 2; ++b) {
                  // (dense batches: also with the relabelling stores of planner.h, plain and with the next targets rotated into
                  //  the low positions; a plan whose last sweep cannot restore the canonical layout pays a relabel sweep)
                  static const bool rotate_ok = !(std::getenv("Q1T_MID_ROTATE") && std::atoi(std::getenv("Q1T_MID_ROTATE")) == 0);      // A/B switch
                  for (int mid = 0; mid <= (mid_ok && attempt == 1 ? (rotate_ok ? 2 : 1) : 0); ++mid) {
                    if (mid_ok && attempt == 1 && mid == 0) continue;
                    Planner trial(n_, (int)tile_bits_, cbits, b != 0, mid);
                    if (mid >= 2) {
                        std::vector<int> la;
                        for (const LoweredGate &g : q)
                            if (g.kind == LoweredGate::G1) la.push_back(g.target);
                        trial.set_lookahead(la);
                    }
                    for (const LoweredGate &g : q) trial.add(g);
                    trial.finish();
                    plan[b] = trial.take();
                    uint64_t extra_sweeps = 0;
                    if (mid > 0 && final_relabel && !plan[b].empty()) {
                        std::vector<int> pm(perm_.begin(), perm_.end()), dstpos(n_);
                        for (size_t si = 0; si + 1 < plan[b].size(); ++si)
                            if (!plan[b][si].mid_dstpos.empty())
                                for (int l = 0; l < n_; ++l) pm[l] = plan[b][si].mid_dstpos[pm[l]];
                        bool idn = true;
                        for (int l = 0; l < n_; ++l) { dstpos[pm[l]] = l; idn = idn && pm[l] == l; }
                        if (!idn && !can_fuse_relabel(plan[b].back().prog, dstpos)) extra_sweeps = 1;
                    }
                    bool all_ladder = true;
                    for (const PlannedSweep &ps : plan[b]) all_ladder = all_ladder && sweep_uses_ladder_kernel(ps.prog);
                    if (attempt == 0 && !all_ladder) continue;        // a 64-byte plan only pays when it runs tracked
                    const bool tracked = sparse_start && all_ladder;
                    const uint64_t c[3] = { tracked ? tracked_bytes(plan[b]) : ~0ull - 1, trial.stats.sweeps + extra_sweeps, trial.stats.rounds };
                    if (c[0] < best_cost[0] || (c[0] == best_cost[0] && (c[1] < best_cost[1] || (c[1] == best_cost[1] && c[2] < best_cost[2])))) {
                        best_cost[0] = c[0]; best_cost[1] = c[1]; best_cost[2] = c[2];
                        best = plan[b];
                    }
                  }
                }
            }
            plan[0].swap(best);
            const int pick = 0;
            if (q.size() >= 16) {
                std::lock_guard<std::mutex> lk(g_plan_mu);
                if (g_plan_cache.size() >= kPlanCacheMax) {       // evict the least recently used plan
                    size_t lru = 0;
                    for (size_t i = 1; i < g_plan_cache.size(); ++i)
                        if (g_plan_cache[i].stamp < g_plan_cache[lru].stamp) lru = i;
                    g_plan_cache.erase(g_plan_cache.begin() + lru);
                }
                g_plan_cache.push_back({ key, key2, q.size(), ++g_plan_clock, plan[pick] });
            }
            cur_plan_key_ = q.size() >= 16 ? (key ? key : 1) : 0;
            const int rcs = run_sweeps(plan[pick], which, final_relabel);
            cur_plan_key_ = 0;
            return rcs;
        }
    }
    Planner pl(n_, (int)tile_bits_, (int)coalesce_bits_, balance);
    for (const LoweredGate &g : q) {
        if (g.kind == LoweredGate::POLY || g.kind == LoweredGate::G1) {
            pl.add(g);
            continue;
        }
        LoweredGate gg;
        to_generic(g, gg);
        uint64_t m = 0;
        for (int p : gg.pos) m |= 1ull << p;
        pl.flush_diag_touching(m);
        pl.cut();
        std::vector<PlannedSweep> sw = pl.take();
        rc = run_sweeps(sw, which, false);
        if (rc) return rc;
        rc = materialize_all();
        if (rc) return rc;
        rc = run_generic(gg, which);
        if (rc) return rc;
    }
    pl.finish();
    std::vector<PlannedSweep> sw = pl.take();
    return run_sweeps(sw, which, final_relabel);
}

// Relabel in place?  Forced by the "inplace_relabel" option, else only when the live columns plus one
// scratch column cannot fit into 90 % of the device memory.
bool DeviceVectorState::want_inplace_relabel()
{
    if (inplace_relabel_ > 0) return n_ >= 5;
    if (inplace_relabel_ < 0) return false;
    size_t live = 0;
    for (const Column &c : cols_)
        if (c.buf) ++live;
    static size_t total_of[64] = { 0 };
    size_t total_b = (device_ >= 0 && device_ < 64) ? total_of[device_] : 0;
    if (!total_b) {
        size_t free_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return false; }
        if (device_ >= 0 && device_ < 64) total_of[device_] = total_b;
    }
    return (live + 1) * (sizeof(double2) << n_) > total_b / 100 * 90;
}

// undo the zero-byte swap relabelling: one out-of-place relabel sweep per column
// (or a few in-place passes when memory is tight)
int DeviceVectorState::canonicalize()
{
    bool ident = true;
    for (int l = 0; l < n_; ++l)
        if (perm_[l] != l) ident = false;
    if (ident) return Q1T_OK;
    std::vector<int> dstpos(n_);
    for (int l = 0; l < n_; ++l) dstpos[perm_[l]] = l;
    if (want_inplace_relabel()) {
        // no room for a second column buffer (128 GiB shards): a few tile-closed passes with
        // source == destination instead of one out-of-place sweep (planner.cpp, plan_inplace_relabel)
        const std::vector<InplacePass> passes = plan_inplace_relabel(n_, (int)tile_bits_, 3, dstpos);
        if (passes.empty()) return fail(Q1T_ERR_UNSUPPORTED, "in-place relabelling needs at least 5 tile bits");
        for (const InplacePass &pass : passes) {
            PlannedSweep ps = build_permute_sweep(n_, (int)tile_bits_, pass.dstpos, &pass.tile);
            for (size_t c = 0; c < cols_.size(); ++c) {
                Column &col = cols_[c];
                if (col.basis) continue;     // lazy basis columns are stored by logical index; nothing to move
                double2 *h[2] = { col.buf, col.buf };
                // d_pair_ is rewritten per launch: stream order keeps the previous launch's copy intact until it has run
                CK(cudaMemcpyAsync(d_pair_, h, sizeof h, cudaMemcpyHostToDevice, stream_));
                time_begin();
                stats.h2d_bytes += sizeof(SweepProgram); CK(launch_sweep(ps.prog, d_pair_, d_pair_ + 1, 1, d_ptabs_, nullptr, stream_));
                time_end(stats.sweep_ms);
                stats.kernel_launches++;
                stats.sweeps++;
                stats.permute_sweeps++;
                stats.sweep_column_passes++;
                stats.sweep_bytes += 32ull << n_;
            }
        }
        for (int l = 0; l < n_; ++l) perm_[l] = l;
        return Q1T_OK;
    }
    PlannedSweep ps = build_permute_sweep(n_, (int)tile_bits_, dstpos);
    for (size_t c = 0; c < cols_.size(); ++c) {
        Column &col = cols_[c];
        if (col.basis) {
            // lazy basis columns are stored by logical index; nothing to move
            continue;
        }
        double2 *scratch = nullptr;
        int rc = alloc_column(&scratch);
        if (rc) return rc;
        double2 *h[2] = { col.buf, scratch };
        CK(cudaMemcpyAsync(d_pair_, h, sizeof h, cudaMemcpyHostToDevice, stream_));
        time_begin();
        stats.h2d_bytes += sizeof(SweepProgram); CK(launch_sweep(ps.prog, d_pair_, d_pair_ + 1, 1, d_ptabs_, nullptr, stream_));
        time_end(stats.sweep_ms);
        stats.kernel_launches++;
        stats.sweeps++;
        stats.permute_sweeps++;
        stats.sweep_column_passes++;
        stats.sweep_bytes += 32ull << n_;
        double2 *old = col.buf;
        col.buf = scratch;
        release_column(old);
    }
    for (int l = 0; l < n_; ++l) perm_[l] = l;
    return Q1T_OK;
}

int DeviceVectorState::flush_async()
{
    int rc = ensure_device();
    if (rc) return rc;
    rc = run_queue(true);
    if (rc) return rc;
    rc = resolve_pending_remap();
    if (rc) return rc;
    rc = canonicalize();
    if (rc) return rc;
    return apply_pending_scale();
}

// a scalar that no sweep has taken along: one scaling pass over the dense columns (lazy basis columns
// cannot carry a factor, they are materialised)
int DeviceVectorState::apply_pending_scale()
{
    if (pending_scale_ == 1.0) return Q1T_OK;
    const double f = pending_scale_;
    pending_scale_ = 1.0;
    for (Column &c : cols_) {
        if (c.basis && c.basis_idx == UINT64_MAX) continue;            // all-zero stays all-zero
        int rc = materialize(c);
        if (rc) return rc;
        CK(launch_scale2(c.buf, c.buf, nullptr, n_, f, 0.0, stream_));
        stats.kernel_launches++;
        stats.sweep_column_passes++;
        stats.sweep_bytes += 32ull << n_;
    }
    stats.sweeps++;
    return Q1T_OK;
}

// every column times re + i im.  The modulus (with the sign of a real factor) is deferred like the Hadamard
// normalisations; a genuine phase is queued as a global-phase gate.
int DeviceVectorState::scale_all(double re, double im)
{
    if (!queue_cols_.empty()) { int rc = run_queue(); if (rc) return rc; }
    if (im == 0.0) { pending_scale_ *= re; return Q1T_OK; }
    const double mod = std::hypot(re, im);
    pending_scale_ *= mod;
    if (mod == 0.0) return Q1T_OK;
    const double m[8] = { re / mod, im / mod, 0, 0, 0, 0, re / mod, im / mod };
    const size_t bit = 0;
    return apply_gate(m, 2, &bit, 1, "phase");
}

int DeviceVectorState::flush()
{
    int rc = flush_async();
    if (rc) return rc;
    CK(cudaStreamSynchronize(stream_));
    return Q1T_OK;
}

// Replay of a gate run that was lowered before FROM THE IDENTITY LAYOUT (Circuit::execute on a fresh state: the
// lowering of a gate depends on the qubit relabelling in force, and the same run from the same start reproduces
// the same sequence of relabellings).  Skips matrix(), classification and control extraction of every gate.
int DeviceVectorState::apply_lowered(const std::vector<LoweredGate> &lgs)
{
    if (!queue_cols_.empty()) { int rc = run_queue(); if (rc) return rc; }
    for (int l = 0; l < n_; ++l)
        if (perm_[l] != l) return fail(Q1T_ERR_INVALID_ARGUMENT, "apply_lowered: the state is not in the identity layout");
    for (const LoweredGate &lg : lgs) {
        int rc = enqueue_lowered(lg);
        if (rc) return rc;
        if (queue_.size() >= 8192) { rc = run_queue(); if (rc) return rc; }
    }
    return Q1T_OK;
}

// vectorstate.rs:166-178
int DeviceVectorState::apply_gate(const double *mat, size_t dim, const size_t *bits, size_t k, const char *desc)
{
    if (!queue_cols_.empty()) { int rc = run_queue(); if (rc) return rc; }
    int rc = lower_and_queue(mat, dim, bits, k, desc);
    if (rc) return rc;
    if (queue_.size() >= 8192) return run_queue();
    return Q1T_OK;
}

// vectorstate.rs:180-189 -- the n one-qubit gates are queued and fused into a
// handful of sweeps instead of n passes
int DeviceVectorState::apply_unary_gate_all(const double *mat, size_t dim, const char *desc)
{
    for (size_t b = 0; b < (size_t)n_; ++b) {
        int rc = apply_gate(mat, dim, &b, 1, desc);
        if (rc) return rc;
    }
    return Q1T_OK;
}

// vectorstate.rs:193-227 + qustate.rs:100-127
int DeviceVectorState::apply_conditional_gate(const uint8_t *control, size_t ncontrol, const double *mat, size_t dim,
                                              const size_t *bits, size_t k, const char *desc)
{
    if (ncontrol != shots_) {
        char buf[320];
        std::snprintf(buf, sizeof buf, "The number of runs is %zu, but received %zu control bits for controlled %s operation",
                      shots_, ncontrol, desc ? desc : "gate");
        return fail(Q1T_ERR_INVALID_NR_CONTROL_BITS, buf);
    }
    size_t gate_bits = 0;
    while (((size_t)1 << gate_bits) < dim) ++gate_bits;
    if (((size_t)1 << gate_bits) != dim || gate_bits != k) {
        char buf[256];
        std::snprintf(buf, sizeof buf, "Expected %zu bits for \"%s\", got %zu", gate_bits, desc ? desc : "gate", k);
        return fail(Q1T_ERR_INVALID_NR_BITS, buf);
    }
    if (!control && ncontrol) return fail(Q1T_ERR_INVALID_ARGUMENT, "NULL control array");
    int rc = ensure_device();
    if (rc) return rc;
    rc = run_queue();
    if (rc) return rc;
    // collect_conditional_ranges (qustate.rs:100-127)
    struct Range { int icol; size_t len; bool apply; };
    std::vector<Range> ranges;
    size_t off = 0;
    for (size_t icol = 0; icol < cols_.size(); ++icol) {
        const size_t count = cols_[icol].count;
        if (count == 0) continue;
        size_t begin = off;
        bool prev = control[off] != 0;
        for (size_t ibit = off + 1; ibit < off + count; ++ibit) {
            if ((control[ibit] != 0) != prev) {
                ranges.push_back({ (int)icol, ibit - begin, prev });
                begin = ibit;
                prev = !prev;
            }
        }
        if (begin < off + count) ranges.push_back({ (int)icol, off + count - begin, prev });
        off += count;
    }
    // new column list: the last range of a column takes over its buffer, earlier ranges get copies
    std::vector<Column> nc(ranges.size());
    std::vector<int> last_of(cols_.size(), -1);
    for (size_t r = 0; r < ranges.size(); ++r) last_of[ranges[r].icol] = (int)r;
    std::vector<int> flagged;
    for (size_t r = 0; r < ranges.size(); ++r) {
        Column &src = cols_[ranges[r].icol];
        Column &dst = nc[r];
        dst.count = ranges[r].len;
        if (last_of[ranges[r].icol] == (int)r) {
            dst.buf = src.buf; dst.basis = src.basis; dst.basis_idx = src.basis_idx;
            src.buf = nullptr;
        } else if (src.basis) {
            dst.basis = true; dst.basis_idx = src.basis_idx;
        } else {
            rc = alloc_column(&dst.buf);
            if (rc) {
                // out of device memory half-way: the state keeps its columns as they were (buffers that already moved
                // into the new list go back to their columns, copies made so far are released)
                for (size_t u = 0; u < r; ++u) {
                    if (last_of[ranges[u].icol] == (int)u) cols_[ranges[u].icol].buf = nc[u].buf;
                    else if (nc[u].buf) release_column(nc[u].buf);
                }
                return rc;
            }
            CK(cudaMemcpyAsync(dst.buf, src.buf, sizeof(double2) << n_, cudaMemcpyDeviceToDevice, stream_));
            stats.sweep_column_passes++;     // a column copy is one read + one write of the column
        }
        if (ranges[r].apply) flagged.push_back((int)r);
    }
    for (Column &c : cols_)
        if (c.buf) release_column(c.buf);     // columns that had no shots
    cols_.swap(nc);
    if (flagged.empty()) return Q1T_OK;
    no_relabel_ = true;                       // a Swap on a subset of the columns cannot be a (global) relabel
    rc = lower_and_queue(mat, dim, bits, k, desc);
    no_relabel_ = false;
    if (rc) return rc;
    if (queue_.empty()) return Q1T_OK;        // identity
    queue_cols_ = flagged;
    return run_queue();
}

// ---------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------
int DeviceVectorState::ensure_scratch(size_t ncols)
{
    const int leaf_bits = n_ < kCanonLeafBits ? n_ : kCanonLeafBits;
    const size_t nleaves = (size_t)1 << (n_ - leaf_bits);
    const size_t nblocks = (nleaves + kCanonBlock - 1) / kCanonBlock;
    if (ncols * nleaves > leaf_cap_) {
        if (d_leaf_) { CK(cudaStreamSynchronize(stream_)); scratch_free(device_, d_leaf_, sizeof(double) * leaf_cap_); }
        leaf_cap_ = ncols * nleaves;
        CK(scratch_alloc(device_, (void **)&d_leaf_, sizeof(double) * leaf_cap_));
    }
    if (ncols * nblocks > block_cap_) {
        if (d_block_) { CK(cudaStreamSynchronize(stream_)); scratch_free(device_, d_block_, sizeof(double) * block_cap_); }
        block_cap_ = ncols * nblocks;
        CK(scratch_alloc(device_, (void **)&d_block_, sizeof(double) * block_cap_));
    }
    if (ncols > totals_cap_) {
        if (d_totals_) { CK(cudaStreamSynchronize(stream_)); scratch_free(device_, d_totals_, sizeof(double) * totals_cap_); }
        totals_cap_ = std::max<size_t>(ncols * 2, 16);
        CK(scratch_alloc(device_, (void **)&d_totals_, sizeof(double) * totals_cap_));
    }
    return Q1T_OK;
}

// canonical totals of |amp|^2 over amplitudes with (index & mask) == want, for
// every device-resident column; leaves d_leaf_/d_block_ holding the prefixes
int DeviceVectorState::reduce_columns(uint64_t mask, uint64_t want, std::vector<double> &totals, std::vector<int> &dev_cols,
                                      bool leaf_totals_ready)
{
    int rc = reduce_launch(mask, want, dev_cols, leaf_totals_ready);
    if (rc) return rc;
    return reduce_fetch(totals, dev_cols.size());
}

// column totals back on the host (blocks until the stream has drained)
int DeviceVectorState::reduce_fetch(std::vector<double> &totals, size_t ndev)
{
    totals.assign(ndev, 0.0);
    if (!ndev) return Q1T_OK;
    CK(cudaMemcpyAsync(totals.data(), d_totals_, sizeof(double) * ndev, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    return Q1T_OK;
}

// canonical leaf totals + prefix scan of every dense column, enqueued only
int DeviceVectorState::reduce_launch(uint64_t mask, uint64_t want, std::vector<int> &dev_cols, bool leaf_totals_ready)
{
    dev_cols.clear();
    for (size_t c = 0; c < cols_.size(); ++c)
        if (!cols_[c].basis) dev_cols.push_back((int)c);
    if (dev_cols.empty()) return Q1T_OK;
    int rc = ensure_scratch(dev_cols.size());
    if (rc) return rc;
    rc = upload_colptrs(dev_cols);
    if (rc) return rc;
    time_begin();
    if (!leaf_totals_ready) {              // else: the last sweep's store pass has already left them in d_leaf_
        CK(launch_leaf_totals(d_colptrs_, (int)dev_cols.size(), d_leaf_, n_, mask, want, stream_));
        stats.kernel_launches += 1;
        stats.read_passes += dev_cols.size();
    }
    CK(launch_scan(d_leaf_, d_block_, d_totals_, (int)dev_cols.size(), n_, stream_));
    time_end(stats.read_ms);
    stats.kernel_launches += 2;
    return Q1T_OK;
}

int DeviceVectorState::marginal0(size_t qbit, double *w0_out)
{
    if (qbit >= (size_t)n_) {
        char buf[128];
        std::snprintf(buf, sizeof buf, "Invalid index %zu for a quantum bit", qbit);
        return fail(Q1T_ERR_INVALID_QBIT, buf);
    }
    int rc = flush();
    if (rc) return rc;
    const uint64_t bit = 1ull << (n_ - 1 - (int)qbit);
    std::vector<double> tot;
    std::vector<int> dev;
    rc = reduce_columns(bit, 0, tot, dev);
    if (rc) return rc;
    size_t k = 0;
    for (size_t c = 0; c < cols_.size(); ++c) {
        if (cols_[c].basis) w0_out[c] = (cols_[c].basis_idx == UINT64_MAX || (cols_[c].basis_idx & bit)) ? 0.0 : 1.0;
        else w0_out[c] = tot[k++];
    }
    return Q1T_OK;
}

int DeviceVectorState::column_totals(double *out)
{
    int rc = flush();
    if (rc) return rc;
    std::vector<double> tot;
    std::vector<int> dev;
    rc = reduce_columns(0, 0, tot, dev);
    if (rc) return rc;
    size_t k = 0;
    for (size_t c = 0; c < cols_.size(); ++c) out[c] = cols_[c].basis ? (cols_[c].basis_idx == UINT64_MAX ? 0.0 : 1.0) : tot[k++];
    return Q1T_OK;
}

// ---------------------------------------------------------------------------
// measurement
// ---------------------------------------------------------------------------
// measure_into (vectorstate.rs:237-329) / peek_into (vectorstate.rs:346-393)
int DeviceVectorState::measure_into(size_t qbit, size_t cbit, uint64_t *res, size_t res_len, q1t_rng rng, bool collapse)
{
    if (qbit >= (size_t)n_) {
        char buf[128];
        std::snprintf(buf, sizeof buf, "Invalid index %zu for a quantum bit", qbit);
        return fail(Q1T_ERR_INVALID_QBIT, buf);
    }
    if (res_len < shots_) {
        char buf[192];
        std::snprintf(buf, sizeof buf, "Not enough space to store %zu measurement results in array of length %zu", shots_, res_len);
        return fail(Q1T_ERR_NOT_ENOUGH_SPACE, buf);
    }
    if (cbit >= 64) return fail(Q1T_ERR_INVALID_ARGUMENT, "classical bit index must be < 64");
    if (!res || !rng.next_u64) return fail(Q1T_ERR_INVALID_ARGUMENT, "NULL pointer argument");
    std::vector<double> w0s(cols_.size());
    int rc = marginal0(qbit, w0s.data());
    if (rc) return rc;
    const int bitpos = n_ - 1 - (int)qbit;
    const uint64_t bit = 1ull << bitpos;
    const uint64_t one_mask = 1ull << cbit, zero_mask = ~one_mask;
    // all Binomial draws first, in column order (vectorstate.rs:263-275)
    std::vector<size_t> n0s(cols_.size());
    for (size_t c = 0; c < cols_.size(); ++c) {
        n0s[c] = (size_t)binomial_sample(rng, cols_[c].count, w0s[c] < 1.0 ? w0s[c] : 1.0);
        if (rng_failed(rng)) return fail(Q1T_ERR_RNG, "the injected random generator ran out of words");
    }
    size_t start = 0;
    for (size_t c = 0; c < cols_.size(); ++c) {
        const size_t n0 = n0s[c], cnt = cols_[c].count;
        for (size_t j = start; j < start + n0; ++j) res[j] &= zero_mask;
        for (size_t j = start + n0; j < start + cnt; ++j) res[j] |= one_mask;
        start += cnt;
    }
    if (!collapse) return Q1T_OK;
    return collapse_columns(qbit, w0s.data(), n0s.data());
}

// collapse / renormalise / branch every column given (w0, n0) per column
// (vectorstate.rs:277-326); column order after a split: [0-branch, 1-branch]
int DeviceVectorState::collapse_columns(size_t qbit, const double *w0s, const size_t *n0s)
{
    if (qbit >= (size_t)n_) return fail(Q1T_ERR_INVALID_QBIT, "Invalid index for a quantum bit");
    int rc = flush();
    if (rc) return rc;
    const int bitpos = n_ - 1 - (int)qbit;
    std::vector<Column> nc;
    for (size_t c = 0; c < cols_.size(); ++c) {
        Column col = cols_[c];
        const size_t n0 = n0s[c], cnt = col.count;
        const double w0 = w0s[c];
        const double f0 = 1.0 / std::sqrt(w0), f1 = 1.0 / std::sqrt(1.0 - w0);
        if (col.basis) {
            // a basis state is an eigenstate: w0 is exactly 0 or 1, nothing changes
            nc.push_back(col);
            continue;
        }
        if (n0 == cnt) {
            CK(launch_collapse(col.buf, col.buf, nullptr, n_, bitpos, f0, 0.0, stream_));
            nc.push_back(col);
        } else if (n0 == 0) {
            CK(launch_collapse(col.buf, nullptr, col.buf, n_, bitpos, 0.0, f1, stream_));
            nc.push_back(col);
        } else {
            Column one;
            rc = alloc_column(&one.buf);
            if (rc) return rc;
            CK(launch_collapse(col.buf, col.buf, one.buf, n_, bitpos, f0, f1, stream_));
            col.count = n0;
            one.count = cnt - n0;
            nc.push_back(col);
            nc.push_back(one);
        }
        stats.kernel_launches++;
        stats.sweep_column_passes++;
    }
    stats.sweeps++;
    cols_.swap(nc);
    CK(cudaStreamSynchronize(stream_));
    return Q1T_OK;
}

// ---------------------------------------------------------------------------
// shard primitives: building blocks of the multi-GPU composition
// (q1tsim_b200/sharded.py).  The host chains the per-rank canonical leaf totals
// in rank order, so the distributed reduction is the same fixed geometry as the
// single-GPU one.
// ---------------------------------------------------------------------------
int DeviceVectorState::init_empty()
{
    { const int rcp = resolve_pending_remap(); if (rcp) return rcp; }      // a recorded qubit remap: the peers still read this shard through it
    int rc = ensure_device();
    if (rc) return rc;
    Column c;                   // lazy all-zero column: generated by the first sweep, never memset + read
    c.count = shots_;
    c.basis = true;
    c.basis_idx = UINT64_MAX;
    cols_.push_back(c);
    return Q1T_OK;
}

size_t DeviceVectorState::nr_leaves() const
{
    const int leaf_bits = n_ < kCanonLeafBits ? n_ : kCanonLeafBits;
    return (size_t)1 << (n_ - leaf_bits);
}

// canonical leaf totals of every column (all amplitudes, or those with qubit `qbit` == 0)
int DeviceVectorState::leaf_totals(size_t qbit, double *out)
{
    int rc = flush();
    if (rc) return rc;
    for (Column &c : cols_) {
        rc = materialize(c);
        if (rc) return rc;
    }
    std::vector<int> all;
    for (size_t c = 0; c < cols_.size(); ++c) all.push_back((int)c);
    rc = ensure_scratch(all.size());
    if (rc) return rc;
    rc = upload_colptrs(all);
    if (rc) return rc;
    const uint64_t mask = qbit < (size_t)n_ ? 1ull << (n_ - 1 - (int)qbit) : 0ull;
    time_begin();
    CK(launch_leaf_totals(d_colptrs_, (int)all.size(), d_leaf_, n_, mask, 0, stream_));
    time_end(stats.read_ms);
    stats.kernel_launches++;
    stats.read_passes += all.size();
    CK(cudaMemcpyAsync(out, d_leaf_, sizeof(double) * all.size() * nr_leaves(), cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    return Q1T_OK;
}

// canonical block totals of every column (shards of >= 2^20 amplitudes): leaf totals and the in-block inclusive
// prefixes stay on the device for resolve_draws_blocks(); only ncols * nblocks doubles cross to the host, which
// continues the chain over blocks in rank order
int DeviceVectorState::block_totals(size_t qbit, double *out)
{
    int rc = block_totals_launch(qbit);
    if (rc) return rc;
    return block_totals_fetch(out);
}

// the two halves of block_totals(): everything enqueued / the copy to the host and the wait.  A host layer that drives
// several shards from one thread launches on all of them before it waits for any.
int DeviceVectorState::block_totals_launch(size_t qbit)
{
    if (nr_leaves() < kCanonBlock) return fail(Q1T_ERR_UNSUPPORTED, "block_totals: shards below 2^20 amplitudes chain their leaves on the host");
    // as in measure_all_into: the last queued sweep may produce the leaf totals in its store pass
    want_leaf_fusion_ = fuse_leaf_totals_ && qbit >= (size_t)n_ && cols_.size() == 1;
    leaf_fused_ = false;
    int rc = flush_async();
    const bool leaf_ready = leaf_fused_;
    want_leaf_fusion_ = leaf_fused_ = false;
    if (rc) return rc;
    for (Column &c : cols_) {
        rc = materialize(c);
        if (rc) return rc;
    }
    std::vector<int> all;
    for (size_t c = 0; c < cols_.size(); ++c) all.push_back((int)c);
    rc = ensure_scratch(all.size());
    if (rc) return rc;
    rc = upload_colptrs(all);
    if (rc) return rc;
    const uint64_t mask = qbit < (size_t)n_ ? 1ull << (n_ - 1 - (int)qbit) : 0ull;
    time_begin();
    if (!leaf_ready) {
        CK(launch_leaf_totals(d_colptrs_, (int)all.size(), d_leaf_, n_, mask, 0, stream_));
        stats.kernel_launches++;
        stats.read_passes += all.size();
    }
    CK(launch_block_scan(d_leaf_, d_block_, (int)all.size(), n_, stream_));
    time_end(stats.read_ms);
    stats.kernel_launches++;
    return Q1T_OK;
}

int DeviceVectorState::block_totals_fetch(double *out)
{
    const size_t nb = nr_leaves() / kCanonBlock;
    CK(cudaSetDevice(device_));
    CK(cudaMemcpyAsync(out, d_block_, sizeof(double) * cols_.size() * nb, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    return Q1T_OK;
}

// resolve sorted draws of column `col` after block_totals(): bp[0] = weight in front of this shard, bp[b + 1] = global
// inclusive prefix through this shard's block b
int DeviceVectorState::resolve_draws_blocks(size_t col, const double *bp, const double *chosen, size_t nd, uint64_t *idx)
{
    if (col >= cols_.size() || !bp) return fail(Q1T_ERR_INVALID_ARGUMENT, "column out of range");
    if (nd == 0) return Q1T_OK;
    CK(cudaSetDevice(device_));
    const size_t nl = nr_leaves(), nb = nl / kCanonBlock;
    if (nd > draws_cap_) {
        CK(cudaStreamSynchronize(stream_));
        scratch_free(device_, d_chosen_, sizeof(double) * draws_cap_);
        scratch_free(device_, d_idx_, sizeof(uint64_t) * draws_cap_);
        draws_cap_ = nd;
        CK(scratch_alloc(device_, (void **)&d_chosen_, sizeof(double) * draws_cap_));
        CK(scratch_alloc(device_, (void **)&d_idx_, sizeof(uint64_t) * draws_cap_));
    }
    CK(cudaMemcpyAsync(d_block_ + col * nb, bp + 1, sizeof(double) * nb, cudaMemcpyHostToDevice, stream_));
    CK(cudaMemcpyAsync(d_chosen_, chosen, sizeof(double) * nd, cudaMemcpyHostToDevice, stream_));
    CK(launch_resolve_draws(cols_[col].buf, d_leaf_ + col * nl, d_block_ + col * nb, n_, d_chosen_, nd,
                            reinterpret_cast<unsigned long long *>(d_idx_), bp[0], stream_, bp[0]));
    stats.kernel_launches++;
    CK(cudaMemcpyAsync(idx, d_idx_, sizeof(uint64_t) * nd, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    return Q1T_OK;
}

// resolve sorted draws against caller-supplied inclusive leaf prefixes P (global chain)
int DeviceVectorState::resolve_draws(size_t col, const double *P, double base, const double *chosen, size_t nd, uint64_t *idx)
{
    if (col >= cols_.size()) return fail(Q1T_ERR_INVALID_ARGUMENT, "column out of range");
    int rc = flush();
    if (rc) return rc;
    rc = materialize(cols_[col]);
    if (rc) return rc;
    if (nd == 0) return Q1T_OK;
    rc = ensure_scratch(1);
    if (rc) return rc;
    const size_t nl = nr_leaves(), nb = (nl + kCanonBlock - 1) / kCanonBlock;
    if (nd > draws_cap_) {
        CK(cudaStreamSynchronize(stream_));
        scratch_free(device_, d_chosen_, sizeof(double) * draws_cap_);
        scratch_free(device_, d_idx_, sizeof(uint64_t) * draws_cap_);
        draws_cap_ = nd;
        CK(scratch_alloc(device_, (void **)&d_chosen_, sizeof(double) * draws_cap_));
        CK(scratch_alloc(device_, (void **)&d_idx_, sizeof(uint64_t) * draws_cap_));
    }
    CK(cudaMemcpyAsync(d_leaf_, P, sizeof(double) * nl, cudaMemcpyHostToDevice, stream_));
    CK(cudaMemsetAsync(d_block_, 0, sizeof(double) * nb, stream_));      // P is already the full prefix
    CK(cudaMemcpyAsync(d_chosen_, chosen, sizeof(double) * nd, cudaMemcpyHostToDevice, stream_));
    CK(launch_resolve_draws(cols_[col].buf, d_leaf_, d_block_, n_, d_chosen_, nd, reinterpret_cast<unsigned long long *>(d_idx_), base, stream_));
    stats.kernel_launches++;
    CK(cudaMemcpyAsync(idx, d_idx_, sizeof(uint64_t) * nd, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    return Q1T_OK;
}

// collapse on a qubit that is a rank bit: this rank keeps (scaled by f0 / f1) or zeroes whole
// columns; column structure follows n0 exactly as collapse_columns does
int DeviceVectorState::scale_split_columns(const double *f0, const double *f1, const size_t *n0s)
{
    int rc = flush();
    if (rc) return rc;
    std::vector<Column> nc;
    for (size_t c = 0; c < cols_.size(); ++c) {
        rc = materialize(cols_[c]);
        if (rc) return rc;
        Column col = cols_[c];
        const size_t n0 = n0s[c], cnt = col.count;
        if (n0 == cnt) {
            CK(launch_scale2(col.buf, col.buf, nullptr, n_, f0[c], 0.0, stream_));
            nc.push_back(col);
        } else if (n0 == 0) {
            CK(launch_scale2(col.buf, col.buf, nullptr, n_, f1[c], 0.0, stream_));
            nc.push_back(col);
        } else {
            Column one;
            rc = alloc_column(&one.buf);
            if (rc) return rc;
            CK(launch_scale2(col.buf, col.buf, one.buf, n_, f0[c], f1[c], stream_));
            col.count = n0;
            one.count = cnt - n0;
            nc.push_back(col);
            nc.push_back(one);
        }
        stats.kernel_launches++;
        stats.sweep_column_passes++;
    }
    stats.sweeps++;
    cols_.swap(nc);
    CK(cudaStreamSynchronize(stream_));
    return Q1T_OK;
}

// replace the column list by basis states (idx) or all-zero columns (idx == UINT64_MAX)
int DeviceVectorState::replace_columns(size_t ncols, const uint64_t *idx, const size_t *counts)
{
    { const int rcp = resolve_pending_remap(); if (rcp) return rcp; }      // a recorded qubit remap: the peers still read this shard through it
    int rc = ensure_device();
    if (rc) return rc;
    queue_.clear();
    queue_cols_.clear();
    pending_scale_ = 1.0;
    CK(cudaStreamSynchronize(stream_));
    for (Column &c : cols_)
        if (c.buf) release_column(c.buf);
    cols_.clear();
    for (int l = 0; l < n_; ++l) perm_[l] = l;
    for (size_t k = 0; k < ncols; ++k) {
        Column c;
        c.count = counts[k];
        if (idx[k] != UINT64_MAX && (idx[k] >> n_)) return fail(Q1T_ERR_INVALID_ARGUMENT, "basis index out of range");
        c.basis = true;                 // lazy: |idx>, or the all-zero column for UINT64_MAX
        c.basis_idx = idx[k];
        cols_.push_back(c);
    }
    return Q1T_OK;
}

// peer exchange over CUDA IPC: export a column, swap in place with a partner's mapped column
int DeviceVectorState::ipc_export(size_t col, unsigned char *handle64)
{
    { const int rcp = resolve_pending_remap(); if (rcp) return rcp; }      // a recorded qubit remap: the peers still read this shard through it
    void *p = nullptr;
    int rc = column_ptr(col, &p);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, p));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::memcpy(handle64, &h, 64);
    return Q1T_OK;
}

int DeviceVectorState::peer_swap(size_t col, const unsigned char *peer_handle64, size_t local_qubit, int my_bit)
{
    { const int rcp = resolve_pending_remap(); if (rcp) return rcp; }      // a recorded qubit remap: the peers still read this shard through it
    if (col >= cols_.size() || local_qubit >= (size_t)n_ || n_ < 2) return fail(Q1T_ERR_INVALID_ARGUMENT, "peer_swap: bad argument");
    void *mine = nullptr;
    int rc = column_ptr(col, &mine);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, peer_handle64, 64);
    // mapping a 16 GiB peer allocation costs milliseconds: keep mappings open for the life of the process
    // (the partner's buffers come from its buffer cache, so the same few handles keep coming back)
    static std::mutex ipc_mu;
    static std::vector<std::pair<std::string, void *>> ipc_open;
    // Large shards (> 31 qubits) are not pooled: the partner frees them eagerly, and memory that is
    // still mapped here would stay allocated there -- map and unmap around the kernel instead.
    const bool keep_mapping = n_ <= 31;   // big shards may be freed by their owner right after: never leave them mapped
    void *theirs = nullptr;
    if (keep_mapping) {
        std::lock_guard<std::mutex> lk(ipc_mu);
        const std::string key(reinterpret_cast<const char *>(peer_handle64), 64);
        for (auto &kv : ipc_open)
            if (kv.first == key) theirs = kv.second;
        if (!theirs) {
            CK(cudaIpcOpenMemHandle(&theirs, h, cudaIpcMemLazyEnablePeerAccess));
            ipc_open.push_back(std::make_pair(key, theirs));
        }
    } else {
        CK(cudaIpcOpenMemHandle(&theirs, h, cudaIpcMemLazyEnablePeerAccess));
    }
    const int L = n_ - 1 - (int)local_qubit;
    cudaEventRecord(ev0_, stream_);
    cudaError_t e = launch_peer_swap(static_cast<double2 *>(mine), static_cast<double2 *>(theirs), n_, L, my_bit ? 1 : 0, stream_);
    cudaEventRecord(ev1_, stream_);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    float ms = 0;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, ev0_, ev1_);
    if (!keep_mapping) cudaIpcCloseMemHandle(theirs);
    if (e != cudaSuccess) return cuda_fail(e, "peer_swap");
    stats.peer_swap_ms += ms;
    stats.peer_swap_bytes += (16ull << n_) / 2;      // a quarter shard read remotely + a quarter written remotely
    stats.kernel_launches++;
    return Q1T_OK;
}

// ---------------------------------------------------------------------------
// peer group.  group_export(): this rank registers the (up to) two shard buffers its single column
// alternates between (column + relabel scratch) and a mailbox, and hands out their IPC handles and
// device pointers; group_open(): the peers' buffers and mailboxes are mapped ONCE (IPC handles between
// processes, plain pointers between states of one process).  After that a qubit remap is three
// stream-ordered launches -- barrier, swap, barrier -- with no host synchronisation, no handle exchange
// and no allocation.
// ---------------------------------------------------------------------------
int DeviceVectorState::group_export(unsigned char *handles, void **ptrs)
{
    int rc = ensure_device();
    if (rc) return rc;
    if (cols_.size() != 1) return fail(Q1T_ERR_UNSUPPORTED, "group_export: the state must have exactly one column");
    rc = flush();
    if (rc) return rc;
    if (!grp_.exported) {
        // the column's buffer (or, for a lazy column, the buffer its first sweep will take) and one scratch
        double2 *b0 = cols_[0].buf;
        if (!b0) {
            if (free_bufs_.empty()) {
                rc = alloc_column(&b0);
                if (rc) return rc;
                free_bufs_.push_back(b0);
            } else b0 = free_bufs_.back();
        }
        double2 *b1 = nullptr;
        for (double2 *p : free_bufs_)
            if (p != b0) b1 = p;
        // a second buffer only if two shards fit (128 GiB shards relabel in place and swap in place: one buffer)
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); total_b = 0; }
        static const bool force_single = std::getenv("Q1T_GROUP_SINGLE_BUFFER") && std::atoi(std::getenv("Q1T_GROUP_SINGLE_BUFFER")) != 0;
        const bool two_fit = !force_single && inplace_relabel_ <= 0 && (total_b == 0 || 2 * (sizeof(double2) << n_) <= total_b / 100 * 90);
        if (!b1 && two_fit) {
            std::vector<double2 *> keep;
            keep.swap(free_bufs_);                       // alloc_column() must not hand b0 out again
            rc = alloc_column(&b1);
            free_bufs_.swap(keep);
            if (rc) return rc;
            free_bufs_.insert(free_bufs_.begin(), b1);
        }
        grp_.bufs[0] = b0;
        grp_.bufs[1] = b1;
        // only registered buffers stay in the free list: whatever the column moves into later must be mapped by the peers
        {
            std::vector<double2 *> keep;
            const size_t bytes = sizeof(double2) << n_;
            for (double2 *p : free_bufs_) {
                if (p == b0 || p == b1) keep.push_back(p);
                else if (!pool_give(device_, bytes, p)) cudaFree(p);
            }
            free_bufs_.swap(keep);
        }
        CK(cudaMalloc(&grp_.mail, 4096));
        CK(cudaMemsetAsync(grp_.mail, 0, 4096, stream_));
        CK(cudaStreamSynchronize(stream_));
        CK(cudaEventCreate(&grp_.ev0));
        CK(cudaEventCreate(&grp_.ev1));
        grp_.exported = true;
    }
    void *all[3] = { grp_.bufs[0], grp_.bufs[1], grp_.mail };
    for (int i = 0; i < 3; ++i) {
        if (ptrs) ptrs[i] = all[i];
        std::memset(handles + 64 * i, 0, 64);
        if (!all[i]) continue;
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, all[i]));
        std::memcpy(handles + 64 * i, &h, 64);
    }
    return Q1T_OK;
}

int DeviceVectorState::group_open(size_t P, size_t rank, const unsigned char *all_handles, void *const *all_ptrs)
{
    if (!grp_.exported) return fail(Q1T_ERR_INVALID_ARGUMENT, "group_open: call group_export first");
    if (grp_.open) return fail(Q1T_ERR_INVALID_ARGUMENT, "group_open: already open");
    if (P < 2 || P > 32 || rank >= P || (P & (P - 1)) || (!all_handles && !all_ptrs))
        return fail(Q1T_ERR_INVALID_ARGUMENT, "group_open: bad argument");
    CK(cudaSetDevice(device_));
    std::vector<void *> pb(2 * P, nullptr);
    std::vector<unsigned long long *> pm(P, nullptr);
    for (size_t r = 0; r < P; ++r) {
        if (r == rank) {
            pb[2 * r] = grp_.bufs[0]; pb[2 * r + 1] = grp_.bufs[1]; pm[r] = grp_.mail;
            continue;
        }
        for (int i = 0; i < 3; ++i) {
            void *p = nullptr;
            if (all_ptrs) p = all_ptrs[3 * r + i];                     // same process: the pointer itself
            else {
                const unsigned char *hb = all_handles + 64 * (3 * r + i);
                bool zero = true;
                for (int b = 0; b < 64; ++b) zero = zero && hb[b] == 0;
                if (!zero) {
                    cudaIpcMemHandle_t h;
                    std::memcpy(&h, hb, 64);
                    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
                    grp_.ipc_opened.push_back(p);
                }
            }
            if (i < 2) pb[2 * r + i] = p;
            else pm[r] = static_cast<unsigned long long *>(p);
        }
        if (!pb[2 * r] || !pm[r]) return fail(Q1T_ERR_INVALID_ARGUMENT, "group_open: a peer exported no buffer");
    }
    CK(cudaMalloc(&grp_.d_peer_buf, sizeof(void *) * 2 * P));
    CK(cudaMalloc(&grp_.d_peer_mail, sizeof(void *) * P));
    CK(cudaMemcpy(grp_.d_peer_buf, pb.data(), sizeof(void *) * 2 * P, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(grp_.d_peer_mail, pm.data(), sizeof(void *) * P, cudaMemcpyHostToDevice));
    grp_.P = (int)P;
    grp_.rank = (int)rank;
    grp_.open = true;
    return Q1T_OK;
}

void DeviceVectorState::group_collect_timing()
{
    if (!grp_.timing_pending) return;
    float ms = 0;
    if (cudaEventSynchronize(grp_.ev1) == cudaSuccess && cudaEventElapsedTime(&ms, grp_.ev0, grp_.ev1) == cudaSuccess) stats.peer_swap_ms += ms;
    else cudaGetLastError();
    grp_.timing_pending = false;
}

int DeviceVectorState::group_barrier()
{
    if (!grp_.open) return fail(Q1T_ERR_INVALID_ARGUMENT, "group_barrier: no open peer group");
    CK(cudaSetDevice(device_));
    unsigned long long cur = 0;
    if (cols_.size() == 1 && cols_[0].buf && cols_[0].buf == grp_.bufs[1]) cur = 1;
    CK(launch_group_barrier(grp_.d_peer_mail, grp_.mail, grp_.P, grp_.rank, ++grp_.epoch, cur, stream_));
    stats.kernel_launches++;
    return Q1T_OK;
}

// rank bits rank_bits[j] trade places with local qubits local_qubits[j] (engine qubit numbering: 0 = top index bit)
int DeviceVectorState::group_remap(size_t k, const int *rank_bits, const size_t *local_qubits)
{
    if (!grp_.open) return fail(Q1T_ERR_INVALID_ARGUMENT, "group_remap: no open peer group");
    if (k < 1 || k > (size_t)kMaxRemapBits || (int)k + 1 > n_ || !rank_bits || !local_qubits)
        return fail(Q1T_ERR_INVALID_ARGUMENT, "group_remap: bad argument");
    int rc = group_remap_prepare();
    if (rc) return rc;
    return group_remap_issue(k, rank_bits, local_qubits);
}

// The two halves of group_remap().  prepare: everything that may allocate, free or wait (queued sweeps, relabels, the
// materialisation of a lazy column).  issue: barrier + swap + barrier, launches only.  A host thread that drives several
// shards ON ONE DEVICE must prepare all of them before it issues any: a cudaFree between two shards' barrier launches
// would wait for the first shard's barrier kernel, which waits for the second's.
int DeviceVectorState::group_remap_prepare()
{
    if (!grp_.open) return fail(Q1T_ERR_INVALID_ARGUMENT, "group_remap: no open peer group");
    if (cols_.size() != 1) return fail(Q1T_ERR_UNSUPPORTED, "group_remap: the state must have exactly one column");
    int rc = flush_async();
    if (rc) return rc;
    rc = materialize(cols_[0]);
    if (rc) return rc;
    if (cols_[0].buf != grp_.bufs[0] && cols_[0].buf != grp_.bufs[1])
        return fail(Q1T_ERR_UNSUPPORTED, "group_remap: the column does not live in a registered buffer");
    group_collect_timing();
    return Q1T_OK;
}

int DeviceVectorState::group_remap_issue(size_t k, const int *rank_bits, const size_t *local_qubits)
{
    if (!grp_.open) return fail(Q1T_ERR_INVALID_ARGUMENT, "group_remap: no open peer group");
    if (k < 1 || k > (size_t)kMaxRemapBits || (int)k + 1 > n_ || !rank_bits || !local_qubits)
        return fail(Q1T_ERR_INVALID_ARGUMENT, "group_remap: bad argument");
    if (cols_.size() != 1 || !cols_[0].buf) return fail(Q1T_ERR_UNSUPPORTED, "group_remap: call group_remap_prepare first");
    int rc = Q1T_OK;
    CK(cudaSetDevice(device_));
    GroupRemapArgs a;
    std::memset(&a, 0, sizeof a);
    a.n = n_; a.k = (int)k; a.P = grp_.P; a.rank = grp_.rank;
    uint64_t used = 0;
    for (size_t j = 0; j < k; ++j) {
        if (rank_bits[j] < 0 || (1 << rank_bits[j]) >= grp_.P || local_qubits[j] >= (size_t)n_)
            return fail(Q1T_ERR_INVALID_ARGUMENT, "group_remap: bit out of range");
        a.gb[j] = rank_bits[j];
        a.lp[j] = n_ - 1 - (int)local_qubits[j];
        if ((used >> a.lp[j]) & 1ull) return fail(Q1T_ERR_INVALID_ARGUMENT, "group_remap: duplicate local qubit");
        used |= 1ull << a.lp[j];
    }
    a.split = n_ - 1;
    while ((used >> a.split) & 1ull) --a.split;                       // the highest index bit that is not traded
    std::vector<int> ins(a.lp, a.lp + k);
    ins.push_back(a.split);
    std::sort(ins.begin(), ins.end());
    for (size_t i = 0; i <= k; ++i) a.ins[i] = ins[i];
    static const int interleave = std::getenv("Q1T_SWAP_INTERLEAVE") ? std::atoi(std::getenv("Q1T_SWAP_INTERLEAVE")) : 1;
    a.interleave = interleave;
    double2 *other = cols_[0].buf == grp_.bufs[0] ? grp_.bufs[1] : cols_[0].buf == grp_.bufs[1] ? grp_.bufs[0] : nullptr;
    if (fused_remap_ && other && std::find(free_bufs_.begin(), free_bufs_.end(), other) != free_bufs_.end()) {
        // recorded only: the next dense ladder sweep gathers its tiles through this map (issue_sweeps); anything else
        // that needs the data first runs it as the swap pass below (resolve_pending_remap)
        grp_.pend = a;
        grp_.pending = true;
        return Q1T_OK;
    }
    return group_remap_run(a);
}

// A recorded trade that no sweep could read through.  It still runs as a gather -- every rank reads its peers' current
// buffers and writes its own other registered buffer -- never as the in-place swap: ranks decide independently whether
// their next batch can fuse (their gate lists differ by the rank-selected blocks), and a peer that swaps in place would
// pull the data away under a peer that gathers.
int DeviceVectorState::resolve_pending_remap()
{
    if (!grp_.pending) return Q1T_OK;
    grp_.pending = false;
    if (cols_.size() != 1 || !cols_[0].buf) return fail(Q1T_ERR_UNSUPPORTED, "recorded qubit remap: the state no longer has one dense column");
    Column &col = cols_[0];
    double2 *dst = col.buf == grp_.bufs[0] ? grp_.bufs[1] : col.buf == grp_.bufs[1] ? grp_.bufs[0] : nullptr;
    std::vector<double2 *>::iterator it = std::find(free_bufs_.begin(), free_bufs_.end(), dst);
    if (!dst || it == free_bufs_.end()) return fail(Q1T_ERR_UNSUPPORTED, "recorded qubit remap: the other registered buffer is not free");
    free_bufs_.erase(it);
    CK(cudaSetDevice(device_));
    RemoteGather rg;
    std::memset(&rg, 0, sizeof rg);
    rg.k = grp_.pend.k; rg.rank = grp_.rank; rg.P = grp_.P;
    for (int j = 0; j < rg.k; ++j) { rg.gb[j] = grp_.pend.gb[j]; rg.lp[j] = grp_.pend.lp[j]; }
    rg.peer_buf = grp_.d_peer_buf;
    rg.my_mail = nullptr;      // set after the barrier: the slot of its epoch
    int rc = group_barrier();
    if (rc) return rc;
    rg.my_mail = grp_.mail + 64 * (grp_.epoch & 1ull);
    CK(cudaEventRecord(grp_.ev0, stream_));
    CK(launch_group_gather(col.buf, dst, rg, n_, stream_));
    CK(cudaEventRecord(grp_.ev1, stream_));
    grp_.timing_pending = true;
    stats.kernel_launches++;
    stats.peer_swap_bytes += (((1ull << rg.k) - 1ull) << (n_ - rg.k)) * 16ull;
    double2 *old = col.buf;
    col.buf = dst;
    release_column(old);
    return group_barrier();
}

int DeviceVectorState::group_remap_run(const GroupRemapArgs &a)
{
    const size_t k = (size_t)a.k;
    int rc = Q1T_OK;
    CK(cudaSetDevice(device_));
    rc = group_barrier();                                             // every rank has finished what precedes, and published its buffer
    if (rc) return rc;
    CK(cudaEventRecord(grp_.ev0, stream_));
    CK(launch_group_swap(cols_[0].buf, grp_.d_peer_buf, grp_.mail + 64 * (grp_.epoch & 1ull), a, stream_));
    CK(cudaEventRecord(grp_.ev1, stream_));
    grp_.timing_pending = true;
    stats.kernel_launches++;
    stats.peer_swap_bytes += (((1ull << k) - 1ull) << (n_ - (int)k - 1)) * 32ull;     // remote reads + remote writes of this rank
    return group_barrier();                                           // nobody touches a shard before every swap has landed
}

int DeviceVectorState::group_close()
{
    if (!grp_.exported) return Q1T_OK;
    resolve_pending_remap();
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    group_collect_timing();
    for (void *p : grp_.ipc_opened) cudaIpcCloseMemHandle(p);
    grp_.ipc_opened.clear();
    if (grp_.d_peer_buf) cudaFree(grp_.d_peer_buf);
    if (grp_.d_peer_mail) cudaFree(grp_.d_peer_mail);
    grp_.d_peer_buf = nullptr;
    grp_.d_peer_mail = nullptr;
    grp_.open = false;
    return Q1T_OK;
}

int DeviceVectorState::column_ptr(size_t col, void **ptr)
{
    if (col >= cols_.size() || !ptr) return fail(Q1T_ERR_INVALID_ARGUMENT, "column out of range");
    int rc = flush();
    if (rc) return rc;
    rc = materialize(cols_[col]);
    if (rc) return rc;
    CK(cudaStreamSynchronize(stream_));
    *ptr = cols_[col].buf;
    return Q1T_OK;
}

// measure_all_into / peek_all_into (vectorstate.rs:106-161)
int DeviceVectorState::measure_all_into(const size_t *cbits, size_t ncbits, uint64_t *res, size_t res_len, q1t_rng rng,
                                        bool collapse)
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
    want_leaf_fusion_ = fuse_leaf_totals_;
    leaf_fused_ = false;
    int rc = flush_async();                 // sweeps (and relabel) enqueued, not waited for
    const bool leaf_ready = leaf_fused_;
    want_leaf_fusion_ = leaf_fused_ = false;
    if (rc) return rc;
    std::vector<double> totals;
    std::vector<int> dev;
    rc = reduce_launch(0, 0, dev, leaf_ready);
    if (rc) return rc;
    // While the device works: draw the uniforms of every shot (one word per shot, column order, as
    // vectorstate.rs:120-133 consumes them) and sort them per column.  chosen = u * scale + low is monotone in
    // u, so scaling the sorted u's later gives exactly the sorted `chosen` values.
    std::vector<double> unit;               // value0_1 of every shot of the dense columns, grouped per column, sorted
    {
        size_t ndense = 0;
        for (const Column &c : cols_)
            if (!c.basis) ndense += c.count;
        unit.reserve(ndense);
        for (const Column &c : cols_) {
            if (c.basis) {
                // WeightedIndex over a unit vector: every draw consumes one word and returns basis_idx
                for (size_t j = 0; j < c.count; ++j) (void)rng.next_u64(rng.ctx);
                continue;
            }
            const size_t at = unit.size();
            for (size_t j = 0; j < c.count; ++j) unit.push_back(uniform_unit(rng));
            std::sort(unit.begin() + (long)at, unit.end());
        }
    }
    rc = reduce_fetch(totals, dev.size());
    if (rc) return rc;
    if (rng_failed(rng)) return fail(Q1T_ERR_RNG, "the injected random generator ran out of words");
    const int leaf_bits = n_ < kCanonLeafBits ? n_ : kCanonLeafBits;
    const size_t nleaves = (size_t)1 << (n_ - leaf_bits);
    const size_t nblocks = (nleaves + kCanonBlock - 1) / kCanonBlock;

    std::vector<std::pair<uint64_t, size_t>> state_counts;   // (basis index, multiplicity), grouped per column
    std::vector<double> chosen;
    std::vector<uint64_t> idx;
    size_t k = 0, unit_at = 0;
    for (size_t c = 0; c < cols_.size(); ++c) {
        const size_t cnt = cols_[c].count;
        if (cols_[c].basis && cols_[c].basis_idx == UINT64_MAX) return fail(Q1T_ERR_INVALID_ARGUMENT, "state column has zero norm");
        if (cols_[c].basis) {
            if (cnt) state_counts.push_back(std::make_pair(cols_[c].basis_idx, cnt));
            continue;
        }
        const double total = totals[k];
        if (!(total > 0.0)) return fail(Q1T_ERR_INVALID_ARGUMENT, "state column has zero norm");
        const UniformF64 u = uniform_new(0.0, total);
        chosen.resize(cnt);
        for (size_t j = 0; j < cnt; ++j) chosen[j] = uniform_scale(u, unit[unit_at + j]);
        unit_at += cnt;
        if (cnt > draws_cap_) {
            CK(cudaStreamSynchronize(stream_));
            scratch_free(device_, d_chosen_, sizeof(double) * draws_cap_);
            scratch_free(device_, d_idx_, sizeof(uint64_t) * draws_cap_);
            draws_cap_ = cnt;
            CK(scratch_alloc(device_, (void **)&d_chosen_, sizeof(double) * draws_cap_));
            CK(scratch_alloc(device_, (void **)&d_idx_, sizeof(uint64_t) * draws_cap_));
        }
        idx.resize(cnt);
        if (cnt) {
            CK(cudaMemcpyAsync(d_chosen_, chosen.data(), sizeof(double) * cnt, cudaMemcpyHostToDevice, stream_));
            time_begin();
            CK(launch_resolve_draws(cols_[c].buf, d_leaf_ + k * nleaves, d_block_ + k * nblocks, n_, d_chosen_, cnt,
                                    reinterpret_cast<unsigned long long *>(d_idx_), 0.0, stream_));
            time_end(stats.read_ms);
            stats.kernel_launches++;
            CK(cudaMemcpyAsync(idx.data(), d_idx_, sizeof(uint64_t) * cnt, cudaMemcpyDeviceToHost, stream_));
            CK(cudaStreamSynchronize(stream_));
        }
        for (size_t j = 0; j < cnt; ++j) {
            if (j > 0 && idx[j] == idx[j - 1]) state_counts.back().second++;
            else state_counts.push_back(std::make_pair(idx[j], (size_t)1));
        }
        ++k;
    }
    // route bits: qubit j -> classical bit cbits[j] (support.rs:50-75, vectorstate.rs:135-148)
    uint64_t m = 0;
    for (size_t j = 0; j < ncbits; ++j) m |= 1ull << cbits[j];
    const uint64_t mask = ~m;
    // one 256-entry table per byte of the basis index instead of a loop over the qubits per outcome
    const int nbytes = (n_ + 7) / 8;
    std::vector<uint64_t> route((size_t)nbytes * 256, 0);
    for (int b = 0; b < nbytes; ++b)
        for (int v = 1; v < 256; ++v) {
            const int low = v & (v - 1), bit = 8 * b + __builtin_ctz((unsigned)v);      // index bit `bit` <-> qubit n-1-bit
            route[(size_t)b * 256 + v] = route[(size_t)b * 256 + low] | (bit < n_ ? 1ull << cbits[n_ - 1 - bit] : 0ull);
        }
    size_t off = 0;
    for (const auto &sc : state_counts) {
        uint64_t word = 0;
        for (int b = 0; b < nbytes; ++b) word |= route[(size_t)b * 256 + ((sc.first >> (8 * b)) & 255ull)];
        for (size_t j = off; j < off + sc.second; ++j) res[j] = (res[j] & mask) | word;
        off += sc.second;
    }
    if (collapse) {
        // vectorstate.rs:150-158 allocates a dense (2^n, n_distinct) matrix of unit vectors;
        // here the collapsed columns are kept as basis indices until a gate touches them
        for (Column &c : cols_)
            if (c.buf) release_column(c.buf);
        cols_.clear();
        for (const auto &sc : state_counts) {
            Column c;
            c.basis = true; c.basis_idx = sc.first; c.count = sc.second;
            cols_.push_back(c);
        }
    }
    return Q1T_OK;
}

// vectorstate.rs:402-408
int DeviceVectorState::reset(size_t bit, q1t_rng rng)
{
    { const int rcp = resolve_pending_remap(); if (rcp) return rcp; }      // a recorded qubit remap: the peers still read this shard through it
    std::vector<uint64_t> m(shots_ ? shots_ : 1, 0);
    int rc = measure_into(bit, 0, m.data(), shots_, rng, true);
    if (rc) return rc;
    std::vector<uint8_t> control(shots_ ? shots_ : 1);
    for (size_t j = 0; j < shots_; ++j) control[j] = m[j] != 0;
    static const double X[8] = { 0, 0, 1, 0, 1, 0, 0, 0 };
    return apply_conditional_gate(control.data(), shots_, X, 2, &bit, 1, "X");
}

// vectorstate.rs:410-415
int DeviceVectorState::reset_all()
{
    { const int rcp = resolve_pending_remap(); if (rcp) return rcp; }      // a recorded qubit remap: the peers still read this shard through it
    int rc = ensure_device();
    if (rc) return rc;
    queue_.clear();
    queue_cols_.clear();
    pending_scale_ = 1.0;
    CK(cudaStreamSynchronize(stream_));
    for (Column &c : cols_)
        if (c.buf) release_column(c.buf);
    cols_.clear();
    for (int l = 0; l < n_; ++l) perm_[l] = l;
    return init_zero_state();
}

int DeviceVectorState::counts(size_t *out)
{
    for (size_t c = 0; c < cols_.size(); ++c) out[c] = cols_[c].count;
    return Q1T_OK;
}

int DeviceVectorState::read_amplitudes(size_t col, size_t offset, size_t len, double *out)
{
    if (col >= cols_.size() || offset + len > ((size_t)1 << n_)) return fail(Q1T_ERR_INVALID_ARGUMENT, "amplitude range out of bounds");
    int rc = flush();
    if (rc) return rc;
    if (cols_[col].basis) {
        std::memset(out, 0, sizeof(double) * 2 * len);
        if (cols_[col].basis_idx >= offset && cols_[col].basis_idx < offset + len) out[2 * (cols_[col].basis_idx - offset)] = 1.0;
        return Q1T_OK;
    }
    CK(cudaMemcpyAsync(out, cols_[col].buf + offset, sizeof(double2) * len, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    return Q1T_OK;
}

int DeviceVectorState::write_amplitudes(size_t col, size_t offset, size_t len, const double *in)
{
    if (col >= cols_.size() || offset + len > ((size_t)1 << n_)) return fail(Q1T_ERR_INVALID_ARGUMENT, "amplitude range out of bounds");
    int rc = flush();
    if (rc) return rc;
    rc = materialize(cols_[col]);
    if (rc) return rc;
    CK(cudaMemcpyAsync(cols_[col].buf + offset, in, sizeof(double2) * len, cudaMemcpyHostToDevice, stream_));
    CK(cudaStreamSynchronize(stream_));
    return Q1T_OK;
}

int DeviceVectorState::set_option(const char *key, long value)
{
    if (!key) return fail(Q1T_ERR_INVALID_ARGUMENT, "NULL option key");
    if (!std::strcmp(key, "tile_bits")) {
        if (value < 8 || value > kMaxTileBits) return fail(Q1T_ERR_INVALID_ARGUMENT, "tile_bits must be in 8..13");
        int rc = run_queue();
        if (rc) return rc;
        tile_bits_ = value;
        return Q1T_OK;
    }
    if (!std::strcmp(key, "coalesce_bits")) {
        if (value < 2 || value > 3) return fail(Q1T_ERR_INVALID_ARGUMENT, "coalesce_bits must be 2 or 3");
        int rc = run_queue();
        if (rc) return rc;
        coalesce_bits_ = value;
        return Q1T_OK;
    }
    if (!std::strcmp(key, "balance")) {
        int rc = run_queue();
        if (rc) return rc;
        balance_ = value;
        return Q1T_OK;
    }
    if (!std::strcmp(key, "inplace_relabel")) {
        inplace_relabel_ = value;
        return Q1T_OK;
    }
    if (!std::strcmp(key, "fuse_leaf_totals")) {
        fuse_leaf_totals_ = value != 0;
        return Q1T_OK;
    }
    if (!std::strcmp(key, "sparse_c2")) {
        int rc = run_queue();
        if (rc) return rc;
        sparse_c2_ = value != 0;
        return Q1T_OK;
    }
    if (!std::strcmp(key, "track_support")) {
        track_support_ = value != 0;
        return Q1T_OK;
    }
    if (!std::strcmp(key, "tma")) {
        tma_ = value != 0;
        return Q1T_OK;
    }
    if (!std::strcmp(key, "graphs")) {
        graphs_ = value != 0;
        return Q1T_OK;
    }
    if (!std::strcmp(key, "mid_relabel")) {
        int rc = run_queue();
        if (rc) return rc;
        mid_relabel_ = value;
        return Q1T_OK;
    }
    if (!std::strcmp(key, "fused_remap")) {
        fused_remap_ = value != 0;
        return Q1T_OK;
    }
    if (!std::strcmp(key, "direct")) {
        direct_ = value != 0;
        return Q1T_OK;
    }
    if (!std::strcmp(key, "prefetch_ahead")) {
        prefetch_ahead_ = value < 0 ? 0 : value;
        return Q1T_OK;
    }
    if (!std::strcmp(key, "fuse")) {
        int rc = run_queue();
        if (rc) return rc;
        rc = canonicalize();
        if (rc) return rc;
        fuse_ = value != 0;
        return Q1T_OK;
    }
    return fail(Q1T_ERR_INVALID_ARGUMENT, std::string("unknown option ") + key);
}

}  // namespace q1t
