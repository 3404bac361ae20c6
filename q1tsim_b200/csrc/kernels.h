// kernels.h -- host-callable launchers of kernels.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "program.h"

namespace q1t {

constexpr int kMaxGenericBits = 6;       // dense fallback: up to 64x64 matrices
constexpr int kCanonLeafBits = 10;       // canonical reduction: leaves of 1024 amplitudes
constexpr unsigned kCanonBlock = 1024;   // leaves per chained block

struct GenericGateArgs {
    int sorted_pos[kMaxGenericBits];                  // target positions ascending
    unsigned long long offs[1 << kMaxGenericBits];    // index offset of gate-matrix index g
    unsigned long long cmask;                         // control positions that must be 1
};

// Fused remap read (ladder kernel, dense sweeps): a pending trade of k rank bits gb[j] with k index bits lp[j] of the
// shards is not run as a swap pass of its own; the sweep that follows gathers its source tiles from the peers instead --
// element l of rank r is read from rank r[gb[j] := l_lp[j]] at index l[lp[j] := r_gb[j]] (group_swap_kernel's map) --
// and writes its result into this rank's other registered buffer.  Local tiles are swept while the remote ones arrive
// over NVLink.
constexpr int kMaxRemapBits = 4;       // rank bits traded per remap
struct RemoteGather {
    int k, rank, P;                                  // k = 0: off
    int gb[kMaxRemapBits], lp[kMaxRemapBits];
    void *const *peer_buf;                           // [P][2], device memory (PeerGroup::d_peer_buf)
    const unsigned long long *my_mail;               // this rank's mailbox: [2 r + 1] = buffer index rank r published
};

// d_leaf_out: when prog.leaf_fuse is set (ladder kernel, staged store), receives the canonical leaf totals
// of the written column(s), 2^(n-10) doubles per column (what launch_leaf_totals would compute afterwards)
cudaError_t launch_sweep(const SweepProgram &prog, const double2 *const *d_src_cols, double2 *const *d_dst_cols,
                         int ncols, const PhaseTab *d_ptabs, const unsigned long long *d_gen_idx, cudaStream_t stream,
                         double *d_leaf_out = nullptr, const double2 *const *h_src_cols = nullptr,
                         const SweepProgram *d_prog = nullptr, const RemoteGather *remote = nullptr);
// true if launch_sweep() can run this program with a RemoteGather (dense ladder sweep, cp.async loads)
bool sweep_can_gather_remote(const SweepProgram &prog);
// d_prog: device-resident copy of `prog` (the caller keeps it alive and in sync): kernels other than the ladder
// kernel then read the program from there and nothing is uploaded -- batches replayed as CUDA graphs
// h_src_cols: host copy of the source column pointers, needed (only) by programs in TMA layout (prog.tma_nreq > 0)
// to encode the tensor maps; true if the driver offers cuTensorMapEncodeTiled
bool tma_available();
// true while more than one state of this process is alive on the device: their program uploads into the one constant
// bank are serialised, and graph capture of sweep batches is off there.  States register / unregister their stream.
bool cprog_device_shared(int dev);
void cprog_stream_register(int dev);
void cprog_stream_unregister(int dev);
// can the tile description of a program in TMA layout be encoded for these source columns?
bool tma_can_encode(const SweepProgram &prog, const double2 *const *h_src_cols, int ncols);
// true if launch_sweep() runs this program in the persistent ladder kernel (the one that honours
// SweepProgram::sup_mask / sup_mode)
bool sweep_uses_ladder_kernel(const SweepProgram &prog);
cudaError_t launch_generic_gate(double2 *const *d_cols, int ncols, int n, int k, const GenericGateArgs &g,
                                const double2 *d_mat, cudaStream_t stream);
// dense blocks on kMaxGenericBits .. kMaxBigGenericBits targets: a CTA stages several groups of 2^k amplitudes in shared
// memory, thread i accumulates output i of every staged group from the TRANSPOSED matrix (d_matT[h * 2^k + i])
constexpr int kMaxBigGenericBits = 10;
struct GenericBigArgs {
    int n, k;
    int pos[kMaxBigGenericBits];          // position of the j-th listed target (bit k-1-j of the matrix index)
    int sorted_pos[kMaxBigGenericBits];   // the same positions ascending
    unsigned long long cmask;             // control positions that must be 1
};
cudaError_t launch_generic_gate_big(double2 *const *d_cols, int ncols, const GenericBigArgs &g, const double2 *d_matT, cudaStream_t stream);
cudaError_t launch_leaf_totals(const double2 *const *d_cols, int ncols, double *d_leaf, int n,
                               unsigned long long mask, unsigned long long want, cudaStream_t stream);
cudaError_t launch_scan(double *d_leaf, double *d_block, double *d_totals, int ncols, int n, cudaStream_t stream);
cudaError_t launch_block_scan(double *d_leaf, double *d_block, int ncols, int n, cudaStream_t stream);
cudaError_t launch_resolve_draws(const double2 *d_col, const double *d_leaf, const double *d_block, int n,
                                 const double *d_chosen, unsigned long long ndraws, unsigned long long *d_idx,
                                 double base, cudaStream_t stream, double base0 = 0.0);
cudaError_t launch_scale2(const double2 *d_in, double2 *d_out0, double2 *d_out1, int n, double f0, double f1, cudaStream_t stream);
cudaError_t launch_collapse(const double2 *d_in, double2 *d_out0, double2 *d_out1, int n, int bitpos, double f0,
                            double f1, cudaStream_t stream);
cudaError_t launch_product_state(double2 *d_col, int n, const double2 *d_coefs, cudaStream_t stream);
cudaError_t launch_peer_swap(double2 *d_mine, double2 *d_theirs, int n, int L, int a, cudaStream_t stream);
// peer group: device-side barrier over mapped mailboxes, and the multi-bit qubit remap (kernels.cu)
struct GroupRemapArgs {
    int n, k, P, rank;               // index bits of the shard, bits traded, ranks, this rank
    int gb[kMaxRemapBits];           // rank bits
    int lp[kMaxRemapBits];           // index-bit positions of the shard they trade with
    int split;                       // index bit (not among lp) that splits the pairs between the two ranks of a pair
    int ins[kMaxRemapBits + 1];      // lp and split, ascending
    int interleave;                  // 1: consecutive CTAs serve different partners; 0: one partner after the other
};
cudaError_t launch_group_barrier(unsigned long long *const *d_peer_mail, unsigned long long *d_my_mail, int P, int rank,
                                 unsigned long long epoch, unsigned long long cur, cudaStream_t stream);
cudaError_t launch_group_swap(double2 *d_mine, void *const *d_peer_buf, const unsigned long long *d_my_mail, const GroupRemapArgs &a,
                              cudaStream_t stream);
cudaError_t launch_group_gather(const double2 *d_mine, double2 *d_dst, const RemoteGather &rg, int n, cudaStream_t stream);
cudaError_t launch_set_basis(double2 *d_col, unsigned long long idx, cudaStream_t stream);

}  // namespace q1t
