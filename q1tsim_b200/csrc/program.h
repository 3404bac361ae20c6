// program.h -- the "sweep program": what one fused pass over the state does.
//
// A sweep reads every amplitude of the selected columns once and writes it
// once.  A CTA owns a tile of 2^T amplitudes (T tile bits = the low
// coalescing bits + up to T-3 arbitrary index bits) staged in shared memory;
// inside the tile, work proceeds in rounds: in a round each thread holds the
// 2^R amplitudes spanned by R "register bits" and applies a list of ops to
// them without touching memory.  Shared by the host planner and the kernels.
#pragma once
#include <stdint.h>

namespace q1t {

constexpr int kMaxBits = 40;        // index bits of one column (one GPU's shard)
constexpr int kMaxTileBits = 13;    // 2^13 * 16 B = 128 KiB of shared memory
#ifndef Q1T_REG_BITS
#define Q1T_REG_BITS 5
#endif
constexpr int kRegBits = Q1T_REG_BITS;   // amplitudes per thread = 2^kRegBits (4 or 5)
static_assert(Q1T_REG_BITS == 4 || Q1T_REG_BITS == 5, "4 or 5 register bits");
constexpr int kSlots = 1 << kRegBits;
constexpr int kMaxThrBits = kMaxTileBits - kRegBits;   // 9 -> 512 threads (8 -> 256 with 5 register bits)
constexpr int kMaxRounds = 24;
constexpr int kMaxOps = 96;
constexpr int kMaxPhase = 48;
constexpr int kThrLoBits = 4;       // per-thread phase factor = lo[tid & 15] * hi[tid >> 4]
constexpr int kHiEntries = 1 << (kMaxThrBits - kThrLoBits);
constexpr int kMaxRuns = 9;         // bit-deposit runs (mask, shift) of a thread index
constexpr int kOuterChunkBits = 6;  // outer tile index -> address via 6-bit lookup tables
constexpr int kOuterChunks = 5;
constexpr int kMaxTmaReq = 16;      // TMA requests per tile (one when the tile's line bits form <= 3 runs)
constexpr int kMaxTmaCols = 4;      // columns per launch that can be described by tensor maps in the kernel parameters

enum OpKind : uint8_t {
    OP_G1_GENERIC = 0,   // dense complex 2x2 on slot bit j
    OP_G1_HADAMARD = 1,  // c * [[1,1],[1,-1]], c real
    OP_G1_ANTIDIAG = 2,  // [[0,m01],[m10,0]]
    OP_G1_SWAPX = 3,     // [[0,1],[1,0]]
    OP_PHASE = 4,        // multiply slots with bit j set by a per-thread phase
    OP_G1_DIAG = 5,      // [[m00,0],[0,m11]] (controlled-diagonal fallback)
    OP_H_UNNORM = 6,     // [[1,1],[1,-1]]; the 1/sqrt(2) is folded into SweepProgram::scale
    OP_PHASE_H = 7,      // OP_PHASE followed by OP_H_UNNORM on the same slot bit, fused
    OP_LINPHASE = 8,     // separable diagonal: every slot s times F(thread, tile) * prod_{j: s_j = 1} m[j]
};

enum RoundKind : uint8_t { ROUND_GENERIC = 0, ROUND_PH = 1 };

constexpr uint8_t kFlagC0 = 0x80;
struct OpDesc {          // 112 bytes
    uint8_t kind;
    uint8_t j;           // slot bit
    uint8_t flags;       // PHASE: bit0..kRegBits-2 = partner q[i] is non-unit, kFlagC0 = has c0 factor
    uint8_t pad0;
    uint32_t cslot;      // G1: slot bits that must be 1 (controls that are register bits)
    uint64_t cmask;      // G1: virtual-index bits (outer<<T | tile-local) that must be 1, register bits excluded
    double m[10];        // G1: m00,m01,m10,m11 (re,im).  PHASE: q[0..kRegBits-2] partner factors (other slot bits ascending), m[8..9] = c0 factor
    uint32_t phase_id;   // PHASE: row of the phase tables
    uint32_t pad1;
    uint64_t pad2;
};

struct BitRun {          // contributes ((v & mask) << shift) (>> -shift if negative) to a deposited index
    uint32_t mask;
    int32_t shift;
};

struct RoundDesc {
    uint8_t reg_tb[kRegBits];       // tile bit of slot bit j
    uint8_t thr_tb[kMaxThrBits];    // tile bit of thread-index bit i (ascending)
    uint8_t nruns;
    uint8_t kind;                   // ROUND_GENERIC: op interpreter; ROUND_PH: straight-line PHASE_H ladder
    uint8_t nsteps;                 // ROUND_PH: ops op_begin..op_begin+nsteps-1 act on slot bits 0..nsteps-1
    uint8_t sync_before;            // 0 none, 1 __syncwarp (data only moves inside warps), 2 __syncthreads
    uint8_t pad1[3];
    uint32_t zmask;                 // support tracking: tile-local bits that are thread bits of this round and still
                                    // pinned to the basis value -> a thread that differs there holds only zeros
    uint32_t smask;                 // same for the register bits: slot bits still pinned when the round starts -> a slot
                                    // that differs from the basis value there is zero and is not read
    BitRun runs[kMaxRuns];          // tid -> tile-local index of the thread
    uint32_t sw_slot[kSlots];       // byte offset (swizzled index * 16) contributed by slot s
    uint16_t op_begin, op_end;
};

struct SweepProgram {
    int32_t n;            // index bits per column
    int32_t T;            // tile bits
    int32_t TB;           // thread bits = T - kRegBits (threads per CTA = 2^TB)
    int32_t n_outer;      // n - T
    int32_t nrounds, nops, nphase;
    int32_t relabel;      // 1 if dst positions differ from src positions (out-of-place only)
    int32_t generate;     // 1: the source column is a basis state |gen_idx[col]>, nothing is read
    int32_t ld_nruns, st_nruns;
    int32_t prefetch_ahead;   // TMA mode, >0: the tile this many grid strides ahead is prefetched into L2 when a tile load is issued
    int32_t direct_load;      // round 0 reads its amplitudes straight from global memory (no staging pass)
    int32_t direct_store;     // the last round writes its amplitudes straight to global memory
    int32_t dl_nruns, ds_nruns;
    uint64_t tile_mask_src;   // source positions of the tile bits
    int32_t reserved0;
    int32_t coalesce;         // low index bits kept contiguous in every tile (3 = 128 B, 2 = 64 B)
    double scale;         // applied to every amplitude at the store (deferred Hadamard normalisation)
    // Support tracking (ladder kernel): every amplitude whose index differs from the column's basis
    // index gen_idx[col] in a bit of sup_mask (source positions) is zero and NOT read.  sup_mode 1:
    // tiles outside the support are skipped altogether; 2: they are written as zeros (last sweep of
    // a batch: the column leaves dense); 0: no tracking.
    uint64_t sup_mask;
    int32_t sup_mode;
    int32_t leaf_fuse;    // staged (relabelling) store laid out one canonical leaf per warp: the store pass also
                          // produces the leaf totals of abs(amp)^2 in the canonical order
    double gen_scale;     // value of the basis element of a generated input (1, or the normalisation of the whole batch of sweeps)
    // tile bits are numbered by ascending source position; outer bits likewise
    uint8_t tsrc[kMaxTileBits + 3], tdst[kMaxTileBits + 3];
    uint8_t osrc[kMaxBits], odst[kMaxBits];
    // tile base address of outer index o = OR over chunks c of o_src[c][(o >> 6c) & 63]
    uint64_t o_src[kOuterChunks][1 << kOuterChunkBits];
    uint64_t o_dst[kOuterChunks][1 << kOuterChunkBits];
    // the same maps for a walk index whose bits are the outer bits in ascending DESTINATION order: a kernel
    // that walks w = 0, 1, 2, .. writes its tiles as a few sequential streams (broadcast sweeps: the reads
    // are negligible there, the write pattern is everything)
    uint64_t w_src[kOuterChunks][1 << kOuterChunkBits];
    uint64_t w_dst[kOuterChunks][1 << kOuterChunkBits];
    // load: element e = tid | i<<TB lives at source offset dep(tid) | ld_hi[i], tile index e
    BitRun ld_runs[kMaxRuns];          // tid -> source offset (shift < 64)
    uint64_t ld_hi[kSlots];
    uint32_t ld_sw_hi[kSlots];         // swizzled byte offset of (i << TB)
    // store: element f = tid | i<<TB (ascending destination position)
    uint8_t st_tb[kMaxTileBits + 3];   // tile bit whose destination position is the f-th smallest
    BitRun st_runs[kMaxRuns];          // tid -> destination offset
    BitRun st_lruns[kMaxRuns];         // tid -> tile index
    uint64_t st_off_hi[kSlots];        // destination offset of the high part of f
    uint32_t st_l_hi[kSlots];          // swizzled byte offset of the tile index of the high part of f
    // direct paths: thread -> offset runs and slot -> offset tables of round 0 (source layout)
    // and of the last round (destination layout)
    BitRun dl_runs[kMaxRuns];
    uint64_t dl_slot[kSlots];
    BitRun ds_runs[kMaxRuns];
    uint64_t ds_slot[kSlots];
    RoundDesc rounds[kMaxRounds];
    OpDesc ops[kMaxOps];
    // TMA tile loads (ladder kernel, dense sweeps; planner.cpp apply_tma_layout).  tma_nreq > 0: the tile is
    // kept in shared memory in the order a cp.async.bulk.tensor load delivers it -- smem index m = the tile-local
    // index with its bits permuted (tile bit i -> smem bit tma_pi[i]), 16-byte chunk bits 0..2 xor-ed with the
    // line bits 3..5 (CU_TENSOR_MAP_SWIZZLE_128B) -- and every shared-memory offset table of the program
    // (rounds[].runs, rounds[].sw_slot, st_lruns, st_l_hi) is expressed in that order.  The column is described
    // to the TMA unit as a 5-d tensor of doubles: d0 = the 16 doubles of a 128-byte line, d1..d3 = runs of
    // consecutive index bits taken in smem order (box tma_box[i], byte stride tma_gstride[i-1]), d4 = the line
    // index (stride 128 B), which carries the tile base; request q of a tile adds tma_req_line[q] lines.
    int32_t tma_nreq;
    uint32_t tma_req_bytes;
    uint32_t tma_box[5];
    uint32_t tma_pad;
    uint64_t tma_gstride[4];
    uint64_t tma_gdim[5];
    uint64_t tma_req_line[kMaxTmaReq];
    uint8_t tma_pi[kMaxTileBits + 3];
    // ld_hi / st_off_hi / ds_slot in bytes (fill_byte_tables, right before a launch): the ladder kernel adds them to a
    // byte address instead of scaling 32 element indices per tile
    uint64_t ld_hi_b[kSlots], st_off_hi_b[kSlots], ds_slot_b[kSlots];
};

inline void fill_byte_tables(SweepProgram &P)
{
    for (int i = 0; i < kSlots; ++i) {
        P.ld_hi_b[i] = P.ld_hi[i] << 4;
        P.st_off_hi_b[i] = P.st_off_hi[i] << 4;
        P.ds_slot_b[i] = P.ds_slot[i] << 4;
    }
}

// per PHASE op, in global memory
struct alignas(16) PhaseTab {
    double lo[2 * (1 << kThrLoBits)];                 // complex factors indexed by tid & 15 (16-byte aligned: read as double2)
    double hi[2 * (1 << (kMaxThrBits - kThrLoBits))]; // complex factors indexed by tid >> 4
    double base;                      // half-turns: constant part of the angle for slots with bit j set
    double outer_coef[kMaxBits];      // half-turns per outer-index bit
    double pad;
};
static_assert(sizeof(PhaseTab) % 16 == 0, "PhaseTab must keep 16-byte alignment in arrays");

#ifdef __CUDACC__
#define Q1T_HD __host__ __device__
#else
#define Q1T_HD
#endif
Q1T_HD inline constexpr uint32_t tile_swizzle(uint32_t l) {
    return l ^ (((l >> 3) ^ (l >> 6) ^ (l >> 9) ^ (l >> 12)) & 7u);
}
// shared-memory order of a TMA-loaded tile (CU_TENSOR_MAP_SWIZZLE_128B on 1024-byte aligned memory)
Q1T_HD inline constexpr uint32_t tma_swizzle(uint32_t m) {
    return m ^ ((m >> 3) & 7u);
}

}  // namespace q1t
