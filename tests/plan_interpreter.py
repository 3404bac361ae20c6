"""CPU interpreter of the engine's sweep programs (test infrastructure).

The fusion planner (q1tsim_b200/csrc/planner.cpp) turns a gate list into `SweepProgram`s (csrc/program.h): tiles,
rounds, register-level ops, phase tables, relabelling.  On the GPU those programs are executed by `sweep_kernel` /
`ladder_kernel`; here the same structs (obtained through the `q1t_plan_dump` test hook, no device needed) are executed
with numpy, following the op semantics of csrc/kernels.cu line by line (cited below).  Two uses:
  * the planner -- the largest piece of host logic -- is checked against the CPU oracle without a GPU;
  * the address tables the kernels read (bit-deposit runs, slot tables, outer tables) are checked against the
    logical layout (tsrc / tdst / osrc / odst, reg_tb / thr_tb) they are derived from.
"""
import ctypes as C

import numpy as np

from q1tsim_b200 import engine as E

K_REG, K_MAX_TILE, K_MAX_BITS, K_MAX_ROUNDS, K_MAX_OPS, K_MAX_RUNS = 5, 13, 40, 24, 96, 9
K_SLOTS, K_MAX_THR, K_THR_LO, K_CHUNKS, K_CHUNK_BITS = 1 << K_REG, K_MAX_TILE - K_REG, 4, 5, 6
OP_G1_GENERIC, OP_G1_HADAMARD, OP_G1_ANTIDIAG, OP_G1_SWAPX, OP_PHASE, OP_G1_DIAG, OP_H_UNNORM, OP_PHASE_H, OP_LINPHASE = range(9)
ROUND_PH = 1
K_MAX_TMA_REQ = 16
FLAG_C0 = 0x80


class BitRun(C.Structure):
    _fields_ = [("mask", C.c_uint32), ("shift", C.c_int32)]


class OpDesc(C.Structure):
    _fields_ = [("kind", C.c_uint8), ("j", C.c_uint8), ("flags", C.c_uint8), ("pad0", C.c_uint8), ("cslot", C.c_uint32),
                ("cmask", C.c_uint64), ("m", C.c_double * 10), ("phase_id", C.c_uint32), ("pad1", C.c_uint32), ("pad2", C.c_uint64)]


class RoundDesc(C.Structure):
    _fields_ = [("reg_tb", C.c_uint8 * K_REG), ("thr_tb", C.c_uint8 * K_MAX_THR), ("nruns", C.c_uint8), ("kind", C.c_uint8),
                ("nsteps", C.c_uint8), ("sync_before", C.c_uint8), ("pad1", C.c_uint8 * 3), ("zmask", C.c_uint32),
                ("smask", C.c_uint32), ("runs", BitRun * K_MAX_RUNS), ("sw_slot", C.c_uint32 * K_SLOTS),
                ("op_begin", C.c_uint16), ("op_end", C.c_uint16)]


_TAB = (C.c_uint64 * (1 << K_CHUNK_BITS)) * K_CHUNKS


class SweepProgram(C.Structure):
    _fields_ = [("n", C.c_int32), ("T", C.c_int32), ("TB", C.c_int32), ("n_outer", C.c_int32), ("nrounds", C.c_int32),
                ("nops", C.c_int32), ("nphase", C.c_int32), ("relabel", C.c_int32), ("generate", C.c_int32),
                ("ld_nruns", C.c_int32), ("st_nruns", C.c_int32), ("prefetch_ahead", C.c_int32), ("direct_load", C.c_int32),
                ("direct_store", C.c_int32), ("dl_nruns", C.c_int32), ("ds_nruns", C.c_int32), ("tile_mask_src", C.c_uint64),
                ("reserved0", C.c_int32), ("coalesce", C.c_int32), ("scale", C.c_double), ("sup_mask", C.c_uint64),
                ("sup_mode", C.c_int32), ("leaf_fuse", C.c_int32), ("gen_scale", C.c_double),
                ("tsrc", C.c_uint8 * (K_MAX_TILE + 3)), ("tdst", C.c_uint8 * (K_MAX_TILE + 3)),
                ("osrc", C.c_uint8 * K_MAX_BITS), ("odst", C.c_uint8 * K_MAX_BITS),
                ("o_src", _TAB), ("o_dst", _TAB), ("w_src", _TAB), ("w_dst", _TAB),
                ("ld_runs", BitRun * K_MAX_RUNS), ("ld_hi", C.c_uint64 * K_SLOTS), ("ld_sw_hi", C.c_uint32 * K_SLOTS),
                ("st_tb", C.c_uint8 * (K_MAX_TILE + 3)), ("st_runs", BitRun * K_MAX_RUNS), ("st_lruns", BitRun * K_MAX_RUNS),
                ("st_off_hi", C.c_uint64 * K_SLOTS), ("st_l_hi", C.c_uint32 * K_SLOTS),
                ("dl_runs", BitRun * K_MAX_RUNS), ("dl_slot", C.c_uint64 * K_SLOTS),
                ("ds_runs", BitRun * K_MAX_RUNS), ("ds_slot", C.c_uint64 * K_SLOTS),
                ("rounds", RoundDesc * K_MAX_ROUNDS), ("ops", OpDesc * K_MAX_OPS),
                ("tma_nreq", C.c_int32), ("tma_req_bytes", C.c_uint32), ("tma_box", C.c_uint32 * 5), ("tma_pad", C.c_uint32),
                ("tma_gstride", C.c_uint64 * 4), ("tma_gdim", C.c_uint64 * 5), ("tma_req_line", C.c_uint64 * K_MAX_TMA_REQ),
                ("tma_pi", C.c_uint8 * (K_MAX_TILE + 3)),
                ("ld_hi_b", C.c_uint64 * K_SLOTS), ("st_off_hi_b", C.c_uint64 * K_SLOTS), ("ds_slot_b", C.c_uint64 * K_SLOTS)]


class PhaseTab(C.Structure):
    _fields_ = [("lo", C.c_double * (2 << K_THR_LO)), ("hi", C.c_double * (2 << (K_MAX_THR - K_THR_LO))), ("base", C.c_double),
                ("outer_coef", C.c_double * K_MAX_BITS), ("pad", C.c_double)]


def check_layout():
    out = (C.c_size_t * 10)()
    L = E.lib()
    L.q1t_plan_layout.argtypes = [C.POINTER(C.c_size_t), C.c_size_t]
    L.q1t_plan_layout(out, 10)
    want = [C.sizeof(SweepProgram), C.sizeof(PhaseTab), C.sizeof(OpDesc), C.sizeof(RoundDesc), K_REG, K_MAX_TILE, K_MAX_BITS,
            K_MAX_ROUNDS, K_MAX_OPS, K_MAX_RUNS]
    assert list(out) == want, (list(out), want)


def plan(nr_bits, gates, tile_bits=12, coalesce_bits=3, balance=0, max_sweeps=256, max_ptabs=4096):
    """gates: [(matrix, bits)] -> ([(SweepProgram, [PhaseTab])], perm) with perm[l] = physical position of logical index bit l"""
    check_layout()
    L = E.lib()
    mats = np.concatenate([np.ascontiguousarray(np.asarray(m, dtype=np.complex128)).ravel() for m, _ in gates]).view(np.float64)
    dims = (C.c_size_t * len(gates))(*[np.asarray(m).shape[0] for m, _ in gates])
    bits = (C.c_size_t * sum(len(b) for _, b in gates))(*[int(x) for _, b in gates for x in b])
    nb = (C.c_size_t * len(gates))(*[len(b) for _, b in gates])
    progs = (SweepProgram * max_sweeps)()
    ptabs = (PhaseTab * max_ptabs)()
    counts = (C.c_int * max_sweeps)()
    perm = (C.c_int * nr_bits)()
    L.q1t_plan_dump.restype = C.c_int
    L.q1t_plan_dump.argtypes = [C.c_size_t, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                C.POINTER(C.c_size_t), C.c_long, C.c_long, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                C.POINTER(C.c_int), C.POINTER(C.c_int)]
    k = L.q1t_plan_dump(nr_bits, len(gates), mats.ctypes.data_as(C.POINTER(C.c_double)), dims, bits, nb, tile_bits, coalesce_bits,
                        balance, C.cast(progs, C.c_void_p), max_sweeps, C.cast(ptabs, C.c_void_p), max_ptabs, counts, perm)
    if k < 0:
        raise RuntimeError("q1t_plan_dump failed: %d" % k)
    out, at = [], 0
    for i in range(k):
        out.append((progs[i], [ptabs[at + t] for t in range(counts[i])]))
        at += counts[i]
    return out, list(perm)


def _deposit(v, positions):
    """bit i of v -> bit positions[i]"""
    r = np.zeros_like(v)
    for i, p in enumerate(positions):
        r |= ((v >> i) & 1) << int(p)
    return r


def _run_bits(v, runs, nruns):
    """kernels.cu run_bits: sum over runs of ((v & mask) << shift) (>> -shift)"""
    r = np.zeros_like(v, dtype=np.uint64)
    for k in range(nruns):
        m = (v & np.uint64(runs[k].mask)).astype(np.uint64)
        sh = runs[k].shift
        r |= (m << np.uint64(sh)) if sh >= 0 else (m >> np.uint64(-sh))
    return r


def tile_swizzle(l):
    return l ^ (((l >> 3) ^ (l >> 6) ^ (l >> 9) ^ (l >> 12)) & 7)


def run_sweep(P, ptabs, psi):
    """one sweep over a column in physical layout; returns the new column (kernels.cu sweep_kernel / ladder_kernel)"""
    n, T, TB, n_outer = P.n, P.T, P.TB, P.n_outer
    assert psi.shape == (1 << n,)
    tsrc, tdst = [P.tsrc[i] for i in range(T)], [P.tdst[i] for i in range(T)]
    osrc, odst = [P.osrc[i] for i in range(n_outer)], [P.odst[i] for i in range(n_outer)]
    l = np.arange(1 << T, dtype=np.int64)
    o = np.arange(1 << n_outer, dtype=np.int64)
    src = _deposit(o, osrc)[:, None] | _deposit(l, tsrc)[None, :]
    dst = _deposit(o, odst)[:, None] | _deposit(l, tdst)[None, :]
    a = psi[src]                                              # (tiles, 2^T) by tile-local index
    for r in range(P.nrounds):
        R = P.rounds[r]
        reg_tb = [R.reg_tb[j] for j in range(K_REG)]
        thr_tb = [R.thr_tb[i] for i in range(TB)]
        slot = np.zeros_like(l)
        for j, tb in enumerate(reg_tb):
            slot |= ((l >> tb) & 1) << j
        tid = np.zeros_like(l)
        for i, tb in enumerate(thr_tb):
            tid |= ((l >> tb) & 1) << i
        v = (o[:, None] << T) | l[None, :]                    # virtual index: outer << T | tile-local

        def F_of(op):
            pt = ptabs[op.phase_id]
            lo = np.array(pt.lo[:]).view(np.complex128)
            hi = np.array(pt.hi[:]).view(np.complex128)
            ang = pt.base + sum(pt.outer_coef[i] * ((o >> i) & 1) for i in range(n_outer))
            tilef = np.exp(1j * np.pi * np.asarray(ang, dtype=np.float64))
            return (lo[tid & 15] * hi[tid >> K_THR_LO])[None, :] * (tilef[:, None] if n_outer else tilef)

        for k in range(R.op_begin, R.op_end):
            op = P.ops[k]
            ladder = R.kind == ROUND_PH
            J = (K_REG - R.nsteps + (k - R.op_begin)) if ladder else op.j
            if ladder:
                assert op.kind == OP_PHASE_H and op.j == J and (op.flags >> J) == 0
            tb = reg_tb[J]
            L0 = l[((l >> tb) & 1) == 0]
            L1 = L0 | (1 << tb)
            x, y = a[:, L0], a[:, L1]
            kind = op.kind
            if kind in (OP_PHASE, OP_PHASE_H, OP_LINPHASE):
                F = F_of(op)
                if kind == OP_LINPHASE:
                    # linphase_apply: every slot times F * prod_{j: slot bit j set} m[j]
                    f = F.copy()
                    for j in range(K_REG):
                        q = complex(op.m[2 * j], op.m[2 * j + 1])
                        f = np.where(((slot >> j) & 1)[None, :] == 1, f * q, f)
                    a = a * f
                    continue
                # phase_factors: partner i of the compacted index (slot bits other than J, ascending)
                f = F[:, L0] if F.shape[0] == a.shape[0] else np.broadcast_to(F, a.shape)[:, L0]
                s0 = slot[L0]
                others = [j for j in range(K_REG) if j != J]
                for i, sb in enumerate(others):
                    use = (sb < J) if ladder else bool(op.flags & (1 << i))
                    if use:
                        q = complex(op.m[2 * i], op.m[2 * i + 1])
                        f = np.where(((s0 >> sb) & 1)[None, :] == 1, f * q, f)
                if kind == OP_PHASE:
                    a[:, L1] = y * f
                    if op.flags & FLAG_C0:
                        a[:, L0] = x * complex(op.m[8], op.m[9])
                else:                                          # phase_h_apply / LadderStep: (x, y) -> (x + f y, x - f y)
                    a[:, L0], a[:, L1] = x + f * y, x - f * y
                continue
            if kind == OP_H_UNNORM:
                a[:, L0], a[:, L1] = x + y, x - y
                continue
            # G1 ops: controls among the register bits (cslot) and among thread / outer bits (cmask)
            ok = ((slot[L0] & op.cslot) == op.cslot)[None, :] & ((v[:, L0] & np.int64(op.cmask)) == np.int64(op.cmask))
            m = [complex(op.m[2 * i], op.m[2 * i + 1]) for i in range(4)]
            if kind == OP_G1_HADAMARD:
                c = op.m[0]
                nx, ny = (x + y) * c, (x - y) * c
            elif kind == OP_G1_ANTIDIAG:
                nx, ny = m[1] * y, m[2] * x
            elif kind == OP_G1_SWAPX:
                nx, ny = y, x
            elif kind == OP_G1_DIAG:
                nx, ny = m[0] * x, m[3] * y
            else:
                nx, ny = m[0] * x + m[1] * y, m[2] * x + m[3] * y
            a[:, L0], a[:, L1] = np.where(ok, nx, x), np.where(ok, ny, y)
    out = np.empty_like(psi)
    out[dst] = a * P.scale
    return out


def run_plan(sweeps, perm, psi):
    """psi in canonical order -> result in canonical order (perm undone at the end, as canonicalize() does)"""
    n = sweeps[0][0].n if sweeps else len(perm)
    cur = np.array(psi, dtype=np.complex128)
    for P, ptabs in sweeps:
        cur = run_sweep(P, ptabs, cur)
    idx = np.arange(1 << n, dtype=np.int64)
    phys = _deposit(idx, perm)                                # logical index -> physical index
    return cur[phys]


def check_tables(P):
    """the address tables the kernels read, against the logical layout they are derived from"""
    n, T, TB, n_outer = P.n, P.T, P.TB, P.n_outer
    tsrc, tdst = [P.tsrc[i] for i in range(T)], [P.tdst[i] for i in range(T)]
    osrc, odst = [P.osrc[i] for i in range(n_outer)], [P.odst[i] for i in range(n_outer)]
    assert sorted(tsrc + osrc) == list(range(n)) and sorted(tdst + odst) == list(range(n))
    assert P.tile_mask_src == sum(1 << p for p in tsrc)
    o = np.arange(1 << n_outer, dtype=np.int64)
    for tab, pos in ((P.o_src, osrc), (P.o_dst, odst)):
        got = np.zeros_like(o)
        for c in range(K_CHUNKS):
            if c * K_CHUNK_BITS < max(n_outer, 1):
                got |= np.array(tab[c][:], dtype=np.uint64).astype(np.int64)[(o >> (K_CHUNK_BITS * c)) & 63]
        assert np.array_equal(got, _deposit(o, pos))
    # the destination-ordered walk visits every tile once, in ascending destination address
    ws = np.zeros_like(o)
    wd = np.zeros_like(o)
    for c in range(K_CHUNKS):
        if c * K_CHUNK_BITS < max(n_outer, 1):
            ws |= np.array(P.w_src[c][:], dtype=np.uint64).astype(np.int64)[(o >> (K_CHUNK_BITS * c)) & 63]
            wd |= np.array(P.w_dst[c][:], dtype=np.uint64).astype(np.int64)[(o >> (K_CHUNK_BITS * c)) & 63]
    assert np.all(np.diff(wd) > 0) or n_outer == 0
    pairs = sorted(zip(_deposit(o, osrc).tolist(), _deposit(o, odst).tolist()))
    assert sorted(zip(ws.tolist(), wd.tolist())) == pairs
    # load: element e = tid | i << TB of the tile lives at source offset dep(tid) | ld_hi[i], swizzled slot ld_sw_hi[i]
    tid = np.arange(1 << TB, dtype=np.uint64)
    lsrc = _deposit(np.arange(1 << T, dtype=np.int64), tsrc)
    dep = _run_bits(tid, P.ld_runs, P.ld_nruns).astype(np.int64)
    for i in range(K_SLOTS):
        e = tid.astype(np.int64) | (i << TB)
        assert np.array_equal(dep | np.int64(P.ld_hi[i]), lsrc[e])
        assert P.ld_sw_hi[i] == tile_swizzle(i << TB) * 16
    # store: element f = tid | i << TB in ascending destination order: tile index via st_lruns / st_l_hi,
    # destination offset via st_runs / st_off_hi; together a bijection of the tile onto its destination set
    ldst = _deposit(np.arange(1 << T, dtype=np.int64), tdst)
    doff = _run_bits(tid, P.st_runs, P.st_nruns).astype(np.int64)
    llo = _run_bits(tid, P.st_lruns, P.st_nruns).astype(np.int64)
    seen = np.zeros(1 << T, dtype=bool)
    for i in range(K_SLOTS):
        lidx = llo | np.int64(tile_swizzle(P.st_l_hi[i] >> 4))
        assert np.array_equal(doff | np.int64(P.st_off_hi[i]), ldst[lidx])
        seen[lidx] = True
    assert seen.all()
    # rounds: thread -> tile-local index (runs) and slot -> swizzled offset (sw_slot)
    for r in range(P.nrounds):
        R = P.rounds[r]
        reg_tb, thr_tb = [R.reg_tb[j] for j in range(K_REG)], [R.thr_tb[i] for i in range(TB)]
        assert sorted(reg_tb + thr_tb) == list(range(T))
        thrL = _run_bits(tid, R.runs, R.nruns).astype(np.int64)
        assert np.array_equal(thrL, _deposit(tid.astype(np.int64), thr_tb))
        for s in range(K_SLOTS):
            ls = sum(((s >> j) & 1) << reg_tb[j] for j in range(K_REG))
            assert R.sw_slot[s] == tile_swizzle(ls) * 16


def tma_swizzle(m):
    return m ^ ((m >> 3) & 7)


def check_tma_tables(P):
    """a program in TMA layout (planner.cpp apply_tma_layout): the tensor description delivers every tile element to
    the shared-memory position the rewritten offset tables expect, and every LDS/STS.128 of a quarter warp is
    conflict-free under the 128-byte TMA swizzle"""
    assert P.tma_nreq > 0
    n, T, TB = P.n, P.T, P.TB
    tsrc, tdst = [P.tsrc[i] for i in range(T)], [P.tdst[i] for i in range(T)]
    pi = [P.tma_pi[i] for i in range(T)]
    assert sorted(pi) == list(range(T)) and pi[:3] == [0, 1, 2]
    l = np.arange(1 << T, dtype=np.int64)
    m_of_l = _deposit(l, pi)
    l_of_m = np.empty_like(l)
    l_of_m[m_of_l] = l
    lsrc = _deposit(l, tsrc)
    # what the TMA unit does with the 5-d box: element (i3, i2, i1, chunk) of request q
    box = [P.tma_box[i] for i in range(5)]
    gs = [P.tma_gstride[i] for i in range(4)]
    assert box[0] == 16 and box[4] == 1 and gs[3] == 128 and P.tma_gdim[4] == 1 << (n - 3)
    assert all(1 <= b <= 256 for b in box) and all(g % 16 == 0 and 0 < g < (1 << 40) for g in gs)
    assert P.tma_nreq * P.tma_req_bytes == 16 << T and P.tma_req_bytes == 128 * box[1] * box[2] * box[3]
    src_at_m = np.empty(1 << T, dtype=np.int64)
    for q in range(P.tma_nreq):
        i1, i2, i3, c = np.meshgrid(np.arange(box[1]), np.arange(box[2]), np.arange(box[3]), np.arange(8), indexing="ij")
        off = (i1 * gs[0] + i2 * gs[1] + i3 * gs[2]) // 16 + P.tma_req_line[q] * 8 + c
        mlin = q * (P.tma_req_bytes // 16) + ((i3 * box[2] + i2) * box[1] + i1) * 8 + c
        src_at_m[mlin.ravel()] = off.ravel()
    assert np.array_equal(src_at_m[m_of_l], lsrc)           # the element at smem index m(l) is tile element l
    tid = np.arange(1 << TB, dtype=np.uint64)

    def conflict_free(addr16):                               # addr16: (threads,) 16-byte unit addresses of one LDS/STS.128
        banks = (addr16 & 7).reshape(-1, 8)                  # a quarter warp = one 128-byte wavefront
        return all(len(set(row.tolist())) == 8 for row in banks)

    for r in range(P.nrounds):
        R = P.rounds[r]
        reg_tb, thr_tb = [R.reg_tb[j] for j in range(K_REG)], [R.thr_tb[i] for i in range(TB)]
        thrM = _run_bits(tid, R.runs, R.nruns).astype(np.int64)
        assert np.array_equal(thrM, _deposit(tid.astype(np.int64), [pi[tb] for tb in thr_tb]))
        for s in range(K_SLOTS):
            ms = sum(((s >> j) & 1) << pi[reg_tb[j]] for j in range(K_REG))
            assert R.sw_slot[s] == tma_swizzle(ms) * 16
            assert conflict_free((tma_swizzle(thrM) * 16 ^ R.sw_slot[s]) >> 4)
    ldst = _deposit(l, tdst)
    doff = _run_bits(tid, P.st_runs, P.st_nruns).astype(np.int64)
    mlo = _run_bits(tid, P.st_lruns, P.st_nruns).astype(np.int64)
    seen = np.zeros(1 << T, dtype=bool)
    for i in range(K_SLOTS):
        midx = mlo | np.int64(tma_swizzle(P.st_l_hi[i] >> 4))
        assert np.array_equal(doff | np.int64(P.st_off_hi[i]), ldst[l_of_m[midx]])
        seen[midx] = True
        if not P.direct_store:
            assert conflict_free((tma_swizzle(mlo) * 16 ^ P.st_l_hi[i]) >> 4)
    assert seen.all()
