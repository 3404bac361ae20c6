"""A numpy stand-in for the per-rank local state, with the interface ShardedState expects
from q1tsim_b200.sharded.EngineLocal.  TEST INFRASTRUCTURE: it lets the world_size>1 host
logic (qubit remap planning, matrix block selection, rank-ordered canonical chaining,
draw ownership) run under gloo on CPU; the product path always uses the CUDA engine."""
import numpy as np

from oracle import oracle as O
from tests import np_ref

ZERO_COLUMN = 0xFFFFFFFFFFFFFFFF


def canonical_leaf_totals(amps, mask_bit=None):
    """DESIGN.md 4.2 leaf geometry, vectorised: lanes accumulate sequentially, xor butterfly"""
    n = amps.size
    leaf = min(n, 1024)
    p = amps.real * amps.real + amps.imag * amps.imag
    if mask_bit is not None:
        idx = np.arange(n)
        p = np.where((idx >> mask_bit) & 1, 0.0, p)
    p = p.reshape(n // leaf, leaf)
    if leaf >= 32:
        acc = np.cumsum(p.reshape(-1, leaf // 32, 32), axis=1)[:, -1, :]
    else:
        acc = np.zeros((p.shape[0], 32))
        acc[:, :leaf] = p
    lanes = np.arange(32)
    for off in (16, 8, 4, 2, 1):
        acc = acc + acc[:, lanes ^ off]
    return acc[:, 0].copy()


class NumpyLocal:
    def __init__(self, n_local, shots, device, empty):
        self.n_local, self.shots = n_local, shots
        col = np.zeros(1 << n_local, dtype=np.complex128)
        if not empty:
            col[0] = 1.0
        self.cols = [col]
        self._counts = [shots]

    def apply_gate(self, mat, qubits):
        self.cols = [np_ref.apply_gate(c, mat, list(qubits), self.n_local) for c in self.cols]

    def apply_conditional_gate(self, control, mat, qubits):
        ranges = O.collect_conditional_ranges(self._counts, list(control))
        nc, ncnt = [], []
        for icol, ln, ap in ranges:
            c = self.cols[icol].copy()
            if ap:
                c = np_ref.apply_gate(c, mat, list(qubits), self.n_local)
            nc.append(c)
            ncnt.append(ln)
        self.cols, self._counts = nc, ncnt

    @property
    def ncols(self):
        return len(self.cols)

    @property
    def counts(self):
        return list(self._counts)

    @property
    def nleaves(self):
        return (1 << self.n_local) // min(1 << self.n_local, 1024)

    def read_column(self, col):
        return self.cols[col].copy()

    def write_column(self, col, amps):
        self.cols[col] = np.array(amps, dtype=np.complex128)

    def leaf_totals(self, qubit):
        bit = None if qubit is None else self.n_local - 1 - qubit
        return np.stack([canonical_leaf_totals(c, bit) for c in self.cols])

    def resolve_draws(self, col, P, base, chosen):
        a = self.cols[col]
        leaf = min(a.size, 1024)
        out = []
        for ch in chosen:
            lo = int(np.searchsorted(P[:-1], ch, side="right"))
            run = base if lo == 0 else P[lo - 1]
            seg = a[lo * leaf:(lo + 1) * leaf]
            p = seg.real * seg.real + seg.imag * seg.imag
            found, last_nz = None, None
            for e in range(leaf):
                if p[e] > 0:
                    last_nz = e
                run = run + p[e]
                if ch < run:
                    found = e
                    break
            if found is None:
                found = leaf - 1 if last_nz is None else last_nz
            out.append(lo * leaf + found)
        return np.array(out, dtype=np.uint64)

    def _split(self, make0, make1, n0s):
        nc, ncnt = [], []
        for c, (col, n0, cnt) in enumerate(zip(self.cols, n0s, self._counts)):
            if n0 == cnt:
                nc.append(make0(c, col)); ncnt.append(cnt)
            elif n0 == 0:
                nc.append(make1(c, col)); ncnt.append(cnt)
            else:
                nc += [make0(c, col), make1(c, col)]
                ncnt += [n0, cnt - n0]
        self.cols, self._counts = nc, ncnt

    def collapse_columns(self, qubit, w0, n0):
        bit = self.n_local - 1 - qubit
        one = ((np.arange(1 << self.n_local) >> bit) & 1).astype(bool)
        self._split(lambda c, col: np.where(one, 0, col * (1.0 / np.sqrt(w0[c]))),
                    lambda c, col: np.where(one, col * (1.0 / np.sqrt(1.0 - w0[c])), 0), n0)

    def scale_split_columns(self, f0, f1, n0):
        self._split(lambda c, col: col * f0[c], lambda c, col: col * f1[c], n0)

    def replace_columns(self, idx, counts):
        self.cols = []
        for v in idx:
            col = np.zeros(1 << self.n_local, dtype=np.complex128)
            if int(v) != ZERO_COLUMN:
                col[int(v)] = 1.0
            self.cols.append(col)
        self._counts = [int(c) for c in counts]

    def column_tensor(self, col):
        import torch
        self.cols[col] = np.ascontiguousarray(self.cols[col])
        return torch.from_numpy(self.cols[col].view(np.float64))

    def draws(self, rng, total, n):
        # rand 0.7 Uniform(0, total): same restatement as the engine's host code
        max_rand = 1.0 - 2.220446049250313e-16
        scale = total
        while scale * max_rand + 0.0 >= total:
            scale = np.nextafter(scale, -np.inf)
        out = np.empty(n)
        for i in range(n):
            bits = (rng.next_u64() >> 12) | 0x3FF0000000000000
            out[i] = (np.frombuffer(np.uint64(bits).tobytes(), dtype=np.float64)[0] - 1.0) * scale + 0.0
        return out

    def binomial(self, rng, n, p):
        return rng.binomial(n, p)

    # peer-group emulation (the CUDA engine swaps in place over NVLink peer memory; here every rank gathers all
    # shards and picks what the multi-bit remap brings it): lets the gloo tests cover ShardedState._exchange_multi
    def group_setup(self, dist, group, rank, P):
        self._dist, self._group, self._rank, self._P = dist, group, rank, P

    def group_remap(self, rank_bits, local_qubits):
        import torch
        n, r = self.n_local, self._rank
        l = np.arange(1 << n, dtype=np.int64)
        src_rank = np.full_like(l, r)
        src_idx = l.copy()
        for gb, q in zip(rank_bits, local_qubits):
            p = n - 1 - q
            lb = (l >> p) & 1                                     # the element now at (r, l) came from rank bit := l_p, index bit := r_gb
            src_rank = (src_rank & ~(1 << gb)) | (lb << gb)
            src_idx = (src_idx & ~(1 << p)) | (((r >> gb) & 1) << p)
        out = []
        for c in self.cols:
            t = torch.from_numpy(np.ascontiguousarray(c).view(np.float64).copy())
            parts = [torch.empty_like(t) for _ in range(self._P)]
            self._dist.all_gather(parts, t, group=self._group)
            allc = np.stack([p_.numpy().view(np.complex128) for p_ in parts])
            out.append(allc[src_rank, src_idx])
        self.cols = out

    def pack_gates(self, gates):
        return [(np.array(m), list(b)) for m, b in gates]

    def apply_packed(self, packed):
        self.replayed = getattr(self, "replayed", 0) + len(packed)
        for m, b in packed:
            self.apply_gate(m, b)

    def scale(self, s):
        self.cols = [c * complex(s) for c in self.cols]

    def reset_all(self):
        col = np.zeros(1 << self.n_local, dtype=np.complex128)
        col[0] = 1.0
        self.cols, self._counts = [col], [self.shots]

    def set_product_state(self, coefs):
        self.cols = [O.OracleState.from_qubit_coefs(list(coefs), 1).column(0)]
        self._counts = [self.shots]


def factory(n_local, shots, device, empty):
    return NumpyLocalNoGroup(n_local, shots, device, empty)


class NumpyLocalNoGroup(NumpyLocal):
    group_setup = None


def factory_group(n_local, shots, device, empty):
    return NumpyLocal(n_local, shots, device, empty)
