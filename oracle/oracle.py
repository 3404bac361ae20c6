"""ctypes binding of the CPU oracle (oracle/libq1t_oracle.so) + a Python
restatement of the reference's circuit interpreter on top of it.

TEST INFRASTRUCTURE ONLY (see oracle/q1t_oracle.h).  Importable only from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs.  The product package q1tsim_b200 never imports this module.

`OracleCircuit` follows `Circuit` of the reference: builder methods
circuit.rs:161-554, interpreter `do_execute_with` circuit.rs:643-762,
histograms circuit.rs:773-841.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libq1t_oracle.so")

ORC_ERRORS = {
    -1: "InvalidNrBits", -2: "InvalidQBit", -3: "NotEnoughSpace",
    -4: "InvalidNrMeasurementBits", -5: "InvalidNrControlBits", -6: "RngExhausted", -7: "OutOfHostMemory",
}


class OracleError(Exception):
    def __init__(self, code):
        super().__init__(ORC_ERRORS.get(code, str(code)))
        self.code = code
        self.kind = ORC_ERRORS.get(code, str(code))


def build(force=False):
    # make decides: the library is rebuilt whenever q1t_oracle.c / .h are newer
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


class _Rng(C.Structure):
    _fields_ = [("kind", C.c_int), ("s", C.c_uint64), ("arr", C.POINTER(C.c_uint64)),
                ("n", C.c_size_t), ("pos", C.c_size_t), ("exhausted", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    sz, u64p, dp, szp, u8p = C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_size_t), C.POINTER(C.c_uint8)
    vp, rp = C.c_void_p, C.POINTER(_Rng)
    L.orc_rng_seed.argtypes = [rp, C.c_uint64]
    L.orc_rng_array.argtypes = [rp, u64p, sz]
    L.orc_rng_next.argtypes = [rp]; L.orc_rng_next.restype = C.c_uint64
    L.orc_binomial.argtypes = [rp, C.c_uint64, C.c_double]; L.orc_binomial.restype = C.c_uint64
    L.orc_state_new.argtypes = [sz, sz]; L.orc_state_new.restype = vp
    L.orc_state_from_qubit_coefs.argtypes = [dp, sz, sz]; L.orc_state_from_qubit_coefs.restype = vp
    L.orc_state_free.argtypes = [vp]
    L.orc_state_ncols.argtypes = [vp]; L.orc_state_ncols.restype = sz
    L.orc_state_counts.argtypes = [vp, szp]
    L.orc_state_read_column.argtypes = [vp, sz, dp]
    L.orc_state_write_column.argtypes = [vp, sz, dp]
    L.orc_set_threads.argtypes = [C.c_int]
    L.orc_apply_gate.argtypes = [vp, dp, szp, sz, C.c_int]
    L.orc_apply_unary_gate_all.argtypes = [vp, dp, C.c_int]
    L.orc_apply_conditional_gate.argtypes = [vp, u8p, sz, dp, szp, sz, C.c_int]
    L.orc_marginal0.argtypes = [vp, sz, C.c_int, dp]
    L.orc_measure_into.argtypes = [vp, sz, sz, u64p, sz, rp, C.c_int]
    L.orc_peek_into.argtypes = [vp, sz, sz, u64p, sz, rp, C.c_int]
    L.orc_measure_all_into.argtypes = [vp, szp, sz, u64p, sz, C.c_int, rp, C.c_int]
    L.orc_reset.argtypes = [vp, sz, rp, C.c_int, C.c_int]
    L.orc_reset_all.argtypes = [vp]
    L.orc_column_totals.argtypes = [vp, C.c_int, dp]
    L.orc_reverse_bits.argtypes = [C.c_uint64, sz]; L.orc_reverse_bits.restype = C.c_uint64
    L.orc_shuffle_bits.argtypes = [C.c_uint64, szp, sz]; L.orc_shuffle_bits.restype = C.c_uint64
    L.orc_collect_conditional_ranges.argtypes = [szp, sz, u8p, szp, szp, u8p]
    L.orc_collect_conditional_ranges.restype = sz
    L.orc_bit_permutation.argtypes = [sz, szp, sz, szp]
    L.orc_gate_matrix.argtypes = [C.c_char_p, dp, sz, dp]
    _lib = L
    return L


def _szarr(xs):
    xs = list(xs)
    return (C.c_size_t * max(len(xs), 1))(*xs)


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _mat_arg(mat):
    m = np.ascontiguousarray(np.asarray(mat, dtype=np.complex128))
    return m, _dptr(m.view(np.float64))


class Rng:
    """SplitMix64(seed) or an injected array of raw u64 words."""

    def __init__(self, seed=None, words=None):
        self._r = _Rng()
        self._keep = None
        if words is not None:
            self._keep = np.ascontiguousarray(np.asarray(words, dtype=np.uint64))
            lib().orc_rng_array(C.byref(self._r), self._keep.ctypes.data_as(C.POINTER(C.c_uint64)), self._keep.size)
        else:
            lib().orc_rng_seed(C.byref(self._r), C.c_uint64(0 if seed is None else seed))

    def next_u64(self):
        return int(lib().orc_rng_next(C.byref(self._r)))

    def binomial(self, n, p):
        return int(lib().orc_binomial(C.byref(self._r), n, p))

    @property
    def consumed(self):
        return int(self._r.pos)

    @property
    def ref(self):
        return C.byref(self._r)


def splitmix64_words(seed, n):
    r = Rng(seed=seed)
    return np.array([r.next_u64() for _ in range(n)], dtype=np.uint64)


def gate_matrix(name, params=()):
    """`matrix()` of a built-in gate by name (composite.rs:287-445 table)."""
    params = np.asarray(list(params), dtype=np.float64)
    out = np.zeros(2 * 64, dtype=np.float64)
    nb = lib().orc_gate_matrix(name.encode(), _dptr(params) if params.size else None, params.size, _dptr(out))
    if nb == -1:
        raise KeyError("unknown gate %r" % name)
    if nb == -2:
        raise ValueError("wrong number of parameters for %r" % name)
    g = 1 << nb
    return out[:2 * g * g].view(np.complex128).reshape(g, g).copy()


def reverse_bits(idx, n):
    return int(lib().orc_reverse_bits(idx, n))


def shuffle_bits(idx, bits):
    return int(lib().orc_shuffle_bits(idx, _szarr(bits), len(bits)))


def collect_conditional_ranges(counts, control):
    n = len(control) + len(counts) + 1
    ic, ln, ap = (C.c_size_t * n)(), (C.c_size_t * n)(), (C.c_uint8 * n)()
    ctl = (C.c_uint8 * max(len(control), 1))(*[1 if c else 0 for c in control])
    nr = lib().orc_collect_conditional_ranges(_szarr(counts), len(counts), ctl, ic, ln, ap)
    return [(int(ic[i]), int(ln[i]), bool(ap[i])) for i in range(nr)]


def bit_permutation(nr_bits, bits):
    out = (C.c_size_t * (1 << nr_bits))()
    rc = lib().orc_bit_permutation(nr_bits, _szarr(bits), len(bits), out)
    if rc != 0:
        raise ValueError("invalid permutation")
    return list(out)


class OracleState:
    """`VectorState` of the reference (vectorstate.rs:25-415).
    mode: 0 faithful loop structure, 1 fast.  order: 0 reference summation
    order, 1 canonical blocked order (shared with the GPU engine)."""

    def __init__(self, nr_bits, nr_shots, mode=1, order=1, _ptr=None):
        self.nr_bits, self.nr_shots, self.mode, self.order = nr_bits, nr_shots, mode, order
        self._p = _ptr if _ptr is not None else lib().orc_state_new(nr_bits, nr_shots)

    @classmethod
    def from_qubit_coefs(cls, coefs, nr_shots, mode=1, order=1):
        c = np.ascontiguousarray(np.asarray(coefs, dtype=np.complex128))
        assert c.size % 2 == 0
        p = lib().orc_state_from_qubit_coefs(_dptr(c.view(np.float64)), c.size // 2, nr_shots)
        return cls(c.size // 2, nr_shots, mode, order, _ptr=p)

    def __del__(self):
        if getattr(self, "_p", None):
            lib().orc_state_free(self._p)
            self._p = None

    def _chk(self, rc):
        if rc != 0:
            raise OracleError(rc)

    @property
    def ncols(self):
        return int(lib().orc_state_ncols(self._p))

    @property
    def counts(self):
        out = (C.c_size_t * self.ncols)()
        lib().orc_state_counts(self._p, out)
        return list(out)

    def column(self, col):
        out = np.empty(2 << self.nr_bits, dtype=np.float64)
        lib().orc_state_read_column(self._p, col, _dptr(out))
        return out.view(np.complex128)

    def set_column(self, col, amps):
        a = np.ascontiguousarray(np.asarray(amps, dtype=np.complex128))
        assert a.size == 1 << self.nr_bits
        lib().orc_state_write_column(self._p, col, _dptr(a.view(np.float64)))

    def states(self):
        """(2^n, C) matrix like the reference's `states` field."""
        return np.stack([self.column(c) for c in range(self.ncols)], axis=1)

    def apply_gate(self, mat, bits):
        m, mp = _mat_arg(mat)
        if m.shape[0] != 1 << len(bits):
            raise OracleError(-1)
        self._chk(lib().orc_apply_gate(self._p, mp, _szarr(bits), len(bits), self.mode))

    def apply_unary_gate_all(self, mat):
        m, mp = _mat_arg(mat)
        self._chk(lib().orc_apply_unary_gate_all(self._p, mp, self.mode))

    def apply_conditional_gate(self, control, mat, bits):
        m, mp = _mat_arg(mat)
        ctl = np.ascontiguousarray(np.asarray(control, dtype=np.uint8))
        if ctl.size != self.nr_shots:
            raise OracleError(-5)
        if m.shape[0] != 1 << len(bits):
            raise OracleError(-1)
        self._chk(lib().orc_apply_conditional_gate(self._p, ctl.ctypes.data_as(C.POINTER(C.c_uint8)), ctl.size,
                                                   mp, _szarr(bits), len(bits), self.mode))

    def marginal0(self, qbit, order=None):
        out = np.zeros(self.ncols, dtype=np.float64)
        self._chk(lib().orc_marginal0(self._p, qbit, self.order if order is None else order, _dptr(out)))
        return out

    def column_totals(self, order=None):
        out = np.zeros(self.ncols, dtype=np.float64)
        lib().orc_column_totals(self._p, self.order if order is None else order, _dptr(out))
        return out

    @staticmethod
    def _res(res):
        assert res.dtype == np.uint64 and res.flags.c_contiguous
        return res.ctypes.data_as(C.POINTER(C.c_uint64))

    def measure_into(self, qbit, cbit, res, rng):
        self._chk(lib().orc_measure_into(self._p, qbit, cbit, self._res(res), res.size, rng.ref, self.order))

    def measure(self, qbit, rng):
        res = np.zeros(self.nr_shots, dtype=np.uint64)
        self.measure_into(qbit, 0, res, rng)
        return res

    def peek_into(self, qbit, cbit, res, rng):
        self._chk(lib().orc_peek_into(self._p, qbit, cbit, self._res(res), res.size, rng.ref, self.order))

    def measure_all_into(self, cbits, res, rng, collapse=True):
        self._chk(lib().orc_measure_all_into(self._p, _szarr(cbits), len(cbits), self._res(res), res.size,
                                             1 if collapse else 0, rng.ref, self.order))

    def measure_all(self, rng):
        res = np.zeros(self.nr_shots, dtype=np.uint64)
        self.measure_all_into(list(range(self.nr_bits)), res, rng)
        return res

    def peek_all_into(self, cbits, res, rng):
        self.measure_all_into(cbits, res, rng, collapse=False)

    def reset(self, bit, rng):
        self._chk(lib().orc_reset(self._p, bit, rng.ref, self.order, self.mode))

    def reset_all(self):
        lib().orc_reset_all(self._p)


H_MAT = None


def _basis_mats():
    global H_MAT
    if H_MAT is None:
        H_MAT = {k: gate_matrix(k) for k in ("h", "s", "sdg")}
    return H_MAT


class OracleCircuit:
    """`Circuit` of the reference, vector backend forced (circuit.rs)."""

    def __init__(self, nr_qbits, nr_cbits, mode=1, order=1):
        self.nr_qbits, self.nr_cbits = nr_qbits, nr_cbits
        self.mode, self.order = mode, order
        self.ops = []
        self.q_state = None
        self.c_state = None

    # builder (circuit.rs:161-554); matrices are evaluated at execute time
    def add_gate(self, name, bits, params=()):
        for b in bits:
            if b >= self.nr_qbits:
                raise OracleError(-2)
        self.ops.append(("gate", name, tuple(params), list(bits)))

    def add_matrix_gate(self, mat, bits):
        self.ops.append(("gate", np.asarray(mat, dtype=np.complex128), (), list(bits)))

    def add_conditional_gate(self, control, target, name, bits, params=()):
        self.ops.append(("cond", list(control), int(target), name, tuple(params), list(bits)))

    def measure_basis(self, qbit, cbit, basis="Z"):
        self.ops.append(("measure", qbit, cbit, basis.upper()))

    def measure(self, qbit, cbit):
        self.measure_basis(qbit, cbit, "Z")

    def measure_all_basis(self, cbits, basis="Z"):
        self.ops.append(("measure_all", list(cbits), basis.upper()))

    def measure_all(self, cbits):
        self.measure_all_basis(cbits, "Z")

    def peek_basis(self, qbit, cbit, basis="Z"):
        self.ops.append(("peek", qbit, cbit, basis.upper()))

    def peek_all_basis(self, cbits, basis="Z"):
        self.ops.append(("peek_all", list(cbits), basis.upper()))

    def reset(self, qbit):
        self.ops.append(("reset", qbit))

    def reset_all(self):
        self.ops.append(("reset_all",))

    def barrier(self, qbits):
        self.ops.append(("barrier", list(qbits)))

    # execution (circuit.rs:562-641)
    def execute(self, nr_shots, rng, q_state=None):
        self.q_state = q_state if q_state is not None else OracleState(self.nr_qbits, nr_shots, self.mode, self.order)
        self.c_state = np.zeros(nr_shots, dtype=np.uint64)
        self.reexecute(rng)

    @staticmethod
    def _mat(name, params):
        if isinstance(name, np.ndarray):
            return name
        params = [p() if callable(p) else p for p in params]
        return gate_matrix(name, params)

    def reexecute(self, rng):
        if self.c_state is None or self.q_state is None:
            raise RuntimeError("The circuit has not been executed yet")
        q, cs, bm = self.q_state, self.c_state, _basis_mats()
        for op in self.ops:                       # circuit.rs:643-762
            kind = op[0]
            if kind == "gate":
                q.apply_gate(self._mat(op[1], op[2]), op[3])
            elif kind == "cond":
                control, target, name, params, bits = op[1:]
                word = np.zeros(cs.size, dtype=np.uint64)
                for idst, isrc in enumerate(control):           # first control index = LSB
                    word |= ((cs >> np.uint64(isrc)) & np.uint64(1)) << np.uint64(idst)
                q.apply_conditional_gate(word == np.uint64(target), self._mat(name, params), bits)
            elif kind in ("measure", "peek"):
                qbit, cbit, basis = op[1:]
                f = q.measure_into if kind == "measure" else q.peek_into
                if basis == "X":
                    q.apply_gate(bm["h"], [qbit]); f(qbit, cbit, cs, rng); q.apply_gate(bm["h"], [qbit])
                elif basis == "Y":
                    q.apply_gate(bm["sdg"], [qbit]); q.apply_gate(bm["h"], [qbit])
                    f(qbit, cbit, cs, rng)
                    q.apply_gate(bm["h"], [qbit]); q.apply_gate(bm["s"], [qbit])
                else:
                    f(qbit, cbit, cs, rng)
            elif kind in ("measure_all", "peek_all"):
                cbits, basis = op[1:]
                f = q.measure_all_into if kind == "measure_all" else q.peek_all_into
                if basis == "X":
                    q.apply_unary_gate_all(bm["h"]); f(cbits, cs, rng); q.apply_unary_gate_all(bm["h"])
                elif basis == "Y":
                    q.apply_unary_gate_all(bm["sdg"]); q.apply_unary_gate_all(bm["h"])
                    f(cbits, cs, rng)
                    q.apply_unary_gate_all(bm["h"]); q.apply_unary_gate_all(bm["s"])
                else:
                    f(cbits, cs, rng)
            elif kind == "reset":
                q.reset(op[1], rng)
            elif kind == "reset_all":
                q.reset_all()
            elif kind == "barrier":
                pass

    # histograms (circuit.rs:773-841)
    def histogram(self):
        keys, cnt = np.unique(self.c_state, return_counts=True)
        return {int(k): int(c) for k, c in zip(keys, cnt)}

    def histogram_vec(self):
        out = [0] * (1 << self.nr_cbits)
        for k in self.c_state:
            out[int(k)] += 1
        return out

    def histogram_string(self):
        return {format(k, "0%db" % self.nr_cbits): c for k, c in self.histogram().items()}
