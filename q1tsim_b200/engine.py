"""Host-side Python mirror of the reference's `VectorState` / `QuState` interface
(src/vectorstate.rs, src/qustate.rs) over the C ABI of include/q1t_engine.h.

This is a thin ctypes veneer used by the tests, the benchmark and Python
drivers: every method maps 1:1 to one C-ABI entry point, names and argument
meaning follow the reference (apply_gate, apply_conditional_gate, measure_into,
peek_into, measure_all_into, peek_all_into, reset, reset_all).  There is no CPU
fallback: without the CUDA library or without a device every call fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("Q1T_LIB") or os.path.join(_HERE, "lib", "libq1tsim.so")   # Q1T_LIB: A/B builds

ERROR_KINDS = {
    -1: "InvalidNrBits", -2: "InvalidQBit", -3: "NotEnoughSpace", -4: "InvalidNrMeasurementBits",
    -5: "InvalidNrControlBits", -6: "RngExhausted", -7: "CudaError", -8: "InvalidArgument", -9: "Unsupported",
}


class EngineError(Exception):
    """Mirrors `q1tsim::error::Error` (error.rs:134-255): `.kind` is the variant,
    the message is the reference's Display text."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code
        self.kind = ERROR_KINDS.get(code, str(code))


class _RngHandle(C.Structure):
    _fields_ = [("next_u64", C.c_void_p), ("ctx", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("gates_queued", C.c_uint64), ("sweeps", C.c_uint64), ("sweep_column_passes", C.c_uint64),
                ("read_passes", C.c_uint64), ("kernel_launches", C.c_uint64), ("permute_sweeps", C.c_uint64),
                ("fallback_sweeps", C.c_uint64), ("fused_relabels", C.c_uint64), ("sweep_bytes", C.c_uint64), ("sweep_ms", C.c_double), ("read_ms", C.c_double),
                ("peer_swap_ms", C.c_double), ("peer_swap_bytes", C.c_uint64), ("plan_cache_hits", C.c_uint64),
                ("tma_sweeps", C.c_uint64), ("graph_captures", C.c_uint64), ("graph_replays", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("fused_remaps", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    """Load libq1tsim.so.  Raises if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("q1tsim_b200: %s is missing -- run `python -m q1tsim_b200.build` "
                           "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    sz, dp, szp, u8p, u64p, vp = C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_size_t), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.c_void_p
    R = _RngHandle
    sig = {
        "q1t_state_new": (C.c_int, [sz, sz, C.c_int, C.POINTER(vp)]),
        "q1t_state_from_qubit_coefs": (C.c_int, [dp, sz, sz, C.c_int, C.POINTER(vp)]),
        "q1t_state_free": (None, [vp]),
        "q1t_apply_gate": (C.c_int, [vp, dp, sz, szp, sz, C.c_char_p]),
        "q1t_apply_unary_gate_all": (C.c_int, [vp, dp, sz, C.c_char_p]),
        "q1t_apply_conditional_gate": (C.c_int, [vp, u8p, sz, dp, sz, szp, sz, C.c_char_p]),
        "q1t_measure": (C.c_int, [vp, sz, u64p, sz, R]),
        "q1t_measure_into": (C.c_int, [vp, sz, sz, u64p, sz, R]),
        "q1t_measure_all": (C.c_int, [vp, u64p, sz, R]),
        "q1t_measure_all_into": (C.c_int, [vp, szp, sz, u64p, sz, R]),
        "q1t_peek_into": (C.c_int, [vp, sz, sz, u64p, sz, R]),
        "q1t_peek_all_into": (C.c_int, [vp, szp, sz, u64p, sz, R]),
        "q1t_reset": (C.c_int, [vp, sz, R]),
        "q1t_reset_all": (C.c_int, [vp]),
        "q1t_nr_bits": (sz, [vp]), "q1t_nr_shots": (sz, [vp]), "q1t_nr_columns": (sz, [vp]),
        "q1t_counts": (C.c_int, [vp, szp]),
        "q1t_read_amplitudes": (C.c_int, [vp, sz, sz, sz, dp]),
        "q1t_write_amplitudes": (C.c_int, [vp, sz, sz, sz, dp]),
        "q1t_marginal0": (C.c_int, [vp, sz, dp]),
        "q1t_column_totals": (C.c_int, [vp, dp]),
        "q1t_flush": (C.c_int, [vp]),
        "q1t_last_error": (C.c_char_p, [vp]),
        "q1t_get_stats": (C.c_int, [vp, C.POINTER(Stats)]),
        "q1t_reset_stats": (C.c_int, [vp]),
        "q1t_set_timing": (C.c_int, [vp, C.c_int]),
        "q1t_set_option": (C.c_int, [vp, C.c_char_p, C.c_long]),
        "q1t_rng_splitmix64": (vp, [C.c_uint64]),
        "q1t_rng_from_words": (vp, [u64p, sz]),
        "q1t_rng_entropy": (vp, []),
        "q1t_rng_consumed": (sz, [vp]),
        "q1t_rng_free": (None, [vp]),
        "q1t_rng_handle": (R, [vp]),
        "q1t_binomial": (C.c_uint64, [R, C.c_uint64, C.c_double]),
        "q1t_gate_matrix": (C.c_int, [C.c_char_p, dp, sz, dp]),
        "q1t_plan_dry_run": (C.c_int, [sz, sz, dp, szp, szp, szp, C.c_long, u64p]),
        "q1t_plan_inplace_relabel": (C.c_int, [sz, C.c_long, C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), sz]),
        "q1t_composite_matrix": (C.c_int, [C.c_char_p, dp, sz, C.c_char_p, sz]),
        "q1t_eval_expression": (C.c_int, [C.c_char_p, dp, szp, C.c_char_p, sz]),
        "q1t_device_count": (C.c_int, []),
        "q1t_version": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


INNER_ABI_SYMBOLS = [
    "q1t_state_new", "q1t_state_from_qubit_coefs", "q1t_state_free", "q1t_apply_gate", "q1t_apply_unary_gate_all",
    "q1t_apply_conditional_gate", "q1t_measure", "q1t_measure_into", "q1t_measure_all", "q1t_measure_all_into",
    "q1t_peek_into", "q1t_peek_all_into", "q1t_reset", "q1t_reset_all", "q1t_nr_bits", "q1t_nr_shots", "q1t_nr_columns",
    "q1t_counts", "q1t_read_amplitudes", "q1t_write_amplitudes", "q1t_marginal0", "q1t_column_totals", "q1t_flush",
    "q1t_last_error", "q1t_get_stats", "q1t_reset_stats", "q1t_set_timing", "q1t_set_option", "q1t_rng_splitmix64",
    "q1t_rng_from_words", "q1t_rng_entropy", "q1t_rng_consumed", "q1t_rng_free", "q1t_rng_handle", "q1t_binomial",
    "q1t_gate_matrix", "q1t_plan_dry_run", "q1t_plan_inplace_relabel", "q1t_composite_matrix", "q1t_eval_expression", "q1t_device_count", "q1t_version",
]


def _szarr(xs):
    xs = list(xs)
    return (C.c_size_t * max(len(xs), 1))(*xs)


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Rng:
    """Caller-owned generator handed to the engine (the `R: Rng` of qustate.rs):
    SplitMix64(seed), an injected sequence of raw u64 words, or OS entropy."""

    def __init__(self, seed=None, words=None):
        L = lib()
        if words is not None:
            w = np.ascontiguousarray(np.asarray(words, dtype=np.uint64))
            self._p = L.q1t_rng_from_words(w.ctypes.data_as(C.POINTER(C.c_uint64)), w.size)
        elif seed is not None:
            self._p = L.q1t_rng_splitmix64(seed)
        else:
            self._p = L.q1t_rng_entropy()
        self.handle = L.q1t_rng_handle(self._p)

    def __del__(self):
        if getattr(self, "_p", None):
            lib().q1t_rng_free(self._p)
            self._p = None

    @property
    def consumed(self):
        return int(lib().q1t_rng_consumed(self._p))

    def binomial(self, n, p):
        return int(lib().q1t_binomial(self.handle, n, p))


def gate_matrix(name, params=()):
    """`matrix()` of a built-in gate (composite.rs:287-445 names)."""
    params = np.asarray([float(p) for p in params], dtype=np.float64)
    out = np.zeros(2 * 64, dtype=np.float64)
    nb = lib().q1t_gate_matrix(name.encode(), _dptr(params) if params.size else None, params.size, _dptr(out))
    if nb == -1:
        raise KeyError("Unknown gate %s" % name)
    if nb < 0:
        raise ValueError("Invalid number of arguments for gate %s" % name)
    g = 1 << nb
    return out[:2 * g * g].view(np.complex128).reshape(g, g).copy()


class ParseError(ValueError):
    """error.rs:71-127 (ParseError), message = its Display text"""


def composite_matrix(description):
    """`Composite::from_string(name, description).matrix()` (composite.rs:273-450, :480-485)."""
    out = np.zeros(2 << 20, dtype=np.float64)
    err = C.create_string_buffer(512)
    k = lib().q1t_composite_matrix(description.encode(), _dptr(out), out.size, err, 512)
    if k < 0:
        raise ParseError(err.value.decode())
    g = 1 << k
    return out[:2 * g * g].view(np.complex128).reshape(g, g).copy()


def eval_expression(text):
    """`Expression::parse(text)` + `eval()` (expression.rs:304-392): returns (value, unparsed rest)."""
    v = C.c_double()
    used = C.c_size_t()
    err = C.create_string_buffer(512)
    rc = lib().q1t_eval_expression(text.encode(), C.byref(v), C.byref(used), err, 512)
    if rc:
        raise ParseError(err.value.decode())
    return v.value, text.encode()[used.value:].decode()


def plan_dry_run(nr_bits, gates, tile_bits=12):
    """gates: list of (matrix, bits).  Returns dict(sweeps, rounds, ops, fallback, permute)."""
    mats = np.concatenate([np.ascontiguousarray(np.asarray(m, dtype=np.complex128)).reshape(-1).view(np.float64) for m, _ in gates]) \
        if gates else np.zeros(1)
    dims = _szarr([np.asarray(m).shape[0] for m, _ in gates])
    bits = _szarr([b for _, bs in gates for b in bs])
    nb = _szarr([len(bs) for _, bs in gates])
    out = (C.c_uint64 * 5)()
    rc = lib().q1t_plan_dry_run(nr_bits, len(gates), _dptr(mats), dims, bits, nb, tile_bits, out)
    if rc:
        raise EngineError(rc, "plan_dry_run failed")
    return dict(zip(("sweeps", "rounds", "ops", "fallback", "permute"), [int(v) for v in out]))


def plan_inplace_relabel(dstpos, tile_bits=12, coalesce_bits=3, max_passes=64):
    """The tile-closed passes the engine runs when it has to restore canonical order in place (no room for a
    second column buffer).  dstpos[p] = destination position of index bit p.  Returns [(tile, pass_dstpos), ...]."""
    n = len(dstpos)
    T = min(tile_bits, n)
    dp = (C.c_int * n)(*[int(d) for d in dstpos])
    tiles = (C.c_int * (max_passes * T))()
    outs = (C.c_int * (max_passes * n))()
    k = lib().q1t_plan_inplace_relabel(n, tile_bits, coalesce_bits, dp, tiles, outs, max_passes)
    if k < 0:
        raise EngineError(k, "plan_inplace_relabel failed")
    return [(list(tiles[i * T:(i + 1) * T]), list(outs[i * n:(i + 1) * n])) for i in range(k)]


class VectorState:
    """HBM-resident `VectorState` (vectorstate.rs:25-415)."""

    def __init__(self, nr_bits, nr_shots, device=0, _coefs=None):
        L = lib()
        p = C.c_void_p()
        if _coefs is None:
            rc = L.q1t_state_new(nr_bits, nr_shots, device, C.byref(p))
        else:
            rc = L.q1t_state_from_qubit_coefs(_dptr(_coefs), nr_bits, nr_shots, device, C.byref(p))
        if rc:
            raise EngineError(rc, L.q1t_last_error(None).decode())
        self._p = p
        self.nr_bits, self.nr_shots = nr_bits, nr_shots

    @classmethod
    def from_qubit_coefs(cls, coefs, nr_shots, device=0):
        c = np.ascontiguousarray(np.asarray(coefs, dtype=np.complex128))
        assert c.size % 2 == 0, "Length of coefficient array is not even"
        return cls(c.size // 2, nr_shots, device, _coefs=c.view(np.float64))

    def close(self):
        if getattr(self, "_p", None):
            lib().q1t_state_free(self._p)
            self._p = None

    __del__ = close

    def _chk(self, rc):
        if rc:
            raise EngineError(rc, lib().q1t_last_error(self._p).decode())

    # ---- QuState ----
    def apply_gate(self, mat, bits, desc="gate"):
        m = np.ascontiguousarray(np.asarray(mat, dtype=np.complex128))
        self._chk(lib().q1t_apply_gate(self._p, _dptr(m.view(np.float64)), m.shape[0], _szarr(bits), len(bits), desc.encode()))

    def apply_unary_gate_all(self, mat, desc="gate"):
        m = np.ascontiguousarray(np.asarray(mat, dtype=np.complex128))
        self._chk(lib().q1t_apply_unary_gate_all(self._p, _dptr(m.view(np.float64)), m.shape[0], desc.encode()))

    def apply_conditional_gate(self, control, mat, bits, desc="gate"):
        m = np.ascontiguousarray(np.asarray(mat, dtype=np.complex128))
        ctl = np.ascontiguousarray(np.asarray(control, dtype=np.uint8))
        self._chk(lib().q1t_apply_conditional_gate(self._p, ctl.ctypes.data_as(C.POINTER(C.c_uint8)), ctl.size,
                                                   _dptr(m.view(np.float64)), m.shape[0], _szarr(bits), len(bits), desc.encode()))

    @staticmethod
    def _res(res):
        assert res.dtype == np.uint64 and res.flags.c_contiguous
        return res.ctypes.data_as(C.POINTER(C.c_uint64))

    def measure(self, qbit, rng):
        res = np.zeros(self.nr_shots, dtype=np.uint64)
        self._chk(lib().q1t_measure(self._p, qbit, self._res(res), res.size, rng.handle))
        return res

    def measure_into(self, qbit, cbit, res, rng):
        self._chk(lib().q1t_measure_into(self._p, qbit, cbit, self._res(res), res.size, rng.handle))

    def measure_all(self, rng):
        res = np.zeros(self.nr_shots, dtype=np.uint64)
        self._chk(lib().q1t_measure_all(self._p, self._res(res), res.size, rng.handle))
        return res

    def measure_all_into(self, cbits, res, rng):
        self._chk(lib().q1t_measure_all_into(self._p, _szarr(cbits), len(cbits), self._res(res), res.size, rng.handle))

    def peek_into(self, qbit, cbit, res, rng):
        self._chk(lib().q1t_peek_into(self._p, qbit, cbit, self._res(res), res.size, rng.handle))

    def peek_all_into(self, cbits, res, rng):
        self._chk(lib().q1t_peek_all_into(self._p, _szarr(cbits), len(cbits), self._res(res), res.size, rng.handle))

    def reset(self, bit, rng):
        self._chk(lib().q1t_reset(self._p, bit, rng.handle))

    def reset_all(self):
        self._chk(lib().q1t_reset_all(self._p))

    # ---- accessors ----
    @property
    def ncols(self):
        return int(lib().q1t_nr_columns(self._p))

    @property
    def counts(self):
        out = (C.c_size_t * max(self.ncols, 1))()
        self._chk(lib().q1t_counts(self._p, out))
        return list(out)[:self.ncols]

    def column(self, col, offset=0, length=None):
        length = (1 << self.nr_bits) - offset if length is None else length
        out = np.empty(2 * length, dtype=np.float64)
        self._chk(lib().q1t_read_amplitudes(self._p, col, offset, length, _dptr(out)))
        return out.view(np.complex128)

    def set_column(self, col, amps, offset=0):
        a = np.ascontiguousarray(np.asarray(amps, dtype=np.complex128))
        self._chk(lib().q1t_write_amplitudes(self._p, col, offset, a.size, _dptr(a.view(np.float64))))

    def states(self):
        return np.stack([self.column(c) for c in range(self.ncols)], axis=1)

    def marginal0(self, qbit):
        out = np.zeros(max(self.ncols, 1), dtype=np.float64)
        self._chk(lib().q1t_marginal0(self._p, qbit, _dptr(out)))
        return out[:self.ncols]

    def column_totals(self):
        out = np.zeros(max(self.ncols, 1), dtype=np.float64)
        self._chk(lib().q1t_column_totals(self._p, _dptr(out)))
        return out[:self.ncols]

    def flush(self):
        self._chk(lib().q1t_flush(self._p))

    def stats(self):
        s = Stats()
        self._chk(lib().q1t_get_stats(self._p, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        self._chk(lib().q1t_reset_stats(self._p))

    def set_timing(self, on=True):
        self._chk(lib().q1t_set_timing(self._p, 1 if on else 0))

    def set_option(self, key, value):
        self._chk(lib().q1t_set_option(self._p, key.encode(), int(value)))


class ShardedProcessState:
    """A state sharded over the devices of THIS process (csrc/sharded.h, q1t_sharded_* of include/q1t_engine.h): the C++
    form of q1tsim_b200.sharded.ShardedState for gates, measure_all / peek_all and read-out.  `devices` may repeat a
    device (several shards on one GPU)."""

    def __init__(self, nr_bits, nr_shots, devices):
        L = lib()
        vp, sz = C.c_void_p, C.c_size_t
        L.q1t_sharded_new.restype = C.c_int
        L.q1t_sharded_new.argtypes = [sz, sz, sz, C.POINTER(C.c_int), C.POINTER(vp)]
        L.q1t_sharded_free.restype = None
        L.q1t_sharded_free.argtypes = [vp]
        L.q1t_sharded_apply_gate.restype = C.c_int
        L.q1t_sharded_apply_gate.argtypes = [vp, C.POINTER(C.c_double), sz, C.POINTER(sz), sz, C.c_char_p]
        L.q1t_sharded_set_initial_layout.restype = C.c_int
        L.q1t_sharded_set_initial_layout.argtypes = [vp, C.POINTER(C.c_int), sz]
        for f in (L.q1t_sharded_measure_all_into, L.q1t_sharded_peek_all_into):
            f.restype = C.c_int
            f.argtypes = [vp, C.POINTER(sz), sz, C.POINTER(C.c_uint64), sz, _RngHandle]
        L.q1t_sharded_reset_all.restype = C.c_int
        L.q1t_sharded_reset_all.argtypes = [vp]
        L.q1t_sharded_read_amplitudes.restype = C.c_int
        L.q1t_sharded_read_amplitudes.argtypes = [vp, sz, sz, C.POINTER(C.c_double)]
        L.q1t_sharded_column_total.restype = C.c_int
        L.q1t_sharded_column_total.argtypes = [vp, C.POINTER(C.c_double)]
        L.q1t_sharded_counters.restype = C.c_int
        L.q1t_sharded_counters.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.q1t_sharded_last_error.restype = C.c_char_p
        L.q1t_sharded_last_error.argtypes = [vp]
        self._L, self.nr_bits, self.nr_shots = L, nr_bits, nr_shots
        p = vp()
        dv = (C.c_int * len(devices))(*[int(d) for d in devices])
        rc = L.q1t_sharded_new(nr_bits, nr_shots, len(devices), dv, C.byref(p))
        if rc:
            raise EngineError(rc, L.q1t_sharded_last_error(None).decode())
        self._p = p

    def close(self):
        if getattr(self, "_p", None):
            self._L.q1t_sharded_free(self._p)
            self._p = None

    __del__ = close

    def _chk(self, rc):
        if rc:
            raise EngineError(rc, self._L.q1t_sharded_last_error(self._p).decode())

    def apply_gate(self, mat, bits, desc="gate"):
        m = np.ascontiguousarray(np.asarray(mat, dtype=np.complex128))
        self._chk(self._L.q1t_sharded_apply_gate(self._p, _dptr(m.view(np.float64)), m.shape[0], _szarr(bits), len(bits), desc.encode()))

    def set_initial_layout(self, dest):
        arr = (C.c_int * len(dest))(*[int(d) for d in dest])
        self._chk(self._L.q1t_sharded_set_initial_layout(self._p, arr, len(dest)))

    def measure_all_into(self, cbits, res, rng, collapse=True):
        f = self._L.q1t_sharded_measure_all_into if collapse else self._L.q1t_sharded_peek_all_into
        self._chk(f(self._p, _szarr(cbits), len(cbits), res.ctypes.data_as(C.POINTER(C.c_uint64)), res.size, rng.handle))

    def peek_all_into(self, cbits, res, rng):
        self.measure_all_into(cbits, res, rng, collapse=False)

    def reset_all(self):
        self._chk(self._L.q1t_sharded_reset_all(self._p))

    def amplitudes(self, offset=0, length=None):
        length = (1 << self.nr_bits) - offset if length is None else length
        out = np.zeros(2 * length, dtype=np.float64)
        self._chk(self._L.q1t_sharded_read_amplitudes(self._p, offset, length, _dptr(out)))
        return out.view(np.complex128)

    def column_total(self):
        v = C.c_double()
        self._chk(self._L.q1t_sharded_column_total(self._p, C.byref(v)))
        return v.value

    def counters(self):
        out = (C.c_uint64 * 3)()
        self._chk(self._L.q1t_sharded_counters(self._p, out))
        return {"remaps": int(out[0]), "exchanged_qubits": int(out[1]), "local_relabels": int(out[2])}
