"""`Circuit` -- the reference's circuit API (src/circuit.rs, python/q1tsim.py) on the
B200 engine, through the ffi.rs-compatible C ABI of include/q1tsim_ffi.h.

Method names, argument meaning and error text follow the reference:
add_gate / add_conditional_gate / measure* / peek* / reset / reset_all / barrier /
execute / reexecute / histogram / cstate, plus the convenience builders h, x, y, z,
s, sdg, rx, ry, rz, u1, u2, u3, cx (circuit.rs:427-539).
"""
import ctypes as C

import numpy as np

from . import engine as E


class _Param(C.Structure):
    _fields_ = [("value", C.c_double), ("value_ptr", C.POINTER(C.c_double))]


class _Result(C.Structure):
    _fields_ = [("data", C.c_void_p), ("length", C.c_size_t), ("size", C.c_size_t), ("restype", C.c_uint32)]


class _HistElem(C.Structure):
    _fields_ = [("key", C.c_char_p), ("count", C.c_size_t)]


class _HistElemU64(C.Structure):
    _fields_ = [("key", C.c_uint64), ("count", C.c_size_t)]


RESULT_ERROR, RESULT_EMPTY, RESULT_STRING, RESULT_HISTOGRAM, RESULT_CSTATE, RESULT_HISTOGRAM_U64 = 0, 1, 2, 3, 5, 6

OUTER_ABI_SYMBOLS = [
    "result_free", "circuit_new", "circuit_free", "circuit_nr_qbits", "circuit_nr_cbits", "circuit_cstate",
    "circuit_add_gate", "circuit_add_conditional_gate", "circuit_measure", "circuit_measure_all", "circuit_reset",
    "circuit_reset_all", "circuit_execute", "circuit_reexecute", "circuit_histogram", "circuit_latex",
    "circuit_open_qasm", "circuit_c_qasm",
    # additive
    "circuit_add_matrix_gate", "circuit_add_conditional_matrix_gate", "circuit_barrier", "circuit_execute_with_rng",
    "circuit_reexecute_with_rng", "circuit_execute_with_qubit_coefs", "circuit_histogram_u64", "circuit_cstate_into",
    "circuit_set_cstate", "circuit_set_device", "circuit_state", "circuit_engine_stats", "circuit_add_composite_gate",
    "circuit_add_loop_gate", "circuit_nr_ops",
]

_bound = False


def _lib():
    global _bound
    L = E.lib()
    if _bound:
        return L
    sz, szp, vp, pp, u64p, dp = C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.POINTER(_Param), C.POINTER(C.c_uint64), C.POINTER(C.c_double)
    R, RNG = _Result, E._RngHandle
    sig = {
        "result_free": (None, [R]),
        "circuit_new": (vp, [sz, sz]), "circuit_free": (None, [vp]),
        "circuit_nr_qbits": (sz, [vp]), "circuit_nr_cbits": (sz, [vp]),
        "circuit_cstate": (R, [vp]),
        "circuit_add_gate": (R, [vp, C.c_char_p, szp, sz, pp, sz]),
        "circuit_add_conditional_gate": (R, [vp, szp, sz, C.c_uint64, C.c_char_p, szp, sz, pp, sz]),
        "circuit_measure": (R, [vp, sz, sz, C.c_char, C.c_uint8]),
        "circuit_measure_all": (R, [vp, szp, sz, C.c_char, C.c_uint8]),
        "circuit_reset": (R, [vp, sz]), "circuit_reset_all": (R, [vp]),
        "circuit_execute": (R, [vp, sz]), "circuit_reexecute": (R, [vp]),
        "circuit_histogram": (R, [vp]), "circuit_latex": (R, [vp]), "circuit_open_qasm": (R, [vp]), "circuit_c_qasm": (R, [vp]),
        "circuit_add_matrix_gate": (R, [vp, C.c_char_p, dp, sz, szp, sz]),
        "circuit_add_conditional_matrix_gate": (R, [vp, szp, sz, C.c_uint64, C.c_char_p, dp, sz, szp, sz]),
        "circuit_barrier": (R, [vp, szp, sz]),
        "circuit_add_composite_gate": (R, [vp, C.c_char_p, C.c_char_p, szp, sz, sz]),
        "circuit_add_loop_gate": (R, [vp, C.c_char_p, C.c_char_p, szp, sz, sz]),
        "circuit_nr_ops": (sz, [vp]),
        "circuit_execute_with_rng": (R, [vp, sz, RNG]), "circuit_reexecute_with_rng": (R, [vp, RNG]),
        "circuit_execute_with_qubit_coefs": (R, [vp, sz, RNG, dp]),
        "circuit_histogram_u64": (R, [vp]),
        "circuit_cstate_into": (sz, [vp, u64p, sz]),
        "circuit_set_cstate": (R, [vp, u64p, sz]),
        "circuit_set_device": (C.c_int, [vp, C.c_int]),
        "circuit_state": (vp, [vp]),
        "circuit_engine_stats": (C.c_int, [vp, C.POINTER(E.Stats)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _bound = True
    return L


class CircuitError(Exception):
    pass


def _unpack(res):
    """python/q1tsimffi.py:85-108 (unpack_result)"""
    if res.restype == RESULT_EMPTY:       # owns no data (ffi.rs:139-169): nothing to free
        return None
    L = _lib()
    try:
        if res.restype == RESULT_ERROR:
            raise CircuitError(C.cast(res.data, C.c_char_p).value.decode("utf-8"))
        if res.restype == RESULT_EMPTY:
            return None
        if res.restype == RESULT_STRING:
            return C.cast(res.data, C.c_char_p).value.decode("utf-8")
        if res.restype == RESULT_HISTOGRAM:
            el = C.cast(res.data, C.POINTER(_HistElem))
            return {el[i].key.decode("utf-8"): el[i].count for i in range(res.length)}
        if res.restype == RESULT_HISTOGRAM_U64:
            el = C.cast(res.data, C.POINTER(_HistElemU64))
            return {int(el[i].key): int(el[i].count) for i in range(res.length)}
        if res.restype == RESULT_CSTATE:
            p = C.cast(res.data, C.POINTER(C.c_uint64))
            return np.ctypeslib.as_array(p, shape=(res.length,)).copy() if res.length else np.zeros(0, dtype=np.uint64)
        raise CircuitError("Unknown data type code: %d" % res.restype)
    finally:
        L.result_free(res)


class RefParam:
    """By-reference gate parameter (python/q1tsim.py:5-38, parameter.rs FFIRef):
    the gate reads the value at *execute* time."""

    def __init__(self, value):
        self._v = C.c_double(value)

    def __float__(self):
        return self._v.value

    def assign(self, value):
        self._v.value = value

    def pointer(self):
        return C.pointer(self._v)


_ARR_T = {}          # ctypes array types by (element type, length): creating them is the slow part of a call
_NAME_B = {}         # gate name -> bytes


def _arr_t(elem, n):
    t = _ARR_T.get((elem, n))
    if t is None:
        t = _ARR_T[(elem, n)] = elem * n
    return t


def _name_b(name):
    b = _NAME_B.get(name)
    if b is None:
        b = _NAME_B[name] = name.encode()
    return b


def _params(values):
    arr = _arr_t(_Param, max(len(values), 1))()
    for i, v in enumerate(values):
        if isinstance(v, RefParam):
            arr[i].value = 0.0
            arr[i].value_ptr = v.pointer()
        else:
            arr[i].value = float(v)
            arr[i].value_ptr = None
    return arr


_SZ_CACHE = {}       # tuple of indices -> (read-only c_size_t array, length); the library copies what it is given
_PARAM_CACHE = {}    # tuple of plain float parameters -> read-only parameter_t array


def _sz(xs):
    key = tuple(xs)
    hit = _SZ_CACHE.get(key)
    if hit is None:
        n = len(key)
        hit = (_arr_t(C.c_size_t, max(n, 1))(*[int(x) for x in key]), n)
        if len(_SZ_CACHE) < 65536:
            _SZ_CACHE[key] = hit
    return hit


class Circuit:
    def __init__(self, nr_qbits, nr_cbits=0, device=0):
        self._L = _lib()
        self._p = self._L.circuit_new(nr_qbits, nr_cbits)
        self._keep = []          # keeps RefParam storage alive
        if device:
            self._L.circuit_set_device(self._p, device)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        if getattr(self, "_p", None):
            self._L.circuit_free(self._p)
            self._p = None

    __del__ = close

    def nr_qbits(self):
        return int(self._L.circuit_nr_qbits(self._p))

    def nr_cbits(self):
        return int(self._L.circuit_nr_cbits(self._p))

    # ---- builder ----
    def add_gate(self, name, qbits, params=()):
        if not isinstance(name, str):
            return self.add_matrix_gate(name, qbits)
        q, nq = _sz(qbits)
        if not params:
            return _unpack(self._L.circuit_add_gate(self._p, _name_b(name), q, nq, None, 0))
        params = tuple(params)
        arr = _PARAM_CACHE.get(params)             # only tuples of plain numbers are ever stored
        if arr is None:
            refs = [p for p in params if isinstance(p, RefParam)]
            arr = _params(params)
            if refs:
                self._keep.extend(refs)
            elif len(_PARAM_CACHE) < 65536:
                _PARAM_CACHE[params] = arr
        return _unpack(self._L.circuit_add_gate(self._p, _name_b(name), q, nq, arr, len(params)))

    def add_composite_gate(self, name, description, qbits, nr_iterations=1):
        """`Composite::from_string(name, description)` on `qbits` (composite.rs:273-450), body repeated
        `nr_iterations` times (`Loop`, staticloop.rs:71-92); flattened into the circuit's gate list."""
        q, nq = _sz(qbits)
        return _unpack(self._L.circuit_add_composite_gate(self._p, name.encode(), description.encode(), q, nq, int(nr_iterations)))

    def add_loop_gate(self, label, body_description, qbits, nr_iterations):
        """`Loop::new(label, nr_iterations, body)` (staticloop.rs:38-50): flattened like a composite, exported as a loop"""
        q, nq = _sz(qbits)
        return _unpack(self._L.circuit_add_loop_gate(self._p, label.encode(), body_description.encode(), q, nq, int(nr_iterations)))

    def nr_ops(self):
        return int(self._L.circuit_nr_ops(self._p))

    def add_matrix_gate(self, matrix, qbits, description="user gate"):
        m = np.ascontiguousarray(np.asarray(matrix, dtype=np.complex128))
        q, nq = _sz(qbits)
        return _unpack(self._L.circuit_add_matrix_gate(self._p, description.encode(), m.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)),
                                                       m.shape[0], q, nq))

    def add_conditional_gate(self, control, target, name, qbits, params=()):
        params = list(params or ())
        self._keep.extend(p for p in params if isinstance(p, RefParam))
        c, nc = _sz(control)
        q, nq = _sz(qbits)
        if not isinstance(name, str):
            m = np.ascontiguousarray(np.asarray(name, dtype=np.complex128))
            return _unpack(self._L.circuit_add_conditional_matrix_gate(self._p, c, nc, int(target), b"user gate",
                                                                       m.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)), m.shape[0], q, nq))
        return _unpack(self._L.circuit_add_conditional_gate(self._p, c, nc, int(target), name.encode(), q, nq,
                                                            _params(params) if params else None, len(params)))

    def h(self, q): return self.add_gate("h", [q])
    def x(self, q): return self.add_gate("x", [q])
    def y(self, q): return self.add_gate("y", [q])
    def z(self, q): return self.add_gate("z", [q])
    def s(self, q): return self.add_gate("s", [q])
    def sdg(self, q): return self.add_gate("sdg", [q])
    def rx(self, theta, q): return self.add_gate("rx", [q], [theta])
    def ry(self, theta, q): return self.add_gate("ry", [q], [theta])
    def rz(self, lam, q): return self.add_gate("rz", [q], [lam])
    def u1(self, lam, q): return self.add_gate("u1", [q], [lam])
    def u2(self, phi, lam, q): return self.add_gate("u2", [q], [phi, lam])
    def u3(self, theta, phi, lam, q): return self.add_gate("u3", [q], [theta, phi, lam])
    def cx(self, control, target): return self.add_gate("cx", [control, target])

    def measure_basis(self, qbit, cbit, basis="Z"):
        return _unpack(self._L.circuit_measure(self._p, qbit, cbit, basis.encode()[:1], 1))

    def measure_x(self, q, c): return self.measure_basis(q, c, "X")
    def measure_y(self, q, c): return self.measure_basis(q, c, "Y")
    def measure_z(self, q, c): return self.measure_basis(q, c, "Z")
    def measure(self, q, c): return self.measure_basis(q, c, "Z")

    def measure_all_basis(self, cbits, basis="Z"):
        c, n = _sz(cbits)
        return _unpack(self._L.circuit_measure_all(self._p, c, n, basis.encode()[:1], 1))

    def measure_all(self, cbits): return self.measure_all_basis(cbits, "Z")

    def peek_basis(self, qbit, cbit, basis="Z"):
        return _unpack(self._L.circuit_measure(self._p, qbit, cbit, basis.encode()[:1], 0))

    def peek(self, q, c): return self.peek_basis(q, c, "Z")

    def peek_all_basis(self, cbits, basis="Z"):
        c, n = _sz(cbits)
        return _unpack(self._L.circuit_measure_all(self._p, c, n, basis.encode()[:1], 0))

    def peek_all(self, cbits): return self.peek_all_basis(cbits, "Z")

    def reset(self, qbit):
        return _unpack(self._L.circuit_reset(self._p, qbit))

    def reset_all(self):
        return _unpack(self._L.circuit_reset_all(self._p))

    def barrier(self, qbits):
        q, n = _sz(qbits)
        return _unpack(self._L.circuit_barrier(self._p, q, n))

    # ---- execution (circuit.rs:562-641) ----
    def execute(self, nr_shots, rng=None, qubit_coefs=None):
        if qubit_coefs is not None:
            c = np.ascontiguousarray(np.asarray(qubit_coefs, dtype=np.complex128))
            rng = rng or E.Rng()
            return _unpack(self._L.circuit_execute_with_qubit_coefs(self._p, nr_shots, rng.handle,
                                                                    c.view(np.float64).ctypes.data_as(C.POINTER(C.c_double))))
        if rng is None:
            return _unpack(self._L.circuit_execute(self._p, nr_shots))
        return _unpack(self._L.circuit_execute_with_rng(self._p, nr_shots, rng.handle))

    def reexecute(self, rng=None):
        if rng is None:
            return _unpack(self._L.circuit_reexecute(self._p))
        return _unpack(self._L.circuit_reexecute_with_rng(self._p, rng.handle))

    # ---- results ----
    def cstate(self):
        return _unpack(self._L.circuit_cstate(self._p))

    def set_cstate(self, words):
        w = np.ascontiguousarray(np.asarray(words, dtype=np.uint64))
        return _unpack(self._L.circuit_set_cstate(self._p, w.ctypes.data_as(C.POINTER(C.c_uint64)), w.size))

    def histogram(self):
        """keys = bit strings, last character = classical bit 0 (circuit.rs:823-835)"""
        return _unpack(self._L.circuit_histogram(self._p))

    def histogram_u64(self):
        return _unpack(self._L.circuit_histogram_u64(self._p))

    def histogram_vec(self):
        out = [0] * (1 << self.nr_cbits())
        for k, v in self.histogram_u64().items():
            out[k] = v
        return out

    # ---- export (circuit.rs:877-1146; python/q1tsim.py open_qasm / c_qasm / latex) ----
    def open_qasm(self):
        return _unpack(self._L.circuit_open_qasm(self._p))

    def c_qasm(self):
        return _unpack(self._L.circuit_c_qasm(self._p))

    def latex(self):
        return _unpack(self._L.circuit_latex(self._p))

    def set_devices(self, devices):
        """execute() on a state sharded over these devices of this process (power of two >= 2; may repeat a device)"""
        L = self._L
        L.circuit_set_devices.restype = C.c_int
        L.circuit_set_devices.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_size_t]
        arr = (C.c_int * max(len(devices), 1))(*[int(d) for d in devices])
        if L.circuit_set_devices(self._p, arr, len(devices)):
            raise CircuitError("set_devices: the number of devices must be 0 or a power of two >= 2")

    def sharded_amplitudes(self, offset=0, length=None):
        L = self._L
        L.circuit_sharded_amplitudes.restype = C.c_int
        L.circuit_sharded_amplitudes.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_double)]
        n = int(self._L.circuit_nr_qbits(self._p))
        length = (1 << n) - offset if length is None else length
        out = np.zeros(2 * length, dtype=np.float64)
        if L.circuit_sharded_amplitudes(self._p, offset, length, out.ctypes.data_as(C.POINTER(C.c_double))):
            raise CircuitError("no sharded state (set_devices + execute first)")
        return out.view(np.complex128)

    def sharded_counters(self):
        L = self._L
        L.circuit_sharded_counters.restype = C.c_int
        L.circuit_sharded_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        out = (C.c_uint64 * 3)()
        L.circuit_sharded_counters(self._p, out)
        return {"remaps": int(out[0]), "exchanged_qubits": int(out[1]), "local_relabels": int(out[2])}

    def engine_stats(self):
        s = E.Stats()
        rc = self._L.circuit_engine_stats(self._p, C.byref(s))
        if rc:
            raise CircuitError("no engine state (circuit not executed)")
        return s.as_dict()

    def state_columns(self):
        """test hook: (2^n, C) amplitudes of the live state + counts"""
        st = self._L.circuit_state(self._p)
        if not st:
            raise CircuitError("The circuit has not been executed yet")
        L = self._L
        n = int(L.q1t_nr_bits(st))
        ncols = int(L.q1t_nr_columns(st))
        counts = (C.c_size_t * max(ncols, 1))()
        L.q1t_counts(st, counts)
        cols = []
        for c in range(ncols):
            out = np.empty(2 << n, dtype=np.float64)
            rc = L.q1t_read_amplitudes(st, c, 0, 1 << n, out.ctypes.data_as(C.POINTER(C.c_double)))
            if rc:
                raise CircuitError(L.q1t_last_error(st).decode())
            cols.append(out.view(np.complex128))
        return np.stack(cols, axis=1), list(counts)[:ncols]
