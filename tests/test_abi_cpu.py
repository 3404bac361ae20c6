"""CPU-side checks of the product library (no GPU needed): the C-ABI library loads
and exports every symbol include/*.h declares, host logic (gate table, Binomial,
fusion planner, builder validation) matches the oracle / the reference's rules,
and the engine fails loudly without a device."""
import os
import re

import numpy as np
import pytest

from oracle import oracle as O
from q1tsim_b200 import circuit as QC
from q1tsim_b200 import engine as E
from q1tsim_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b((?:q1t|circuit|result)_[a-z0-9_]+)\s*\(", txt))
    return {n for n in names if not n.endswith("_fn")}


def test_library_exports_every_declared_symbol():
    L = E.lib()
    QC._lib()
    inner, outer = _declared("q1t_engine.h"), _declared("q1tsim_ffi.h")
    assert len(inner) >= 35 and len(outer) >= 28
    for name in sorted(inner | outer):
        assert hasattr(L, name), name
    assert set(E.INNER_ABI_SYMBOLS) <= inner
    assert set(QC.OUTER_ABI_SYMBOLS) <= outer | inner


@pytest.mark.skipif(not os.path.exists("/root/reference/python/q1tsimffi.py"), reason="reference not mounted")
def test_reference_python_cdef_binds_against_our_library():
    """python/q1tsimffi.py:35-82 is the authoritative C declaration of ffi.rs; cffi must be
    able to dlopen our .so with exactly that cdef and call the no-device entry points."""
    import cffi
    src = open("/root/reference/python/q1tsimffi.py").read()
    cdef = src.split('ffi.cdef("""')[1].split('""")')[0]
    ffi = cffi.FFI()
    ffi.cdef(cdef)
    lib = ffi.dlopen(E.LIB_PATH)
    c = lib.circuit_new(3, 3)
    assert lib.circuit_nr_qbits(c) == 3 and lib.circuit_nr_cbits(c) == 3
    r = lib.circuit_add_gate(c, b"h", [0], 1, ffi.NULL, 0)
    assert r.restype == 1
    lib.result_free(r)
    r = lib.circuit_add_gate(c, b"nope", [0], 1, ffi.NULL, 0)
    assert r.restype == 0 and ffi.string(ffi.cast("const char*", r.data)) == b'Unknown gate "nope"'
    lib.result_free(r)
    r = lib.circuit_measure(c, 0, 0, b"q", 1)
    assert r.restype == 0
    lib.result_free(r)
    r = lib.circuit_histogram(c)
    assert r.restype == 0 and ffi.string(ffi.cast("const char*", r.data)) == b"The circuit has not been executed yet"
    lib.result_free(r)
    lib.circuit_free(c)


def test_gate_table_matches_oracle():
    rs = np.random.default_rng(1)
    for name, npar in [("h", 0), ("i", 0), ("x", 0), ("y", 0), ("z", 0), ("s", 0), ("sdg", 0), ("t", 0), ("tdg", 0), ("v", 0),
                       ("vdg", 0), ("rx", 1), ("ry", 1), ("rz", 1), ("u1", 1), ("u2", 2), ("u3", 3), ("cx", 0), ("cy", 0),
                       ("cz", 0), ("ch", 0), ("cs", 0), ("csdg", 0), ("ct", 0), ("ctdg", 0), ("cv", 0), ("cvdg", 0),
                       ("swap", 0), ("crx", 1), ("cry", 1), ("crz", 1), ("cu1", 1), ("cu2", 2), ("cu3", 3), ("ccx", 0),
                       ("ccz", 0), ("ccrx", 1), ("ccry", 1), ("ccrz", 1)]:
        p = rs.uniform(-4, 4, size=npar)
        assert np.array_equal(E.gate_matrix(name, p), O.gate_matrix(name, p)), name
        assert np.array_equal(E.gate_matrix(name.upper(), p), O.gate_matrix(name, p))
    with pytest.raises(KeyError):
        E.gate_matrix("cch")
    with pytest.raises(ValueError):
        E.gate_matrix("u3", [1.0])


def test_host_sampling_matches_oracle_word_for_word():
    words = O.splitmix64_words(99, 20000)
    ro, re_ = O.Rng(words=words), E.Rng(words=words)
    cases = [(1024, 0.5), (8192, 0.125), (100, 0.01), (50000, 0.9991), (7, 0.3), (1, 0.5), (3000, 0.25), (10**6, 0.43), (12, 1.0), (12, 0.0)]
    for n, p in cases * 40:
        assert ro.binomial(n, p) == re_.binomial(n, p)
        assert ro.consumed == re_.consumed
    assert E.Rng(seed=5).binomial(100, 0.3) == O.Rng(seed=5).binomial(100, 0.3)


def _gates(ops):
    return [(E.gate_matrix(o[1], o[2]), o[3]) for o in ops if o[0] == "gate"]


def test_planner_fuses_qft30_into_three_sweeps():
    r = E.plan_dry_run(30, _gates(W.qft_ops(30, measure=False)), 12)
    assert r["sweeps"] == 3 and r["fallback"] == 0 and r["permute"] == 1
    assert r["ops"] <= 60                    # 30 H + <=30 fused phase ops for 480 gates
    r = E.plan_dry_run(34, _gates(W.qft_ops(34, measure=False)), 12)
    assert r["sweeps"] <= 4
    r = E.plan_dry_run(20, _gates(W.random_circuit_ops(20, 100, measure=False)), 12)
    assert r["fallback"] == 0 and r["sweeps"] < 1500 / 8
    r = E.plan_dry_run(24, _gates(W.ghz_branching_ops(24)), 12)
    assert r["sweeps"] <= 3 and r["fallback"] == 0


def test_planner_dense_blocks_fall_back():
    u = np.linalg.qr(np.random.default_rng(0).normal(size=(8, 8)) + 0j)[0]
    r = E.plan_dry_run(10, [(E.gate_matrix("h"), [0]), (u, [1, 5, 7]), (E.gate_matrix("cu1", [0.3]), [2, 3])], 12)
    assert r["fallback"] == 1
    with pytest.raises(E.EngineError):
        E.plan_dry_run(10, [(E.gate_matrix("cx"), [1, 1])], 12)


def test_builder_validation_without_device():
    c = QC.Circuit(3, 2)
    c.h(0); c.cx(0, 1); c.add_gate("cu1", [0, 2], [0.5]); c.measure(0, 1); c.barrier([0, 1])
    for call, msg in [(lambda: c.h(3), "Invalid index 3 for a quantum bit"),
                      (lambda: c.measure(0, 2), "Invalid index 2 for a classical bit"),
                      (lambda: c.measure(5, 0), "Invalid index 5 for a quantum bit"),
                      (lambda: c.measure_all([0, 7, 1]), "Invalid index 7 for a classical bit"),
                      (lambda: c.add_conditional_gate([4], 1, "x", [0]), "Invalid index 4 for a classical bit"),
                      (lambda: c.add_conditional_gate([0], 1, "x", [9]), "Invalid index 9 for a quantum bit"),
                      (lambda: c.reset(3), "Invalid index 3 for a quantum bit"),
                      (lambda: c.add_gate("frob", [0]), 'Unknown gate "frob"'),
                      (lambda: c.add_gate("u2", [0], [1.0]), 'Expected 2 arguments to "U2" gate, got 1'),
                      (lambda: c.measure_basis(0, 0, "Q"), "Invalid measurement basis '81'"),
                      (lambda: c.reexecute(), "The circuit has not been executed yet"),
                      (lambda: c.histogram(), "The circuit has not been executed yet"),
                      (lambda: c.cstate(), "Circuit has not been run yet")]:
        with pytest.raises(QC.CircuitError) as ei:
            call()
        assert str(ei.value) == msg
    assert c.nr_qbits() == 3 and c.nr_cbits() == 2


def test_no_cpu_fallback():
    """without a device the product path must fail loudly, never compute on the CPU"""
    if E.lib().q1t_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(E.EngineError) as ei:
        E.VectorState(3, 10)
    assert ei.value.kind == "CudaError" and "no CPU fallback" in str(ei.value)
    c = QC.Circuit(2, 2)
    c.h(0)
    with pytest.raises(QC.CircuitError) as ei:
        c.execute(10)
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "q1tsim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read().lower()
                assert "oracle" not in txt, os.path.join(dirpath, f)


def test_workload_generators_are_deterministic():
    a, b = W.random_circuit_ops(20, 100), W.random_circuit_ops(20, 100)
    assert a == b and W.gate_count(a) == 50 * 20 + 50 * 10
    q = W.qft_ops(30)
    assert W.gate_count(q) == 480 and q[-1][0] == "measure_all"
    g = W.ghz_branching_ops(24)
    assert sum(1 for o in g if o[0] == "measure") == 3 and sum(1 for o in g if o[0] == "cond") == 2
    assert W.SplitMix64(0).next_u64() == 0xE220A8397B1DCDAF
