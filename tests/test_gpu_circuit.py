"""GPU tests of the circuit layer through the ffi.rs-compatible C ABI
(include/q1tsim_ffi.h) against the oracle's restatement of circuit.rs:643-762.
Includes the reference's deterministic circuit tests (circuit.rs:1456-1985)."""
import math

import numpy as np
import pytest

from oracle import oracle as O
from q1tsim_b200 import circuit as QC
from q1tsim_b200 import engine as E
from q1tsim_b200 import workloads as W
from tests.test_oracle_reference_kats import measurement_ok

pytestmark = pytest.mark.gpu


def both(nq, nc, ops, shots, seed=1, nwords=None):
    words = O.splitmix64_words(seed, nwords or (4 * shots + 64))
    c = QC.Circuit(nq, nc)
    W.load_ops(c, ops)
    c.execute(shots, E.Rng(words=words))
    o = O.OracleCircuit(nq, nc, mode=1, order=1)
    W.load_ops(o, ops)
    o.execute(shots, O.Rng(words=words))
    return c, o


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def test_reference_test_execute():
    # circuit.rs:1456-1468
    c = QC.Circuit(2, 2)
    c.x(0); c.x(1); c.cx(0, 1); c.measure(0, 0); c.measure(1, 1)
    c.execute(5)
    assert list(c.cstate()) == [0b01] * 5


def test_reference_test_conditional():
    # circuit.rs:1619-1651
    c = QC.Circuit(2, 2)
    c.add_conditional_gate([0, 1], 1, "x", [1]); c.measure_all([0, 1])
    c.execute(5)
    assert list(c.cstate()) == [0] * 5
    for ctl, tgt, qb, exp in (([0, 1], 1, [1], [0b10, 0, 0, 0, 0]), ([0, 1], 2, [1], [0, 0b10, 0b10, 0, 0]), ([1], 1, [0], [0, 0b01, 0b01, 0b01, 0])):
        c = QC.Circuit(2, 2)
        c.set_cstate([0b01, 0b10, 0b10, 0b11, 0b00])
        c.add_conditional_gate(ctl, tgt, "x", qb); c.measure_all([0, 1])
        c.reexecute()
        assert list(c.cstate()) == exp


def test_reference_test_measure_all_and_basis():
    n = 1024
    c = QC.Circuit(2, 2); c.x(0); c.measure_all([0, 1]); c.execute(n)
    assert c.histogram_vec() == [0, n, 0, 0]                      # circuit.rs:1660-1665
    c = QC.Circuit(2, 2); c.x(0); c.measure_all([1, 0]); c.execute(n)
    assert c.histogram_vec() == [0, 0, n, 0]                      # circuit.rs:1667-1673
    c = QC.Circuit(2, 2); c.h(0); c.h(1); c.measure_all([0, 1]); c.execute(n)
    assert all(measurement_ok(v, n, 0.25, 1e-5) for v in c.histogram_vec())
    c = QC.Circuit(2, 2); c.h(0); c.h(1); c.measure_all_basis([0, 1], "X"); c.execute(n)
    assert c.histogram_vec() == [n, 0, 0, 0]                      # circuit.rs:1693-1700
    c = QC.Circuit(2, 2); c.x(0); c.h(0); c.h(1); c.measure_all_basis([0, 1], "X"); c.execute(n)
    assert c.histogram_vec() == [0, n, 0, 0]
    c = QC.Circuit(2, 2); c.x(0); c.h(0); c.h(1); c.s(0); c.s(1); c.measure_all_basis([0, 1], "Y"); c.execute(n)
    assert c.histogram_vec() == [0, n, 0, 0]
    c = QC.Circuit(2, 2); c.measure_all_basis([0, 1], "Y"); c.execute(n)
    assert all(measurement_ok(v, n, 0.25, 1e-5) for v in c.histogram_vec())


def test_reference_test_reset():
    n = 1024
    c = QC.Circuit(2, 2); c.h(0); c.z(0); c.reset(0); c.measure(0, 0); c.measure(1, 1); c.execute(n)
    assert c.histogram_vec() == [n, 0, 0, 0]                      # circuit.rs:1929-1940
    c = QC.Circuit(2, 2); c.h(0); c.z(0); c.x(1); c.reset(0); c.measure(0, 0); c.measure(1, 1); c.execute(n)
    assert c.histogram_vec() == [0, 0, n, 0]
    c = QC.Circuit(2, 2); c.h(0); c.z(0); c.h(1); c.reset(0); c.measure(0, 0); c.measure(1, 1); c.execute(n)
    hv = c.histogram_vec()
    assert measurement_ok(hv[0], n, 0.5, 1e-5) and hv[1] == 0 and measurement_ok(hv[2], n, 0.5, 1e-5) and hv[3] == 0
    c = QC.Circuit(5, 5); c.h(0); c.z(0); c.x(4); c.h(3); c.reset_all(); c.measure_all([0, 1, 2, 3, 4]); c.execute(n)
    hv = c.histogram_vec()
    assert hv[0] == n and not any(hv[1:])                         # circuit.rs:1969-1985


def test_cfg1_readme_qft3_through_ffi_abi():
    """README.md:46-80 with 8192 runs; bit-exact against the oracle for the same words"""
    ops = W.qft_ops(3, measure=True)
    c, o = both(3, 3, ops, 8192, seed=42)
    assert np.array_equal(c.cstate(), o.c_state)
    h = c.histogram()
    assert sum(h.values()) == 8192 and len(h) == 8
    assert all(measurement_ok(v, 8192, 0.125, 1e-5) for v in h.values())
    assert h == o.histogram_string()
    # measure_all collapses every distinct outcome into its own (lazy) basis column
    st, counts = c.state_columns()
    assert counts == o.q_state.counts and np.array_equal(st, o.q_state.states())


@pytest.mark.parametrize("n,depth", [(8, 12), (12, 30), (20, 100)])
def test_cfg2_random_circuit(n, depth):
    ops = W.random_circuit_ops(n, depth, measure=False) + [("peek_all", list(range(n)), "Z")]
    O.lib().orc_set_threads(8)
    c, o = both(n, n, ops, 1024)
    O.lib().orc_set_threads(1)
    st, counts = c.state_columns()
    assert counts == [1024]
    assert rel_l2(st[:, 0], o.q_state.column(0)) < 1e-10
    assert np.array_equal(c.cstate(), o.c_state)


@pytest.mark.parametrize("n", [5, 10, 14, 18])
def test_cfg4_ghz_branching(n):
    ops = W.ghz_branching_ops(n)
    ops_peek = ops[:-1] + [("peek_all", list(range(n)), "Z")]
    c, o = both(n, n, ops_peek, 1024, seed=5)
    assert np.array_equal(c.cstate(), o.c_state)
    st, counts = c.state_columns()
    assert counts == o.q_state.counts and 2 <= len(counts) <= 16
    so = o.q_state.states()
    for k in range(len(counts)):
        assert rel_l2(st[:, k], so[:, k]) < 1e-10
    # full version with the collapsing measure_all
    c, o = both(n, n, ops, 1024, seed=6)
    assert np.array_equal(c.cstate(), o.c_state)
    assert c.histogram_u64() == o.histogram()


def test_peek_sandwiches_and_y_basis():
    n = 6
    ops = W.u3_layer_ops(n, seed=3) + [("peek", 0, 0, "X"), ("peek", 3, 1, "Y"), ("measure", 5, 2, "Y"), ("peek_all", [3, 4, 5, 0, 1, 2], "X"),
                                      ("gate", "cx", (), [0, 5]), ("measure_all", [5, 4, 3, 2, 1, 0], "Y")]
    c, o = both(n, 6, ops, 512, seed=8)
    assert np.array_equal(c.cstate(), o.c_state)
    assert c.state_columns()[1] == o.q_state.counts


def test_reference_parameters_are_read_at_execute_time():
    """python/test.py: a RefParam changed between execute and reexecute takes effect"""
    ang = QC.RefParam(0.0)
    c = QC.Circuit(1, 1)
    c.rx(ang, 0); c.peek(0, 0)
    c.execute(1000, E.Rng(seed=1))
    assert c.histogram() == {"0": 1000}
    ang.assign(math.pi)
    c.reexecute(E.Rng(seed=2))
    assert c.histogram() == {"1": 1000}


def test_user_matrix_gates():
    n = 7
    rs = np.random.default_rng(3)
    u2 = np.linalg.qr(rs.normal(size=(4, 4)) + 1j * rs.normal(size=(4, 4)))[0]
    u3 = np.linalg.qr(rs.normal(size=(8, 8)) + 1j * rs.normal(size=(8, 8)))[0]
    c, o = QC.Circuit(n, n), O.OracleCircuit(n, n)
    for q in range(n):
        c.h(q); o.add_gate("h", [q])
    c.add_gate(u2, [4, 1]); o.add_matrix_gate(u2, [4, 1])
    c.add_gate(u3, [6, 0, 3]); o.add_matrix_gate(u3, [6, 0, 3])
    c.add_conditional_gate([0], 0, u2, [2, 5]); o.add_conditional_gate([0], 0, u2, [2, 5])
    words = O.splitmix64_words(1, 64)
    c.execute(16, E.Rng(words=words)); o.execute(16, O.Rng(words=words))
    st, counts = c.state_columns()
    assert counts == o.q_state.counts
    assert rel_l2(st[:, 0], o.q_state.column(0)) < 1e-10


def test_execute_errors_surface_like_the_reference():
    c = QC.Circuit(2, 2)
    c.add_gate("cx", [0])            # bit-count mismatch is detected at execution time (vectorstate.rs:169-174)
    with pytest.raises(QC.CircuitError) as ei:
        c.execute(4)
    assert str(ei.value) == 'Expected 2 bits for "CX", got 1'
    c = QC.Circuit(2, 2)
    c.measure_all([0])
    with pytest.raises(QC.CircuitError) as ei:
        c.execute(4)
    assert str(ei.value) == "Expected 2 measurement bits, but got 1"


def test_composite_and_loop_gates_flattened():
    """a circuit that uses `Composite::from_string` descriptions and a `Loop`
    (composite.rs:273-450, staticloop.rs:71-92) against the oracle fed with the same sub-gates one by one
    (the reference applies a composite as the product of its sub-gates, composite.rs:487-501) and against
    the composite's own matrix() used as one dense user gate."""
    nq, shots = 9, 256
    inc3 = "CCX 2 1 0; CX 2 1; X 2"
    rot = "RY(pi/3) 0; CU1(pi/4) 0 1; H 1; CRZ(0.5*(1.0-0.26)) 1 0"
    c = QC.Circuit(nq, nq)
    for q in range(nq):
        c.add_gate("h", [q])
    c.add_composite_gate("Inc3", inc3, [6, 2, 4])
    c.add_composite_gate("Rot", rot, [8, 0], nr_iterations=3)
    c.add_composite_gate("Inc3", inc3, [0, 1, 2])
    c.measure_all(list(range(nq)))
    words = O.splitmix64_words(3, 4 * shots + 64)
    c.execute(shots, E.Rng(words=words))

    def sub(desc, bits, reps=1):
        out = []
        for _ in range(reps):
            for part in desc.split(";"):
                name = part.split("(")[0].split()[0]
                args = ()
                if "(" in part:
                    inner = part[part.index("(") + 1:part.rindex(")")]
                    depth, cur, args = 0, "", []
                    for ch in inner:                      # split on top-level commas
                        if ch == "," and depth == 0:
                            args.append(cur); cur = ""
                        else:
                            depth += ch == "("; depth -= ch == ")"; cur += ch
                    args = tuple(E.eval_expression(a)[0] for a in args + [cur])
                    tail = part[part.rindex(")") + 1:]
                else:
                    tail = part.strip()[len(name):]
                out.append(("gate", name.lower(), args, [bits[int(b)] for b in tail.split()]))
        return out
    ops = [("gate", "h", (), [q]) for q in range(nq)] + sub(inc3, [6, 2, 4]) + sub(rot, [8, 0], 3) + sub(inc3, [0, 1, 2])
    o = O.OracleCircuit(nq, nq, mode=1, order=1)
    W.load_ops(o, ops + [("measure_all", list(range(nq)), "Z")])
    o.execute(shots, O.Rng(words=words))
    assert np.array_equal(c.cstate(), o.c_state)
    # the same circuit with each composite as ONE dense matrix gate
    e = E.VectorState(nq, 1)
    o2 = O.OracleState(nq, 1, mode=1, order=1)
    for q in range(nq):
        e.apply_gate(O.gate_matrix("h"), [q], "H")
    e.apply_gate(E.composite_matrix(inc3), [6, 2, 4], "Inc3")
    for _ in range(3):
        e.apply_gate(E.composite_matrix(rot), [8, 0], "Rot")
    e.apply_gate(E.composite_matrix(inc3), [0, 1, 2], "Inc3")
    for op in ops:
        o2.apply_gate(O.gate_matrix(op[1], op[2]), op[3])
    assert rel_l2(e.column(0), o2.column(0)) < 1e-10
