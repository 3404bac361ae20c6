"""Pins the CPU oracle against the reference's OWN deterministic unit tests
(known-answer vectors transcribed from /root/reference/src; tolerance 1e-15 abs
as `assert_complex_matrix_eq!`, cmatrix.rs:87-123).  Each test names the
reference test it transcribes."""
import math

import numpy as np
import pytest

from oracle import oracle as O

z, o, i = 0.0 + 0j, 1.0 + 0j, 1j
x = complex(math.sqrt(0.5))   # COMPLEX_HSQRT2
h = 0.5 * o
hx = 0.5 * x
TOL = 1e-15


def eq(a, b):
    a, b = np.asarray(a, dtype=np.complex128), np.asarray(b, dtype=np.complex128)
    assert a.shape == b.shape
    assert np.all(np.abs(a.real - b.real) <= TOL) and np.all(np.abs(a.imag - b.imag) <= TOL), (a, b)


def G(name, *p):
    return O.gate_matrix(name, p)


@pytest.fixture(params=[0, 1], ids=["faithful", "fast"])
def mode(request):
    return request.param


# ---- vectorstate.rs:425-470 --------------------------------------------
def test_new():
    s = O.OracleState(1, 1024)
    assert s.counts == [1024]
    eq(s.states(), [[o], [z]])
    s = O.OracleState(3, 1500)
    assert s.counts == [1500]
    eq(s.states(), [[o], [z], [z], [z], [z], [z], [z], [z]])


def test_from_qubit_coefs():
    s = O.OracleState.from_qubit_coefs([o, z, z, o], 1)
    assert (s.nr_bits, s.nr_shots, s.counts) == (2, 1, [1])
    eq(s.states(), [[z], [o], [z], [z]])
    s = O.OracleState.from_qubit_coefs([z, o, o, z], 13)
    assert s.counts == [13]
    eq(s.states(), [[z], [z], [o], [z]])
    s = O.OracleState.from_qubit_coefs([o, o, -i, z], 9)
    xx = math.sqrt(0.5) * i
    eq(s.states(), [[-xx], [z], [-xx], [z]])


# ---- vectorstate.rs:472-509 ---------------------------------------------
def test_apply_conditional_gate(mode):
    s = O.OracleState(2, 5, mode)
    s.apply_conditional_gate([0, 0, 1, 1, 0], G("x"), [1])
    assert s.counts == [2, 2, 1]
    eq(s.states(), [[o, z, o], [z, o, z], [z, z, z], [z, z, z]])

    s = O.OracleState(2, 5, mode)
    s.apply_conditional_gate([0, 0, 1, 1, 1], G("x"), [0])
    assert s.counts == [2, 3]
    eq(s.states(), [[o, z], [z, z], [z, o], [z, z]])

    s = O.OracleState(2, 5, mode)
    s.apply_conditional_gate([1, 0, 1, 1, 0], G("h"), [1])
    assert s.counts == [1, 1, 2, 1]
    eq(s.states(), [[x, o, x, o], [x, z, x, z], [z, z, z, z], [z, z, z, z]])

    s = O.OracleState.from_qubit_coefs([o, z, x, x], 5, mode)
    s.apply_conditional_gate([1, 0, 1, 1, 0], G("cx"), [1, 0])
    assert s.counts == [1, 1, 2, 1]
    eq(s.states(), [[x, x, x, x], [z, x, z, x], [z, z, z, z], [x, z, x, z]])

    s = O.OracleState(2, 5, mode)
    s.apply_conditional_gate([1, 1, 1, 0, 0], G("h"), [0])
    assert s.counts == [3, 2]
    eq(s.states(), [[x, o], [z, z], [x, z], [z, z]])
    s.apply_conditional_gate([0, 0, 1, 1, 1], G("h"), [0])
    assert s.counts == [2, 1, 2]
    eq(s.states(), [[x, o, x], [z, z, z], [x, z, x], [z, z, z]])


# ---- vectorstate.rs:511-590 (deterministic parts + structure) ------------
@pytest.mark.parametrize("order", [0, 1])
def test_measure(order):
    rng = O.Rng(seed=7)
    s = O.OracleState(1, 3, order=order)
    assert list(s.measure(0, rng)) == [0, 0, 0]
    eq(s.states(), [[o], [z]])

    s = O.OracleState.from_qubit_coefs([o, z, o, z], 3, order=order)
    assert list(s.measure(1, rng)) == [0, 0, 0]
    eq(s.states(), [[o], [z], [z], [z]])
    assert list(s.measure(0, rng)) == [0, 0, 0]
    eq(s.states(), [[o], [z], [z], [z]])

    s = O.OracleState.from_qubit_coefs([o, o, o, o], 1024, order=order)
    m0 = s.measure(0, rng)
    st = s.states()
    sc = 0
    prev = m0[0]
    for b in m0:
        if b != prev:
            sc += 1
            prev = b
        # tolerance 1e-15 in the reference; renormalisation by 1/sqrt(w0) is exact to 1 ulp
        exp = [x, x, z, z] if b == 0 else [z, z, x, x]
        assert np.allclose(st[:, sc], exp, atol=1e-15, rtol=0)
    m0b = s.measure(0, rng)
    assert np.array_equal(m0, m0b)
    m1 = s.measure(1, rng)
    st = s.states()
    sc = 0
    p0, p1 = m0[0], m1[0]
    table = {(0, 0): [o, z, z, z], (0, 1): [z, o, z, z], (1, 0): [z, z, o, z], (1, 1): [z, z, z, o]}
    for j in range(s.nr_shots):
        if m0[j] != p0 or m1[j] != p1:
            sc += 1
            p0, p1 = m0[j], m1[j]
        assert np.allclose(st[:, sc], table[(int(m0[j]), int(m1[j]))], atol=1e-15, rtol=0)


def get_bounds(nr_shots, p, tol):
    """stats.rs:10-19 (erf_inv from scipy instead of statrs)."""
    from scipy.special import erfinv
    mu = nr_shots * p
    sigma = math.sqrt(nr_shots * p * (1.0 - p))
    quantile = mu + sigma * math.sqrt(2.0) * erfinv(2.0 * tol - 1.0)
    return math.floor(quantile), math.ceil(mu + (mu - quantile))


def measurement_ok(count, nr_shots, p, tol):
    lo, hi = get_bounds(nr_shots, p, tol)
    return lo < count < hi


def test_stats_get_bounds():
    # stats.rs:39-48
    assert get_bounds(1024, 0.5, 1.0e-5) == (443, 581)
    assert get_bounds(1234, 0.25, 1.0e-5) == (243, 374)
    assert get_bounds(1234, 0.75, 1.0e-5) == (860, 991)
    assert get_bounds(1234, 0.75, 1.0e-10) == (828, 1023)
    assert get_bounds(1_000_000, 0.43, 1.0e-4) == (428158, 431842)


# ---- vectorstate.rs:592-638 ----------------------------------------------
@pytest.mark.parametrize("order", [0, 1])
def test_peek_into(order):
    n = 1024
    rng = O.Rng(seed=11)
    m = np.zeros(n, dtype=np.uint64)
    s = O.OracleState(1, n, order=order)
    s.peek_into(0, 0, m, rng)
    assert not m.any()
    eq(s.states(), [[o], [z]])

    s = O.OracleState.from_qubit_coefs([o, o], n, order=order)
    s.peek_into(0, 0, m, rng)
    assert measurement_ok(int(m.sum()), n, 0.5, 1e-5)
    eq(s.states(), [[x], [x]])

    s = O.OracleState.from_qubit_coefs([o, o, o, o], n, order=order)
    s.peek_into(0, 0, m, rng)
    assert measurement_ok(int(m.sum()), n, 0.5, 1e-5)
    m[:] = 0
    s.peek_into(1, 0, m, rng)
    assert measurement_ok(int(m.sum()), n, 0.5, 1e-5)
    eq(s.states(), [[h], [h], [h], [h]])

    s = O.OracleState.from_qubit_coefs([x, x, z, o], n, order=order)
    s.peek_into(0, 0, m, rng)
    assert measurement_ok(int(m.sum()), n, 0.5, 1e-5)
    m[:] = 0
    s.peek_into(1, 0, m, rng)
    assert int(m.sum()) == n
    eq(s.states(), [[z], [x], [z], [x]])


# ---- vectorstate.rs:640-707 ----------------------------------------------
def test_apply_unary_gate(mode):
    s = O.OracleState(3, 1, mode)
    s.apply_gate(G("h"), [0])
    eq(s.states(), [[x], [z], [z], [z], [x], [z], [z], [z]])
    s = O.OracleState(3, 1, mode)
    s.apply_gate(G("h"), [1])
    eq(s.states(), [[x], [z], [x], [z], [z], [z], [z], [z]])
    s = O.OracleState(3, 1, mode)
    s.apply_gate(G("y"), [2])
    eq(s.states(), [[z], [i], [z], [z], [z], [z], [z], [z]])


def test_apply_binary_gate(mode):
    s = O.OracleState(3, 1, mode)
    s.apply_gate(G("cx"), [0, 1])
    eq(s.states(), [[o], [z], [z], [z], [z], [z], [z], [z]])
    s = O.OracleState.from_qubit_coefs([z, o, o, z, o, z], 1, mode)
    s.apply_gate(G("cx"), [0, 1])
    eq(s.states(), [[z], [z], [z], [z], [z], [z], [o], [z]])
    s = O.OracleState.from_qubit_coefs([z, o, o, z, o, z], 1, mode)
    s.apply_gate(G("cx"), [0, 2])
    eq(s.states(), [[z], [z], [z], [z], [z], [o], [z], [z]])
    s = O.OracleState.from_qubit_coefs([z, o, o, z, o, z], 1, mode)
    hh = np.kron(G("h"), G("h"))        # Kron::new(H, H), kron.rs:59-62
    s.apply_gate(hh, [1, 2])
    eq(s.states(), [[z], [z], [z], [z], [h], [h], [h], [h]])


def test_apply_n_ary_gate(mode):
    s = O.OracleState(3, 1, mode)
    s.apply_gate(G("ccx"), [0, 1, 2])
    eq(s.states(), [[o], [z], [z], [z], [z], [z], [z], [z]])
    s = O.OracleState.from_qubit_coefs([z, o, z, o, o, z], 1, mode)
    s.apply_gate(G("ccx"), [0, 2, 1])
    eq(s.states(), [[z], [z], [z], [z], [z], [z], [o], [z]])
    s.apply_gate(G("ccx"), [0, 1, 2])
    eq(s.states(), [[z], [z], [z], [z], [z], [z], [z], [o]])
    s = O.OracleState.from_qubit_coefs([x, -x, x, -x, x, -x], 1, mode)
    s.apply_gate(G("ccx"), [0, 2, 1])
    eq(s.states(), [[hx], [-hx], [-hx], [hx], [-hx], [-hx], [hx], [hx]])


# ---- vectorstate.rs:709-771 ------------------------------------------------
@pytest.mark.parametrize("order", [0, 1])
def test_measure_all(order):
    rng = O.Rng(seed=3)
    s = O.OracleState.from_qubit_coefs([z, o, z, o, z, o], 5, order=order)
    r = s.measure_all(rng)
    assert r.shape == (5,) and np.all(r == 0b111)
    s = O.OracleState.from_qubit_coefs([z, o, z, o, o, z], 5, order=order)
    r = s.measure_all(rng)
    assert np.all(r == 0b011)
    s = O.OracleState(3, 5, order=order)
    s.apply_gate(G("h"), [2])
    r = s.measure_all(rng)
    assert np.all((r & np.uint64(0b011)) == 0)


@pytest.mark.parametrize("order", [0, 1])
def test_peek_all(order):
    n = 1024
    rng = O.Rng(seed=5)
    s = O.OracleState.from_qubit_coefs([z, o, z, o, z, o], n, order=order)
    s.apply_gate(G("h"), [0])
    s.apply_gate(G("h"), [2])
    res = np.zeros(n, dtype=np.uint64)
    s.peek_all_into([0, 1, 2], res, rng)
    assert s.counts == [n]
    assert np.allclose(s.states()[:, 0], [z, z, h, -h, z, z, -h, h], atol=1e-15, rtol=0)
    cnt = [int(((res >> np.uint64(b)) & np.uint64(1)).sum()) for b in range(3)]
    assert measurement_ok(cnt[0], n, 0.5, 1e-5)
    assert cnt[1] == n
    assert measurement_ok(cnt[2], n, 0.5, 1e-5)


# ---- vectorstate.rs:773-830 ----------------------------------------------
def test_reset(mode):
    rng = O.Rng(seed=9)
    s = O.OracleState.from_qubit_coefs([o, z], 10, mode)
    s.reset(0, rng)
    eq(s.states(), [[o], [z]])
    s = O.OracleState.from_qubit_coefs([z, o], 10, mode)
    s.reset(0, rng)
    eq(s.states(), [[o], [z]])
    s = O.OracleState.from_qubit_coefs([z, o, z, o], 10, mode)
    s.reset(0, rng)
    eq(s.states(), [[z], [o], [z], [z]])
    s = O.OracleState.from_qubit_coefs([z, o, z, o], 10, mode)
    s.reset(1, rng)
    eq(s.states(), [[z], [z], [o], [z]])
    s = O.OracleState.from_qubit_coefs([x, -x, o, z], 10, mode)
    s.reset(0, rng)
    st = s.states()
    if s.ncols == 1:
        assert np.allclose(st, [[o], [z], [z], [z]], atol=1e-15, rtol=0) or np.allclose(st, [[-o], [z], [z], [z]], atol=1e-15, rtol=0)
    else:
        assert s.ncols == 2
        assert np.allclose(st, [[o, -o], [z, z], [z, z], [z, z]], atol=1e-15, rtol=0)
    s = O.OracleState.from_qubit_coefs([x, -x, o, z], 10, mode)
    s.reset(1, rng)
    eq(s.states(), [[x], [z], [-x], [z]])


def test_reset_all():
    s = O.OracleState(5, 100)
    s.apply_gate(G("h"), [2])
    s.apply_gate(G("x"), [0])
    s.apply_gate(G("h"), [4])
    s.reset_all()
    assert s.counts == [100]
    exp = np.zeros((32, 1), dtype=np.complex128)
    exp[0, 0] = 1
    eq(s.states(), exp)


# ---- support.rs:99-118 ---------------------------------------------------
def test_reverse_bits():
    assert O.reverse_bits(1, 1) == 1
    assert O.reverse_bits(1, 4) == 8
    assert O.reverse_bits(10, 4) == 5
    assert O.reverse_bits(26, 4) == 5
    assert O.reverse_bits(0xfffffffffffffffa, 4) == 0x5
    assert O.reverse_bits(0xffffffffffffaaaa, 32) == 0x5555ffff
    assert O.reverse_bits(0x1ffffffffffffffa, 64) == 0x5ffffffffffffff8


def test_shuffle_bits():
    assert O.shuffle_bits(0xc, [0, 3, 1, 2]) == 0x6
    assert O.shuffle_bits(0xfffffffffffffffa, [8, 9, 10, 11]) == 0xa00
    assert O.shuffle_bits(0xf555555555555555, [63, 62, 61, 60]) == 0xa000000000000000
    assert O.shuffle_bits(0x3, [3, 2, 1, 0]) == 0xc


# ---- gates: test_matrix of the gate files ---------------------------------
def test_gate_matrices():
    eq(G("h"), [[x, x], [x, -x]])                        # hadamard.rs:186-191
    eq(G("x"), [[z, o], [o, z]])
    eq(G("y"), [[z, -i], [i, z]])
    eq(G("z"), [[o, z], [z, -o]])
    eq(G("s"), [[o, z], [z, i]])
    eq(G("sdg"), [[o, z], [z, -i]])
    t = x + x * i
    eq(G("t"), [[o, z], [z, t]])                         # t.rs:246-252
    eq(G("tdg"), [[o, z], [z, t.conjugate()]])
    eq(G("v"), [[h + h * i, h - h * i], [h - h * i, h + h * i]])
    eq(G("vdg"), [[h - h * i, h + h * i], [h + h * i, h - h * i]])
    eq(G("swap"), [[o, z, z, z], [z, z, o, z], [z, o, z, z], [z, z, z, o]])   # swap.rs:189-199
    eq(G("cx"), [[o, z, z, z], [z, o, z, z], [z, z, z, o], [z, z, o, z]])     # cx.rs:144-153
    eq(G("cz"), [[o, z, z, z], [z, o, z, z], [z, z, o, z], [z, z, z, -o]])
    eq(G("cy"), [[o, z, z, z], [z, o, z, z], [z, z, z, -i], [z, z, i, z]])
    # circuit.rs:1289-1453 builder matrices with parameters
    th = 1.2345
    c, s_ = math.cos(th / 2), math.sin(th / 2)
    eq(G("rx", th), [[c, -1j * s_], [-1j * s_, c]])
    eq(G("ry", th), [[c, -s_], [s_, c]])
    eq(G("rz", th), [[complex(c, -s_), z], [z, complex(c, s_)]])
    eq(G("u1", th), [[o, z], [z, complex(math.cos(th), math.sin(th))]])
    phi, lam = 0.4, -2.2
    eq(G("u2", phi, lam), np.array([[1, -np.exp(1j * lam)], [np.exp(1j * phi), np.exp(1j * (phi + lam))]]) * math.sqrt(0.5))
    eq(G("u3", th, phi, lam), [[c, -np.exp(1j * lam) * s_], [np.exp(1j * phi) * s_, np.exp(1j * (phi + lam)) * c]])
    ccx = np.eye(8, dtype=np.complex128)
    ccx[6:, 6:] = [[0, 1], [1, 0]]
    eq(G("ccx"), ccx)                                    # controlled.rs:60-69, composite tests
    cu1 = np.eye(4, dtype=np.complex128)
    cu1[3, 3] = np.exp(1j * th)
    eq(G("cu1", th), cu1)
    with pytest.raises(KeyError):
        G("foo")
    with pytest.raises(ValueError):
        G("rx")


# ---- gates.rs:53-80 closed form + permutation.rs semantics --------------------
def test_bit_permutation_closed_form():
    import itertools
    for n in (2, 3, 4):
        for k in (2, 3):
            if k > n:
                continue
            for bits in itertools.permutations(range(n), k):
                perm = O.bit_permutation(n, list(bits))
                rest = [q for q in range(n) if q not in bits]
                order = list(bits) + rest          # qubits from MSB to LSB of the permuted index
                for new in range(1 << n):
                    old = 0
                    for j, q in enumerate(order):
                        v = (new >> (n - 1 - j)) & 1
                        old |= v << (n - 1 - q)
                    assert perm[new] == old


# ---- qustate.rs:100-127 -----------------------------------------------------
def test_collect_conditional_ranges():
    assert O.collect_conditional_ranges([5], [0, 0, 1, 1, 0]) == [(0, 2, False), (0, 2, True), (0, 1, False)]
    assert O.collect_conditional_ranges([3, 2], [1, 1, 1, 0, 0]) == [(0, 3, True), (1, 2, False)]
    assert O.collect_conditional_ranges([3, 2], [0, 0, 1, 1, 1]) == [(0, 2, False), (0, 1, True), (1, 2, True)]


# ---- circuit.rs deterministic tests ---------------------------------------------
def test_circuit_execute():
    # circuit.rs:1456-1468
    c = O.OracleCircuit(2, 2)
    c.add_gate("x", [0]); c.add_gate("x", [1]); c.add_gate("cx", [0, 1])
    c.measure(0, 0); c.measure(1, 1)
    c.execute(5, O.Rng(seed=1))
    assert list(c.c_state) == [0b01] * 5


def test_circuit_conditional():
    # circuit.rs:1619-1651
    c = O.OracleCircuit(2, 2)
    c.add_conditional_gate([0, 1], 1, "x", [1])
    c.measure_all([0, 1])
    c.execute(5, O.Rng(seed=1))
    assert list(c.c_state) == [0] * 5
    for ctl, tgt, qb, exp in (([0, 1], 1, [1], [0b10, 0, 0, 0, 0]),
                              ([0, 1], 2, [1], [0, 0b10, 0b10, 0, 0]),
                              ([1], 1, [0], [0, 0b01, 0b01, 0b01, 0])):
        c = O.OracleCircuit(2, 2)
        c.q_state = O.OracleState(2, 5)
        c.c_state = np.array([0b01, 0b10, 0b10, 0b11, 0b00], dtype=np.uint64)
        c.add_conditional_gate(ctl, tgt, "x", qb)
        c.measure_all([0, 1])
        c.reexecute(O.Rng(seed=2))
        assert list(c.c_state) == exp


def test_circuit_measure_all_routing():
    # circuit.rs:1654-1673 (first two cases)
    c = O.OracleCircuit(2, 2)
    c.add_gate("x", [0]); c.measure_all([0, 1])
    c.execute(1024, O.Rng(seed=1))
    assert c.histogram_vec() == [0, 1024, 0, 0]
    c = O.OracleCircuit(2, 2)
    c.add_gate("x", [0]); c.measure_all([1, 0])
    c.execute(1024, O.Rng(seed=1))
    assert c.histogram_vec() == [0, 0, 1024, 0]
    c = O.OracleCircuit(2, 2)
    c.add_gate("h", [0]); c.add_gate("h", [1]); c.measure_all([0, 1])
    c.execute(1024, O.Rng(seed=1))
    assert all(measurement_ok(v, 1024, 0.25, 1e-5) for v in c.histogram_vec())


def test_circuit_measure_all_basis():
    # circuit.rs:1688-1722.  NOTE: the reference routes these all-Clifford
    # circuits to its stabilizer backend; on the vector path the X-basis
    # sandwich leaves ~1e-33 probability on other outcomes, which the
    # canonical sampler never selects (chosen < total always lands on a
    # non-zero weight).
    n = 1024
    c = O.OracleCircuit(2, 2)
    c.add_gate("h", [0]); c.add_gate("h", [1]); c.measure_all_basis([0, 1], "X")
    c.execute(n, O.Rng(seed=1))
    assert c.histogram_vec() == [n, 0, 0, 0]
    c = O.OracleCircuit(2, 2)
    c.add_gate("x", [0]); c.add_gate("h", [0]); c.add_gate("h", [1]); c.measure_all_basis([0, 1], "X")
    c.execute(n, O.Rng(seed=1))
    assert c.histogram_vec() == [0, n, 0, 0]
    c = O.OracleCircuit(2, 2)
    c.add_gate("x", [0]); c.add_gate("h", [0]); c.add_gate("h", [1]); c.add_gate("s", [0]); c.add_gate("s", [1])
    c.measure_all_basis([0, 1], "Y")
    c.execute(n, O.Rng(seed=1))
    assert c.histogram_vec() == [0, n, 0, 0]
    c = O.OracleCircuit(2, 2)
    c.measure_all_basis([0, 1], "Y")
    c.execute(n, O.Rng(seed=1))
    assert all(measurement_ok(v, n, 0.25, 1e-5) for v in c.histogram_vec())


def test_circuit_reset():
    # circuit.rs:1924-1966
    n = 1024
    c = O.OracleCircuit(2, 2)
    c.add_gate("h", [0]); c.add_gate("z", [0]); c.reset(0); c.measure(0, 0); c.measure(1, 1)
    c.execute(n, O.Rng(seed=1))
    assert c.histogram_vec() == [n, 0, 0, 0]
    c = O.OracleCircuit(2, 2)
    c.add_gate("h", [0]); c.add_gate("z", [0]); c.add_gate("x", [1]); c.reset(0); c.measure(0, 0); c.measure(1, 1)
    c.execute(n, O.Rng(seed=1))
    assert c.histogram_vec() == [0, 0, n, 0]
    c = O.OracleCircuit(2, 2)
    c.add_gate("h", [0]); c.add_gate("z", [0]); c.add_gate("h", [1]); c.reset(0); c.measure(0, 0); c.measure(1, 1)
    c.execute(n, O.Rng(seed=1))
    hv = c.histogram_vec()
    assert measurement_ok(hv[0], n, 0.5, 1e-5) and hv[1] == 0 and measurement_ok(hv[2], n, 0.5, 1e-5) and hv[3] == 0


def test_circuit_reset_all():
    # circuit.rs:1969-1985
    c = O.OracleCircuit(5, 5)
    c.add_gate("h", [0]); c.add_gate("z", [0]); c.add_gate("x", [4]); c.add_gate("h", [3])
    c.reset_all(); c.measure_all([0, 1, 2, 3, 4])
    c.execute(1024, O.Rng(seed=1))
    hv = c.histogram_vec()
    assert hv[0] == 1024 and not any(hv[1:])


def test_readme_qft3():
    # README.md:59-68 (cfg1): uniform 1/sqrt(8), histogram ~1024 each
    c = O.OracleCircuit(3, 3)
    c.add_gate("h", [2]); c.add_gate("cs", [1, 2]); c.add_gate("ct", [0, 2])
    c.add_gate("h", [1]); c.add_gate("cs", [0, 1]); c.add_gate("h", [0]); c.add_gate("swap", [0, 2])
    c.peek_all_basis([0, 1, 2], "Z")
    c.execute(8192, O.Rng(seed=42))
    assert np.allclose(c.q_state.states()[:, 0], np.full(8, 1 / math.sqrt(8)), atol=1e-15)
    hv = c.histogram_vec()
    assert sum(hv) == 8192 and all(measurement_ok(v, 8192, 0.125, 1e-5) for v in hv)
