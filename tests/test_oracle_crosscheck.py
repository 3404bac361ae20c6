"""Cross-checks of the C oracle: vs an independent numpy dense restatement,
faithful vs fast structure (bit-exact), QFT closed form,
canonical vs reference summation order, sampler and Binomial sanity."""
import math

import numpy as np
import pytest

from oracle import oracle as O
from q1tsim_b200 import workloads as W
from tests import np_ref
from tests.test_oracle_reference_kats import measurement_ok


def rand_state(n, seed):
    r = np.random.default_rng(seed)
    v = r.normal(size=1 << n) + 1j * r.normal(size=1 << n)
    return v / np.linalg.norm(v)


def rand_unitary(k, seed):
    r = np.random.default_rng(seed)
    a = r.normal(size=(1 << k, 1 << k)) + 1j * r.normal(size=(1 << k, 1 << k))
    q, _ = np.linalg.qr(a)
    return q


@pytest.mark.parametrize("n,k", [(3, 1), (4, 2), (5, 3), (6, 4), (7, 2), (8, 3)])
def test_random_gates_vs_numpy(n, k):
    rs = np.random.default_rng(100 * n + k)
    psi = rand_state(n, n)
    sf, sq = O.OracleState(n, 1, mode=0), O.OracleState(n, 1, mode=1)
    sf.set_column(0, psi); sq.set_column(0, psi)
    ref = psi.copy()
    for t in range(6):
        bits = [int(b) for b in rs.permutation(n)[:k]]
        u = rand_unitary(k, 17 * t + n)
        sf.apply_gate(u, bits); sq.apply_gate(u, bits)
        ref = np_ref.apply_gate(ref, u, bits, n)
    assert np.array_equal(sf.column(0), sq.column(0))          # same arithmetic, different loop structure
    assert np.linalg.norm(sq.column(0) - ref) < 1e-13


def test_named_gates_vs_numpy():
    n = 5
    psi = rand_state(n, 3)
    s = O.OracleState(n, 1)
    s.set_column(0, psi)
    ref = psi.copy()
    for name, params, bits in [("h", (), [4]), ("cx", (), [3, 0]), ("ccx", (), [4, 1, 2]), ("cu1", (0.3,), [0, 4]),
                               ("swap", (), [1, 3]), ("crz", (1.1,), [2, 0]), ("ccrx", (0.7,), [0, 3, 1]),
                               ("u3", (0.1, 0.2, 0.3), [2]), ("cu3", (1.0, 2.0, 3.0), [4, 2]), ("cvdg", (), [1, 0])]:
        m = O.gate_matrix(name, params)
        s.apply_gate(m, bits)
        ref = np_ref.apply_gate(ref, m, bits, n)
    assert np.linalg.norm(s.column(0) - ref) < 1e-14


@pytest.mark.parametrize("n", [3, 4, 5, 6, 9])
def test_qft_closed_form(n):
    for x in (0, 1, (1 << n) - 2, 5 % (1 << n)):
        c = O.OracleCircuit(n, n)
        for q in range(n):
            if (x >> (n - 1 - q)) & 1:
                c.add_gate("x", [q])
        W.load_ops(c, W.qft_ops(n, measure=False))
        c.execute(1, O.Rng(seed=1))
        assert np.linalg.norm(c.q_state.column(0) - np_ref.qft_closed_form(n, x)) < 1e-13


def test_multi_column_apply_matches_per_column():
    n = 4
    s = O.OracleState(n, 6)
    s.apply_gate(O.gate_matrix("h"), [0])
    s.apply_conditional_gate([1, 0, 0, 1, 1, 0], O.gate_matrix("ry", [0.4]), [2])
    assert s.ncols == 4
    before = s.states()
    u = rand_unitary(2, 5)
    s.apply_gate(u, [3, 1])
    after = s.states()
    for c in range(4):
        assert np.linalg.norm(after[:, c] - np_ref.apply_gate(before[:, c], u, [3, 1], n)) < 1e-14


@pytest.mark.parametrize("n", [1, 4, 9, 12, 14])
def test_canonical_vs_reference_order(n):
    s = O.OracleState(n, 1)
    s.set_column(0, rand_state(n, n + 50))
    for q in {0, n // 2, n - 1}:
        w_ref, w_can = s.marginal0(q, order=0)[0], s.marginal0(q, order=1)[0]
        assert abs(w_ref - w_can) < 1e-14
        assert abs(w_can - np_ref.marginal0(s.column(0), q, n)) < 1e-14
    assert abs(s.column_totals(order=1)[0] - 1.0) < 1e-14


def test_measure_all_orders_agree_as_multisets():
    # the two summation orders resolve (almost surely) the same outcomes for the same uniforms
    n, shots = 10, 4096
    psi = rand_state(n, 77)
    words = O.splitmix64_words(2, shots)
    out = []
    for order in (0, 1):
        s = O.OracleState(n, shots, order=order)
        s.set_column(0, psi)
        res = np.zeros(shots, dtype=np.uint64)
        s.peek_all_into(list(range(n)), res, O.Rng(words=words))
        out.append(np.sort(res))
    assert np.array_equal(out[0], out[1])
    # canonical order is grouped ascending in basis index -> bit-reversed words non-decreasing
    s = O.OracleState(n, shots, order=1)
    s.set_column(0, psi)
    res = np.zeros(shots, dtype=np.uint64)
    s.peek_all_into(list(range(n)), res, O.Rng(words=words))
    idx = np.array([O.reverse_bits(int(r), n) for r in res])
    assert np.all(np.diff(idx) >= 0)
    # statistics: chi-square-ish check against |psi|^2 on the 8 most likely outcomes
    p = np.abs(psi) ** 2
    for i in np.argsort(p)[-8:]:
        assert measurement_ok(int((idx == i).sum()), shots, float(p[i]), 1e-6)


def test_measure_all_collapse_columns():
    n, shots = 3, 64
    s = O.OracleState(n, shots, order=1)
    s.apply_gate(O.gate_matrix("h"), [0]); s.apply_gate(O.gate_matrix("h"), [2])
    res = np.zeros(shots, dtype=np.uint64)
    s.measure_all_into([0, 1, 2], res, O.Rng(seed=4))
    st = s.states()
    assert sum(s.counts) == shots and s.ncols == 4
    off = 0
    for c, cnt in enumerate(s.counts):
        idx = int(np.argmax(np.abs(st[:, c])))
        assert st[idx, c] == 1.0 and np.count_nonzero(st[:, c]) == 1
        word = O.shuffle_bits(O.reverse_bits(idx, n), [0, 1, 2])
        assert np.all(res[off:off + cnt] == word)
        off += cnt


def test_binomial_edges_and_statistics():
    r = O.Rng(seed=123)
    assert r.binomial(100, 0.0) == 0 and r.binomial(100, 1.0) == 100
    for n, p in [(20, 0.3), (1024, 0.5), (8192, 0.125), (100000, 0.999), (50, 0.01), (1000, 0.9)]:
        xs = np.array([r.binomial(n, p) for _ in range(4000)], dtype=np.float64)
        assert xs.min() >= 0 and xs.max() <= n
        mu, sd = n * p, math.sqrt(n * p * (1 - p))
        assert abs(xs.mean() - mu) < 5 * sd / math.sqrt(xs.size) + 1e-9
        assert abs(xs.std() - sd) < 0.1 * sd + 1e-9


def test_rng_array_exhaustion_reports():
    s = O.OracleState.from_qubit_coefs([1, 1], 10)
    res = np.zeros(10, dtype=np.uint64)
    with pytest.raises(O.OracleError) as e:
        s.measure_into(0, 0, res, O.Rng(words=[]))
    assert e.value.kind == "RngExhausted"


def test_error_codes():
    s = O.OracleState(2, 4)
    res = np.zeros(2, dtype=np.uint64)
    with pytest.raises(O.OracleError) as e:
        s.measure_into(5, 0, np.zeros(4, dtype=np.uint64), O.Rng(seed=1))
    assert e.value.kind == "InvalidQBit"
    with pytest.raises(O.OracleError) as e:
        s.measure_into(0, 0, res, O.Rng(seed=1))
    assert e.value.kind == "NotEnoughSpace"
    with pytest.raises(O.OracleError) as e:
        s.measure_all_into([0], np.zeros(4, dtype=np.uint64), O.Rng(seed=1))
    assert e.value.kind == "InvalidNrMeasurementBits"
    with pytest.raises(O.OracleError) as e:
        s.apply_conditional_gate([1, 0], O.gate_matrix("x"), [0])
    assert e.value.kind == "InvalidNrControlBits"
    with pytest.raises(O.OracleError) as e:
        s.apply_gate(O.gate_matrix("cx"), [0])
    assert e.value.kind == "InvalidNrBits"


@pytest.mark.parametrize("n", [3, 8, 13])
def test_closed_form_of_qft_on_a_product_state(n):
    """workloads.qft_of_product_state (the size-independent check the GPU tests and bench.py use at n = 30 on a dense
    input) pinned against the oracle applying the QFT gate by gate"""
    from q1tsim_b200 import workloads as W
    coefs = W.product_state_coefs(n, seed=n)
    o = O.OracleState.from_qubit_coefs(coefs, 1)
    for op in W.qft_ops(n, measure=False):
        o.apply_gate(O.gate_matrix(op[1], op[2]), op[3])
    want = W.qft_of_product_state(n, coefs, np.arange(1 << n))
    assert np.linalg.norm(o.column(0) - want) < 1e-13
