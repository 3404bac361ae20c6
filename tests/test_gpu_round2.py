"""GPU tests added in round 2: the advisor's lazy-column / pending-relabel regression, parity at BASELINE sizes
(cfg3 dense input at n = 30 against the closed form, cfg4 GHZ-24 branching against the oracle, measure_all
bit-exact at n = 24), and the TMA-loaded dense ladder sweeps against the cp.async ones."""
import numpy as np
import pytest

from oracle import oracle as O
from q1tsim_b200 import circuit as QC
from q1tsim_b200 import engine as E
from q1tsim_b200 import workloads as W

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def both(nq, nc, build, shots, seed=1, nwords=None):
    words = O.splitmix64_words(seed, nwords or (6 * shots + 64))
    c, o = QC.Circuit(nq, nc), O.OracleCircuit(nq, nc, mode=1, order=1)
    build(c)
    build(o)
    c.execute(shots, E.Rng(words=words))
    o.execute(shots, O.Rng(words=words))
    return c, o


@pytest.mark.parametrize("n", [5, 6, 9, 13])
def test_lazy_columns_with_pending_swap_relabel(n):
    """after measure_all every column is a lazy basis state kept by LOGICAL index; a conditional X
    makes some columns dense and leaves others lazy; a Swap is then a zero-byte relabel.  Materialising the lazy
    columns while the relabel is pending must write the 1 at the PHYSICAL position (engine.cu materialize)."""
    def build(c):
        c.add_gate("x", [1]); c.add_gate("h", [0])
        c.measure_all(list(range(n)))
        c.reset(0)                                   # conditional X: one column dense, one lazy
        c.add_gate("swap", [0, 1]); c.add_gate("h", [2])
        c.add_gate("swap", [1, n - 1]); c.add_gate("ccx", [0, 2, n - 2])
        c.peek_all_basis(list(range(n)), "Z")
    c, o = both(n, n, build, 256, seed=n)
    assert np.array_equal(c.cstate(), o.c_state)
    st, counts = c.state_columns()
    assert counts == o.q_state.counts
    so = o.q_state.states()
    for k in range(len(counts)):
        assert rel_l2(st[:, k], so[:, k]) < 1e-10


@pytest.mark.parametrize("n", [5, 8, 12])
def test_lazy_columns_swap_then_dense_matrix_gate(n):
    """same, through the generic-gate fallback: measure_all -> swap -> user 2-qubit matrix gate"""
    rs = np.random.default_rng(n)
    u2 = np.linalg.qr(rs.normal(size=(4, 4)) + 1j * rs.normal(size=(4, 4)))[0]

    def build(c):
        for q in range(n):
            c.add_gate("h", [q])
        c.measure_all(list(range(n)))
        c.add_gate("swap", [0, n - 1]); c.add_gate("swap", [1, 2])
        c.add_matrix_gate(u2, [0, 2])
        c.add_gate("h", [1])
        c.peek_all_basis(list(range(n)), "Z")
    c, o = both(n, n, build, 64, seed=3 + n)
    assert np.array_equal(c.cstate(), o.c_state)
    st, counts = c.state_columns()
    assert counts == o.q_state.counts
    so = o.q_state.states()
    for k in range(len(counts)):
        assert rel_l2(st[:, k], so[:, k]) < 1e-10


def _dense_qft(n, options):
    coefs = W.product_state_coefs(n, seed=n)
    st = E.VectorState.from_qubit_coefs(coefs, 1)
    for k, v in options.items():
        st.set_option(k, v)
    for op in W.qft_ops(n, measure=False):
        st.apply_gate(E.gate_matrix(op[1], op[2]), op[3], op[1])
    st.flush()
    return st, coefs


@pytest.mark.parametrize("n", [13, 16, 21, 24])
def test_tma_sweeps_equal_cp_async_sweeps(n):
    """dense ladder sweeps load their tiles by TMA (cp.async.bulk.tensor, planner.cpp apply_tma_layout): same
    arithmetic on the same values, so the result is bit-identical to the cp.async path, and both match the oracle"""
    a, coefs = _dense_qft(n, {"tma": 1})
    b, _ = _dense_qft(n, {"tma": 0})
    assert a.stats()["tma_sweeps"] > 0 and b.stats()["tma_sweeps"] == 0
    ca, cb = a.column(0), b.column(0)
    assert np.array_equal(ca, cb)
    if n <= 21:
        o = O.OracleState.from_qubit_coefs(coefs, 1)
        O.lib().orc_set_threads(8)
        for op in W.qft_ops(n, measure=False):
            o.apply_gate(O.gate_matrix(op[1], op[2]), op[3])
        O.lib().orc_set_threads(1)
        assert rel_l2(ca, o.column(0)) < 1e-10
    else:
        assert rel_l2(ca, W.qft_of_product_state(n, coefs, np.arange(1 << n))) < 1e-10


@pytest.mark.parametrize("n", [22, 24])
def test_relabelling_stores_in_the_middle_of_a_plan(n):
    """Planner mid_relabel: dense ladder sweeps store relabelled in the middle of a batch -- into the
    contiguous low block, with the next targets rotated into the coalescing positions -- and the layout is restored by the
    last sweep.  Forced on at every size here (the default starts at 2^24 amplitudes); against the in-place plan, the
    oracle / the closed form"""
    a, coefs = _dense_qft(n, {"mid_relabel": 2})
    b, _ = _dense_qft(n, {"mid_relabel": 0})
    sa, sb = a.stats(), b.stats()
    assert sa["fallback_sweeps"] == 0 and sa["sweeps"] <= sb["sweeps"]
    assert sa["fused_relabels"] > sb["fused_relabels"]                   # at least one store in the middle relabels
    ca, cb = a.column(0), b.column(0)
    assert rel_l2(ca, cb) < 1e-12
    assert rel_l2(ca, W.qft_of_product_state(n, coefs, np.arange(1 << n))) < 1e-10
    assert abs(a.column_totals()[0] - 1.0) < 1e-12
    a.close(); b.close()


def test_cfg3_qft30_dense_input_closed_form():
    """BASELINE cfg3 at full size on a DENSE input (seeded product state, from_qubit_coefs): the QFT of a product state
    has a per-amplitude closed form (workloads.qft_of_product_state, pinned against the oracle in
    tests/test_oracle_crosscheck.py); 2^16-amplitude windows and the norm are checked at 1e-10"""
    n = 30
    st, coefs = _dense_qft(n, {})
    s = st.stats()
    assert s["fallback_sweeps"] == 0 and s["sweeps"] <= 4
    N = 1 << n
    rs = np.random.default_rng(30)
    for off in [0, N - 65536, N // 2 - 32768] + [int(v) & ~4095 for v in rs.integers(0, N - 65536, size=5)]:
        want = W.qft_of_product_state(n, coefs, np.arange(off, off + 65536, dtype=np.int64))
        assert rel_l2(st.column(0, off, 65536), want) < 1e-10
    assert abs(st.column_totals()[0] - 1.0) < 1e-12
    st.close()


def test_cfg4_ghz24_branching_full_size():
    """BASELINE cfg4 at n = 24 (circuit.rs:669-688 measurement sandwiches, vectorstate.rs:193-227 conditional column
    splitting): classical register, column structure and histogram bit-exact against the oracle circuit"""
    n = 24
    ops = W.ghz_branching_ops(n)
    words = O.splitmix64_words(9, 6 * 1024 + 64)
    c = QC.Circuit(n, n)
    W.load_ops(c, ops)
    c.execute(1024, E.Rng(words=words))
    o = O.OracleCircuit(n, n, mode=1, order=1)
    W.load_ops(o, ops)
    O.lib().orc_set_threads(8)
    o.execute(1024, O.Rng(words=words))
    O.lib().orc_set_threads(1)
    assert np.array_equal(c.cstate(), o.c_state)
    assert c.histogram_u64() == o.histogram()


def test_measure_all_bit_exact_n24():
    """measure_all of 8192 shots at n = 24 on identical amplitude bits (from_qubit_coefs is bit-exact between engine
    and oracle): canonical leaf/block scan + draw resolution give the oracle's outcomes word for word"""
    n, shots = 24, 8192
    coefs = W.product_state_coefs(n, seed=24)
    e, o = E.VectorState.from_qubit_coefs(coefs, shots), O.OracleState.from_qubit_coefs(coefs, shots)
    words = O.splitmix64_words(4, shots + 8)
    re_, ro = np.zeros(shots, dtype=np.uint64), np.zeros(shots, dtype=np.uint64)
    e.measure_all_into(list(range(n)), re_, E.Rng(words=words))
    o.measure_all_into(list(range(n)), ro, O.Rng(words=words))
    assert np.array_equal(re_, ro)


@pytest.mark.parametrize("n,depth", [(12, 40), (20, 100)])
def test_launch_bound_batches_replayed_as_cuda_graphs(n, depth):
    """cfg2 shape (random H/U3/CX/CS/CT layers): a hundred short sweeps per run.  From the second run of the same gate list
    on, the whole batch is ONE cudaGraphLaunch of the captured sweeps; results are identical to issuing them one by one"""
    import gc
    gc.collect()                 # (graphs are off while another state is alive on the device: let earlier tests' states go)
    ops = W.random_circuit_ops(n, depth, measure=False)
    gates = [(E.gate_matrix(o[1], o[2]), o[3], o[1]) for o in ops]
    cols = {}
    for graphs in (1, 0):
        st = E.VectorState(n, 16)
        st.set_option("graphs", graphs)
        for rep in range(3):
            st.reset_all()
            for m, b, name in gates:
                st.apply_gate(m, b, name)
            st.flush()
        s = st.stats()
        if graphs:
            assert s["graph_captures"] + s["graph_replays"] >= 3 and s["graph_replays"] >= 1, s
        else:
            assert s["graph_captures"] == 0 and s["graph_replays"] == 0
        cols[graphs] = st.column(0)
        st.close()
    assert np.array_equal(cols[0], cols[1])
    if n <= 12:
        o = O.OracleState(n, 16, mode=1, order=1)
        for m, b, _ in gates:
            o.apply_gate(m, b)
        assert rel_l2(cols[1], o.column(0)) < 1e-10
