"""GPU parity tests: the CUDA engine (through the C ABI of include/q1t_engine.h)
against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): amplitudes within 1e-10 relative L2 in f64;
measurement outcomes, counts and column structure BIT-EXACT when both sides
consume the same injected sequence of u64 words."""
import math

import numpy as np
import pytest

from oracle import oracle as O
from q1tsim_b200 import engine as E
from q1tsim_b200 import workloads as W
from tests import np_ref

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def rand_state(n, seed):
    r = np.random.default_rng(seed)
    v = r.normal(size=1 << n) + 1j * r.normal(size=1 << n)
    return v / np.linalg.norm(v)


def rand_unitary(k, seed):
    r = np.random.default_rng(seed)
    a = r.normal(size=(1 << k, 1 << k)) + 1j * r.normal(size=(1 << k, 1 << k))
    q, _ = np.linalg.qr(a)
    return q


def pair(n, shots=1, seed=None):
    """engine + oracle states holding the same (random) column"""
    e, o = E.VectorState(n, shots), O.OracleState(n, shots, mode=1, order=1)
    if seed is not None:
        psi = rand_state(n, seed)
        e.set_column(0, psi)
        o.set_column(0, psi)
    return e, o


GATE_POOL = [("h", 0), ("x", 0), ("y", 0), ("z", 0), ("s", 0), ("sdg", 0), ("t", 0), ("tdg", 0), ("v", 0), ("vdg", 0),
             ("i", 0), ("rx", 1), ("ry", 1), ("rz", 1), ("u1", 1), ("u2", 2), ("u3", 3), ("cx", 0), ("cy", 0), ("cz", 0),
             ("ch", 0), ("cs", 0), ("csdg", 0), ("ct", 0), ("ctdg", 0), ("cv", 0), ("cvdg", 0), ("swap", 0), ("crx", 1),
             ("cry", 1), ("crz", 1), ("cu1", 1), ("cu2", 2), ("cu3", 3), ("ccx", 0), ("ccz", 0), ("ccrx", 1), ("ccry", 1),
             ("ccrz", 1)]


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 9, 12, 13, 16])
def test_every_builtin_gate_random_bits(n):
    rs = np.random.default_rng(1000 + n)
    e, o = pair(n, seed=n)
    for rep in range(3):
        for name, npar in GATE_POOL:
            m = O.gate_matrix(name, rs.uniform(-3, 3, size=npar))
            k = int(round(math.log2(m.shape[0])))
            if k > n:
                continue
            bits = [int(b) for b in rs.permutation(n)[:k]]
            e.apply_gate(m, bits, name)
            o.apply_gate(m, bits)
    assert rel_l2(e.column(0), o.column(0)) < TOL


@pytest.mark.parametrize("n,k", [(5, 2), (8, 2), (8, 3), (12, 2), (12, 4), (14, 3), (14, 5), (12, 6), (13, 7), (14, 8), (12, 10)])
def test_dense_user_gates(n, k):
    """arbitrary user gates supply only matrix() (lib.rs:148-196)"""
    rs = np.random.default_rng(7 * n + k)
    e, o = pair(n, seed=3 * n)
    for rep in range(5):
        u = rand_unitary(k, rep + n)
        bits = [int(b) for b in rs.permutation(n)[:k]]
        e.apply_gate(u, bits, "user")
        o.apply_gate(u, bits)
        h = O.gate_matrix("h")
        b = [int(rs.integers(n))]
        e.apply_gate(h, b, "H")
        o.apply_gate(h, b)
    assert rel_l2(e.column(0), o.column(0)) < TOL


@pytest.mark.parametrize("n,k,nc", [(12, 3, 1), (12, 5, 2), (13, 6, 1), (13, 7, 2)])
def test_controlled_dense_user_gates(n, k, nc):
    """a user matrix that is the identity unless its first nc qubits are 1 (gates.rs:310-325 applies it as any dense
    matrix): the engine extracts the controls and runs the k-target block (registers for k <= 5, staged groups above)"""
    rs = np.random.default_rng(11 * n + k)
    e, o = pair(n, seed=5 * n + k)
    for rep in range(3):
        u = rand_unitary(k, 3 * rep + n)
        dim = 1 << (k + nc)
        m = np.eye(dim, dtype=np.complex128)
        m[dim - (1 << k):, dim - (1 << k):] = u
        bits = [int(b) for b in rs.permutation(n)[:k + nc]]
        e.apply_gate(m, bits, "ctl-user")
        o.apply_gate(m, bits)
        h = O.gate_matrix("h")
        for b in bits[:nc]:
            e.apply_gate(h, [b], "H")
            o.apply_gate(h, [b])
    assert rel_l2(e.column(0), o.column(0)) < TOL
    assert e.stats()["fallback_sweeps"] >= 3


def test_reference_state_kats():
    """vectorstate.rs:640-707 on the engine (tolerance of the reference: 1e-15 abs)"""
    z, o_, x = 0j, 1 + 0j, complex(math.sqrt(0.5))
    G = O.gate_matrix
    s = E.VectorState(3, 1)
    s.apply_gate(G("h"), [0])
    assert np.allclose(s.column(0), [x, z, z, z, x, z, z, z], atol=1e-15, rtol=0)
    s = E.VectorState(3, 1)
    s.apply_gate(G("y"), [2])
    assert np.allclose(s.column(0), [z, 1j, z, z, z, z, z, z], atol=1e-15, rtol=0)
    s = E.VectorState.from_qubit_coefs([z, o_, o_, z, o_, z], 1)
    s.apply_gate(G("cx"), [0, 2])
    assert np.allclose(s.column(0), [z, z, z, z, z, o_, z, z], atol=1e-15, rtol=0)
    s = E.VectorState.from_qubit_coefs([z, o_, z, o_, o_, z], 1)
    s.apply_gate(G("ccx"), [0, 2, 1])
    assert np.allclose(s.column(0), [z, z, z, z, z, z, o_, z], atol=1e-15, rtol=0)
    hx = 0.5 * x
    s = E.VectorState.from_qubit_coefs([x, -x, x, -x, x, -x], 1)
    s.apply_gate(G("ccx"), [0, 2, 1])
    assert np.allclose(s.column(0), [hx, -hx, -hx, hx, -hx, -hx, hx, hx], atol=1e-15, rtol=0)
    s = E.VectorState.from_qubit_coefs([z, o_, o_, z, o_, z], 1)
    s.apply_gate(np.kron(G("h"), G("h")), [1, 2])
    assert np.allclose(s.column(0), [z, z, z, z, .5, .5, .5, .5], atol=1e-15, rtol=0)


@pytest.mark.parametrize("n", [1, 2, 5, 11, 17])
def test_from_qubit_coefs_bit_exact(n):
    rs = np.random.default_rng(n)
    coefs = rs.normal(size=2 * n) + 1j * rs.normal(size=2 * n)
    e, o = E.VectorState.from_qubit_coefs(coefs, 3), O.OracleState.from_qubit_coefs(coefs, 3)
    assert np.array_equal(e.column(0), o.column(0))


@pytest.mark.parametrize("n", [5, 8, 11, 12, 13, 16, 20, 22])
def test_qft_vs_oracle(n):
    prep = W.u3_layer_ops(n, seed=1)
    ops = prep + W.qft_ops(n, measure=False)
    e, o = pair(n)
    for op in ops:
        m = O.gate_matrix(op[1], op[2])
        e.apply_gate(m, op[3], op[1])
        o.apply_gate(m, op[3])
    assert rel_l2(e.column(0), o.column(0)) < TOL
    st = e.stats()
    if n >= 12:
        assert st["sweeps"] <= 2 + (n + 8) // 9 + 1     # fused: a handful of sweeps, not one per gate
        assert st["fallback_sweeps"] == 0


@pytest.mark.parametrize("n", [24, 26])
def test_qft_closed_form_large(n):
    """size-independent check: QFT|x> = 2^(-n/2) exp(2 pi i rev(x) rev(y) / 2^n)  """
    x = (0b1011 << (n - 5)) | 0b101
    e = E.VectorState(n, 1)
    for q in range(n):
        if (x >> (n - 1 - q)) & 1:
            e.apply_gate(O.gate_matrix("x"), [q], "X")
    for op in W.qft_ops(n, measure=False):
        e.apply_gate(O.gate_matrix(op[1], op[2]), op[3], op[1])
    rs = np.random.default_rng(5)
    N = 1 << n

    def rev(v):
        return int(format(v, "0%db" % n)[::-1], 2)
    for off in [0, N - 4096] + [int(v) for v in rs.integers(0, N - 4096, size=6)]:
        got = e.column(0, off, 4096)
        ys = np.array([rev(y) for y in range(off, off + 4096)], dtype=object)
        ph = np.array([((rev(x) * int(y)) % N) / N for y in ys], dtype=np.float64)
        want = np.exp(2j * np.pi * ph) / math.sqrt(N)
        assert rel_l2(got, want) < 1e-9
    assert abs(e.column_totals()[0] - 1.0) < 1e-12


def test_random_circuit_cfg2_small_and_full():
    for n, depth in ((10, 20), (20, 100)):
        ops = W.random_circuit_ops(n, depth, measure=False)
        e, o = pair(n)
        O.lib().orc_set_threads(8)
        for op in ops:
            m = O.gate_matrix(op[1], op[2])
            e.apply_gate(m, op[3], op[1])
            o.apply_gate(m, op[3])
        O.lib().orc_set_threads(1)
        assert rel_l2(e.column(0), o.column(0)) < TOL


@pytest.mark.parametrize("n", [1, 3, 4, 7, 10, 11, 14, 18])
def test_canonical_reductions_bit_exact(n):
    e, o = pair(n, seed=40 + n)
    for q in sorted({0, n // 2, n - 1}):
        assert e.marginal0(q)[0] == o.marginal0(q, order=1)[0]
    assert e.column_totals()[0] == o.column_totals(order=1)[0]


@pytest.mark.parametrize("n,shots", [(1, 16), (3, 100), (6, 1000), (10, 1024), (12, 5000), (16, 8192)])
def test_measure_all_bit_exact(n, shots):
    words = O.splitmix64_words(2, shots + 8)
    psi = rand_state(n, 90 + n)
    for collapse in (False, True):
        e, o = pair(n, shots)
        e.set_column(0, psi); o.set_column(0, psi)
        cbits = list(np.random.default_rng(n).permutation(n))
        re_, ro = np.full(shots, 1 << 40, dtype=np.uint64), np.full(shots, 1 << 40, dtype=np.uint64)
        rng_e, rng_o = E.Rng(words=words), O.Rng(words=words)
        if collapse:
            e.measure_all_into(cbits, re_, rng_e); o.measure_all_into(cbits, ro, rng_o)
        else:
            e.peek_all_into(cbits, re_, rng_e); o.peek_all_into(cbits, ro, rng_o)
        assert np.array_equal(re_, ro)
        assert rng_e.consumed == rng_o.consumed == shots
        assert e.counts == o.counts
        if n <= 10:
            assert np.array_equal(e.states(), o.states())


@pytest.mark.parametrize("n,shots", [(2, 64), (5, 1024), (10, 1000), (14, 4096)])
def test_measure_peek_sequence_bit_exact(n, shots):
    """mid-circuit Z measurements: same uniforms -> same results, same column structure"""
    words = O.splitmix64_words(3, 4096)
    rng_e, rng_o = E.Rng(words=words), O.Rng(words=words)
    e, o = pair(n, shots, seed=7 + n)
    ce, co = np.zeros(shots, dtype=np.uint64), np.zeros(shots, dtype=np.uint64)
    qs = [0, n - 1, n // 2]
    for i, q in enumerate(qs):
        e.peek_into(q, 10 + i, ce, rng_e); o.peek_into(q, 10 + i, co, rng_o)
        e.measure_into(q, i, ce, rng_e); o.measure_into(q, i, co, rng_o)
        h = O.gate_matrix("h")
        e.apply_gate(h, [(q + 1) % n], "H"); o.apply_gate(h, [(q + 1) % n])
    assert np.array_equal(ce, co)
    assert rng_e.consumed == rng_o.consumed
    assert e.counts == o.counts
    se, so = e.states(), o.states()
    for c in range(o.ncols):
        assert rel_l2(se[:, c], so[:, c]) < TOL
    # w0 of the collapsed columns is exactly reproducible too
    assert np.array_equal(e.marginal0(0), o.marginal0(0, order=1)) or np.allclose(e.marginal0(0), o.marginal0(0, order=1), atol=1e-15)


def test_conditional_gate_column_splitting():
    """vectorstate.rs:472-509 on the engine"""
    z, o_, x = 0j, 1 + 0j, complex(math.sqrt(0.5))
    G = O.gate_matrix
    s = E.VectorState(2, 5)
    s.apply_conditional_gate([0, 0, 1, 1, 0], G("x"), [1], "X")
    assert s.counts == [2, 2, 1]
    assert np.allclose(s.states(), [[o_, z, o_], [z, o_, z], [z, z, z], [z, z, z]], atol=1e-15, rtol=0)
    s = E.VectorState(2, 5)
    s.apply_conditional_gate([1, 0, 1, 1, 0], G("h"), [1], "H")
    assert s.counts == [1, 1, 2, 1]
    assert np.allclose(s.states(), [[x, o_, x, o_], [x, z, x, z], [z, z, z, z], [z, z, z, z]], atol=1e-15, rtol=0)
    s = E.VectorState.from_qubit_coefs([o_, z, x, x], 5)
    s.apply_conditional_gate([1, 0, 1, 1, 0], G("cx"), [1, 0], "CX")
    assert s.counts == [1, 1, 2, 1]
    assert np.allclose(s.states(), [[x, x, x, x], [z, x, z, x], [z, z, z, z], [x, z, x, z]], atol=1e-15, rtol=0)
    s = E.VectorState(2, 5)
    s.apply_conditional_gate([1, 1, 1, 0, 0], G("h"), [0], "H")
    assert s.counts == [3, 2]
    s.apply_conditional_gate([0, 0, 1, 1, 1], G("h"), [0], "H")
    assert s.counts == [2, 1, 2]
    assert np.allclose(s.states(), [[x, o_, x], [z, z, z], [x, z, x], [z, z, z]], atol=1e-15, rtol=0)


@pytest.mark.parametrize("n", [2, 6, 12])
def test_reset_and_reset_all(n):
    words = O.splitmix64_words(11, 256)
    shots = 50
    e, o = pair(n, shots, seed=n)
    rng_e, rng_o = E.Rng(words=words), O.Rng(words=words)
    e.reset(n - 1, rng_e); o.reset(n - 1, rng_o)
    e.reset(0, rng_e); o.reset(0, rng_o)
    assert e.counts == o.counts and rng_e.consumed == rng_o.consumed
    se, so = e.states(), o.states()
    for c in range(o.ncols):
        assert rel_l2(se[:, c], so[:, c]) < TOL
    e.reset_all()
    assert e.counts == [shots]
    want = np.zeros(1 << n, dtype=np.complex128)
    want[0] = 1
    assert np.array_equal(e.column(0), want)


def test_full_circuit_cfg1_readme_qft3():
    """README.md:59-68 through the op interpreter: uniform amplitudes, ~1024 per outcome"""
    from tests.test_oracle_reference_kats import measurement_ok
    words = O.splitmix64_words(42, 8192)
    out = []
    for backend in ("engine", "oracle"):
        c = O.OracleCircuit(3, 3)
        W.load_ops(c, W.qft_ops(3, measure=False) + [("peek_all", [0, 1, 2], "Z")])
        if backend == "engine":
            c.execute(8192, E.Rng(words=words), q_state=E.VectorState(3, 8192))
        else:
            c.execute(8192, O.Rng(words=words))
        assert np.allclose(c.q_state.column(0), np.full(8, 1 / math.sqrt(8)), atol=1e-15)
        out.append(c.c_state.copy())
    assert np.array_equal(out[0], out[1])
    hv = np.bincount(out[0].astype(np.int64), minlength=8)
    assert all(measurement_ok(int(v), 8192, 0.125, 1e-5) for v in hv)


@pytest.mark.parametrize("n", [6, 12, 16])
def test_full_circuit_cfg4_ghz_branching(n):
    """GHZ + X/Y/Z-basis mid-circuit measurements + conditional gates (multi-column branching)"""
    shots = 1024
    words = O.splitmix64_words(5, 3 * shots)
    res = []
    for backend in ("engine", "oracle"):
        c = O.OracleCircuit(n, n)
        W.load_ops(c, W.ghz_branching_ops(n)[:-1] + [("peek_all", list(range(n)), "Z")])
        if backend == "engine":
            c.execute(shots, E.Rng(words=words), q_state=E.VectorState(n, shots))
        else:
            c.execute(shots, O.Rng(words=words))
        res.append((c.c_state.copy(), c.q_state.counts, c.q_state.states()))
    assert np.array_equal(res[0][0], res[1][0])
    assert res[0][1] == res[1][1]
    for col in range(len(res[1][1])):
        assert rel_l2(res[0][2][:, col], res[1][2][:, col]) < TOL


def test_error_behaviour():
    """same variants and Display text as error.rs:192-255"""
    s = E.VectorState(2, 4)
    with pytest.raises(E.EngineError) as ei:
        s.apply_gate(O.gate_matrix("cx"), [0], "CX")
    assert ei.value.kind == "InvalidNrBits" and str(ei.value) == 'Expected 2 bits for "CX", got 1'
    with pytest.raises(E.EngineError) as ei:
        s.measure_into(5, 0, np.zeros(4, dtype=np.uint64), E.Rng(seed=1))
    assert ei.value.kind == "InvalidQBit" and str(ei.value) == "Invalid index 5 for a quantum bit"
    with pytest.raises(E.EngineError) as ei:
        s.measure_into(0, 0, np.zeros(2, dtype=np.uint64), E.Rng(seed=1))
    assert ei.value.kind == "NotEnoughSpace"
    assert str(ei.value) == "Not enough space to store 4 measurement results in array of length 2"
    with pytest.raises(E.EngineError) as ei:
        s.measure_all_into([0], np.zeros(4, dtype=np.uint64), E.Rng(seed=1))
    assert ei.value.kind == "InvalidNrMeasurementBits" and str(ei.value) == "Expected 2 measurement bits, but got 1"
    with pytest.raises(E.EngineError) as ei:
        s.apply_conditional_gate([1, 0], O.gate_matrix("x"), [0], "X")
    assert ei.value.kind == "InvalidNrControlBits"
    assert str(ei.value) == "The number of runs is 4, but received 2 control bits for controlled X operation"
    with pytest.raises(E.EngineError) as ei:
        s.apply_gate(O.gate_matrix("cx"), [1, 1], "CX")      # rejected instead of silently wrong
    assert ei.value.kind == "InvalidArgument"
    # |00>: w0 == 1 exactly, Binomial::sample returns without touching the generator
    s.measure_into(0, 0, np.zeros(4, dtype=np.uint64), E.Rng(words=[]))
    s2 = E.VectorState.from_qubit_coefs([1, 1, 1, 0], 4)
    with pytest.raises(E.EngineError) as ei:
        s2.measure_into(0, 0, np.zeros(4, dtype=np.uint64), E.Rng(words=[]))
    assert ei.value.kind == "RngExhausted"


def test_unfused_path_matches_fused():
    n = 12
    ops = W.u3_layer_ops(n) + W.qft_ops(n, measure=False)
    outs = []
    for fuse in (1, 0):
        e = E.VectorState(n, 1)
        e.set_option("fuse", fuse)
        for op in ops:
            e.apply_gate(O.gate_matrix(op[1], op[2]), op[3], op[1])
        outs.append(e.column(0))
        st = e.stats()
        assert (st["fallback_sweeps"] == 0) == bool(fuse)
    assert rel_l2(outs[0], outs[1]) < TOL


@pytest.mark.parametrize("tile_bits", [8, 10, 11, 13])
def test_tile_sizes(tile_bits):
    n = 15
    ops = W.u3_layer_ops(n) + W.qft_ops(n, measure=False) + W.random_circuit_ops(n, 6, measure=False)
    e, o = pair(n)
    e.set_option("tile_bits", tile_bits)
    for op in ops:
        m = O.gate_matrix(op[1], op[2])
        e.apply_gate(m, op[3], op[1]); o.apply_gate(m, op[3])
    assert rel_l2(e.column(0), o.column(0)) < TOL


def _ladder_ops(n, targets, partners_per_target=None):
    """H on `targets` (in order), each preceded by controlled phases from every earlier-touched or
    arbitrary other qubit: a QFT-like ladder restricted to a subset of the qubits."""
    ops = []
    done = []
    for t in targets:
        for d, c in enumerate(done[::-1][:6]):
            ops.append(("gate", "cu1", (math.pi / (1 << (d + 1)),), [c, t]))
        for c in (partners_per_target or {}).get(t, []):
            ops.append(("gate", "cu1", (0.37 * (c + 1),), [c, t]))
        ops.append(("gate", "h", (), [t]))
        done.append(t)
    return ops


@pytest.mark.parametrize("n", [10, 12, 13, 15, 17])
@pytest.mark.parametrize("case", ["qft", "qft_noswap", "low_half", "high_half", "scattered", "one"])
def test_support_tracking_from_basis_states(n, case):
    """Circuits that start from |0..0> and consist of (phase, Hadamard) ladders run with support
    tracking: only what can be non-zero is read, computed and written.  Same result
    as with tracking off and as the oracle, including the tiles that are written as zeros."""
    if case == "qft":
        ops = W.qft_ops(n, measure=False)
    elif case == "qft_noswap":
        ops = W.qft_ops(n, measure=False, swaps=False)
    elif case == "low_half":
        ops = _ladder_ops(n, list(range(n - 1, n // 2, -1)))
    elif case == "high_half":
        ops = _ladder_ops(n, list(range(n // 2, -1, -1)), {0: [n - 1], 1: [n - 2]})
    elif case == "scattered":
        ops = _ladder_ops(n, [n - 1, 0, n // 2, 3, n - 4, 1], {3: [2, n - 2]})
    else:
        ops = _ladder_ops(n, [n // 2])
    outs = []
    for track in (1, 0):
        e = E.VectorState(n, 1)
        e.set_option("track_support", track)
        for op in ops:
            e.apply_gate(O.gate_matrix(op[1], op[2]), op[3], op[1])
        outs.append(e.column(0))
        assert e.stats()["fallback_sweeps"] == 0
        e.close()
    o = O.OracleState(n, 1, mode=1, order=1)
    for op in ops:
        o.apply_gate(O.gate_matrix(op[1], op[2]), op[3])
    assert rel_l2(outs[0], o.column(0)) < TOL
    assert rel_l2(outs[1], o.column(0)) < TOL
    assert rel_l2(outs[0], outs[1]) < 1e-14


@pytest.mark.parametrize("n,shots", [(10, 16), (13, 40)])
def test_support_tracking_collapsed_columns(n, shots):
    """measure_all leaves one lazy basis column per distinct outcome (vectorstate.rs:150-158): a ladder
    applied afterwards starts from a different basis index in every column."""
    e, o = E.VectorState(n, shots), O.OracleState(n, shots, mode=1, order=1)
    for q in range(n):
        m = O.gate_matrix("h")
        e.apply_gate(m, [q], "H"); o.apply_gate(m, [q])
    words = O.splitmix64_words(11, shots)
    re_, ro = np.zeros(shots, dtype=np.uint64), np.zeros(shots, dtype=np.uint64)
    e.measure_all_into(list(range(n)), re_, E.Rng(words=words))
    o.measure_all_into(list(range(n)), ro, O.Rng(words=words))
    assert np.array_equal(re_, ro)
    for op in W.qft_ops(n, measure=False):
        m = O.gate_matrix(op[1], op[2])
        e.apply_gate(m, op[3], op[1]); o.apply_gate(m, op[3])
    assert e.counts == o.counts
    for c in range(len(o.counts)):
        assert rel_l2(e.column(c), o.column(c)) < TOL


def test_conditional_swap_moves_only_flagged_columns():
    """A conditional Swap (vectorstate.rs:193-227 with swap.rs) must not relabel the unflagged columns."""
    n, shots = 6, 8
    e, o = pair(n, shots, seed=3)
    words = O.splitmix64_words(5, 64)
    re_, ro = np.zeros(shots, dtype=np.uint64), np.zeros(shots, dtype=np.uint64)
    e.measure_into(2, 0, re_, E.Rng(words=words)); o.measure_into(2, 0, ro, O.Rng(words=words))
    assert np.array_equal(re_, ro)
    ctrl = [bool(v & 1) for v in re_]
    m = O.gate_matrix("swap")
    e.apply_conditional_gate(ctrl, m, [0, 5], "Swap"); o.apply_conditional_gate(ctrl, m, [0, 5])
    mh = O.gate_matrix("h")
    e.apply_gate(mh, [0], "H"); o.apply_gate(mh, [0])
    assert e.counts == o.counts
    for c in range(len(o.counts)):
        assert rel_l2(e.column(c), o.column(c)) < TOL


@pytest.mark.parametrize("n,prep", [(12, "zero"), (14, "zero"), (17, "zero"), (20, "zero"), (12, "collapsed"), (14, "collapsed"),
                                    (13, "dense")])
def test_measure_all_after_qft_fused_leaf_totals(n, prep):
    """measure_all right after a QFT: when the last sweep's tile holds whole canonical leaves its store pass
    also produces the leaf totals (no separate read pass).  Outcomes must stay bit-exact (canonical order,
    DESIGN.md 4.2) against the oracle and against the unfused path."""
    shots = 2000
    outs = []
    for fuse in (1, 0):
        e, o = E.VectorState(n, shots), O.OracleState(n, shots, mode=1, order=1)
        e.set_option("fuse_leaf_totals", fuse)
        words = O.splitmix64_words(21, 3 * shots + 64)
        re_, ro = np.zeros(shots, dtype=np.uint64), np.zeros(shots, dtype=np.uint64)
        rng_e, rng_o = E.Rng(words=words), O.Rng(words=words)
        if prep == "collapsed":
            for q in range(0, n, 2):
                m = O.gate_matrix("h"); e.apply_gate(m, [q], "H"); o.apply_gate(m, [q])
            e.measure_all_into(list(range(n)), re_, rng_e); o.measure_all_into(list(range(n)), ro, rng_o)
            assert np.array_equal(re_, ro)
        elif prep == "dense":
            for op in W.u3_layer_ops(n, seed=4):
                m = O.gate_matrix(op[1], op[2]); e.apply_gate(m, op[3], op[1]); o.apply_gate(m, op[3])
            e.flush()
        e.reset_stats()
        for op in W.qft_ops(n, measure=False):
            m = O.gate_matrix(op[1], op[2]); e.apply_gate(m, op[3], op[1]); o.apply_gate(m, op[3])
        e.measure_all_into(list(range(n)), re_, rng_e); o.measure_all_into(list(range(n)), ro, rng_o)
        assert np.array_equal(re_, ro), "outcomes differ from the oracle (fuse=%d)" % fuse
        assert e.counts == o.counts
        st = e.stats()
        if prep == "zero" and n in (12, 20):
            # QFT-12 is one tile, QFT-20 two sweeps of 10 bits: the last tile holds whole leaves of 1024
            assert st["read_passes"] == (0 if fuse else 1)       # fused: no leaf_totals launch at all
        outs.append(re_.copy())
        e.close()
    assert np.array_equal(outs[0], outs[1])

