"""The sharded state behind the C ABI (csrc/sharded.h: q1t_sharded_* of include/q1t_engine.h, circuit_set_devices of
include/q1tsim_ffi.h): one process, P shards.  On a one-GPU box the P shards all live on device 0 -- same kernels, same
peer-group mechanism (mailbox barriers, multi-bit swap), plain pointers instead of IPC mappings; on a multi-GPU box
`Q1T_TEST_DEVICES=0,1,2,3` spreads them."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from q1tsim_b200 import circuit as QC
from q1tsim_b200 import engine as E
from q1tsim_b200 import workloads as W

pytestmark = pytest.mark.gpu

# Every test body runs in a process of its own (a pytest child on this very test): P shards of ONE process on ONE device
# are P streams whose barrier kernels wait for each other, so they need as many hardware queues
# (CUDA_DEVICE_MAX_CONNECTIONS, read when the context is created), and a kernel that trapped there must not take the
# CUDA context of the other GPU tests with it.
INNER = os.environ.get("Q1T_SHARDED_CABI_INNER") == "1"


def in_own_process(fn):
    import inspect
    import subprocess
    import sys

    def wrapper(*args, **kw):
        request = kw.pop("request")
        if INNER:
            return fn(*args, **kw)
        env = dict(os.environ, Q1T_SHARDED_CABI_INNER="1", CUDA_DEVICE_MAX_CONNECTIONS="32")
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "%s::%s" % (os.path.abspath(__file__), request.node.name)],
                           cwd=root, env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]

    sig = inspect.signature(fn)
    params = list(sig.parameters.values()) + [inspect.Parameter("request", inspect.Parameter.KEYWORD_ONLY)]
    wrapper.__signature__ = sig.replace(parameters=params)
    wrapper.__name__, wrapper.__doc__, wrapper.__module__ = fn.__name__, fn.__doc__, fn.__module__
    return wrapper


def _devices(P):
    env = os.environ.get("Q1T_TEST_DEVICES")
    if env:
        ds = [int(x) for x in env.split(",")]
        return [ds[i % len(ds)] for i in range(P)]
    return [0] * P


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("P,n", [(2, 14), (4, 15), (8, 16)])
@in_own_process
def test_random_gates_on_every_qubit(P, n):
    """pinned and free rank bits, rank-selected blocks, remaps (single and multi-bit), relabels: against the oracle"""
    st = E.ShardedProcessState(n, 64, _devices(P))
    ref = O.OracleState(n, 64, mode=1, order=1)
    rs = np.random.default_rng(P * 100 + n)
    names = [("h", 0), ("x", 0), ("u3", 3), ("cx", 0), ("cu1", 1), ("cs", 0), ("swap", 0), ("rz", 1), ("ccx", 0), ("y", 0), ("t", 0), ("cz", 0), ("crx", 1)]
    for rep in range(40):
        name, npar = names[rep % len(names)]
        m = O.gate_matrix(name, list(rs.uniform(-2, 2, size=npar)))
        k = int(np.log2(m.shape[0]))
        bits = [int(b) for b in rs.permutation(n)[:k]]
        st.apply_gate(m, bits, name)
        ref.apply_gate(m, bits)
    assert rel_l2(st.amplitudes(), ref.column(0)) < 1e-10
    assert abs(st.column_total() - 1.0) < 1e-12
    assert st.counters()["remaps"] >= 1
    st.close()


@pytest.mark.parametrize("P,n", [(2, 13), (4, 16), (8, 17), (4, 22)])
@in_own_process
def test_qft_circuit_through_the_ffi_abi(P, n):
    """BASELINE config 5 at test size through circuit_set_devices + circuit_execute: QFT + measure_all; the run starts
    from the layout its Swap relabels turn into the canonical one, so it costs no exchange at all"""
    ops = W.qft_ops(n, measure=False) + [("peek_all", list(range(n)), "Z")]
    shots = 500
    words = O.splitmix64_words(5, 2 * shots + 64)
    c = QC.Circuit(n, n)
    c.set_devices(_devices(P))
    W.load_ops(c, ops)
    c.execute(shots, E.Rng(words=words))
    o = O.OracleCircuit(n, n, mode=1, order=1)
    W.load_ops(o, ops)
    O.lib().orc_set_threads(8)
    o.execute(shots, O.Rng(words=words))
    O.lib().orc_set_threads(1)
    assert c.sharded_counters()["remaps"] == 0
    assert rel_l2(c.sharded_amplitudes(), o.q_state.column(0)) < 1e-10
    assert np.array_equal(c.cstate(), o.c_state)
    # the collapsing measurement and a second execute() of the same circuit object
    ops2 = W.qft_ops(n, measure=True)
    c2 = QC.Circuit(n, n)
    c2.set_devices(_devices(P))
    W.load_ops(c2, ops2)
    o2 = O.OracleCircuit(n, n, mode=1, order=1)
    W.load_ops(o2, ops2)
    for rep in range(2):
        w = O.splitmix64_words(7 + rep, shots + 64)
        c2.execute(shots, E.Rng(words=w))
        o2.execute(shots, O.Rng(words=w))
        assert np.array_equal(c2.cstate(), o2.c_state)
        assert c2.histogram_u64() == o2.histogram()


@pytest.mark.parametrize("P,n", [(4, 15)])
@in_own_process
def test_basis_changes_and_a_dense_circuit(P, n):
    """X- and Y-basis measure_all (circuit.rs:705-735) and a circuit whose global qubits need a remap"""
    ops = W.u3_layer_ops(n, seed=3) + W.qft_ops(n, measure=False) + [("peek_all", list(range(n)), "X"), ("gate", "h", (), [0]),
                                                                      ("peek_all", list(range(n)), "Y"), ("gate", "t", (), [1]),
                                                                      ("measure_all", list(range(n)), "Z")]
    shots = 300
    words = O.splitmix64_words(11, 4 * shots + 64)
    c = QC.Circuit(n, n)
    c.set_devices(_devices(P))
    W.load_ops(c, ops)
    c.execute(shots, E.Rng(words=words))
    o = O.OracleCircuit(n, n, mode=1, order=1)
    W.load_ops(o, ops)
    o.execute(shots, O.Rng(words=words))
    assert np.array_equal(c.cstate(), o.c_state)
    assert c.sharded_counters()["remaps"] >= 1
    c3 = QC.Circuit(n, n)
    c3.set_devices(_devices(P))
    c3.add_gate("h", [0]); c3.measure_all_basis(list(range(n)), "X")
    with pytest.raises(QC.CircuitError) as e:
        c3.execute(4)
    assert "X or Y basis" in str(e.value)


@in_own_process
def test_errors_like_the_single_gpu_state():
    st = E.ShardedProcessState(12, 8, _devices(2))
    with pytest.raises(E.EngineError) as e:
        st.apply_gate(O.gate_matrix("h"), [12])
    assert "Invalid index 12 for a quantum bit" in str(e.value)
    with pytest.raises(E.EngineError) as e:
        st.apply_gate(O.gate_matrix("cx"), [1])
    assert "Expected 2 bits" in str(e.value)
    with pytest.raises(E.EngineError):
        E.ShardedProcessState(10, 8, _devices(2))           # 9 local qubits: below one canonical leaf
    c = QC.Circuit(12, 12)
    c.set_devices(_devices(2))
    c.add_gate("h", [0]); c.measure(0, 0)
    with pytest.raises(QC.CircuitError) as e2:
        c.execute(4)
    assert "sharded" in str(e2.value)
