"""GPU parity tests added late in round 1, kept in one file that sorts after the other GPU suites:
* the in-place relabelling path (planner.cpp plan_inplace_relabel, engine.cu canonicalize), forced
  at small sizes with the "inplace_relabel" option, against the CPU oracle;
* the reference's own gate known-answer vectors (tests/golden/gate_kats.json) through the engine."""
import numpy as np
import pytest

from oracle import oracle as O
from q1tsim_b200 import engine as E
from q1tsim_b200 import workloads as W
from tests.test_gpu_parity import TOL, pair, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,tile_bits", [(5, 12), (9, 12), (13, 12), (14, 12), (16, 10), (17, 12), (20, 12), (20, 13), (22, 12)])
def test_inplace_relabel_matches_out_of_place(n, tile_bits):
    """Shards that leave no room for a second column buffer restore canonical order by tile-closed passes with
    source == destination (planner.cpp plan_inplace_relabel) instead of one out-of-place sweep.  Forced here at
    small sizes: QFT with its swaps and a random swap network, amplitudes against the oracle, outcomes bit-exact."""
    shots = 500
    rs = np.random.default_rng(n)
    ops = W.u3_layer_ops(n, seed=2) + W.qft_ops(n, measure=False)
    for _ in range(2 * n):
        a, b = [int(v) for v in rs.permutation(n)[:2]]
        ops.append(("gate", "swap", (), [a, b]))
    ops += W.u3_layer_ops(n, seed=3)
    e, o = pair(n, shots)
    e.set_option("tile_bits", tile_bits)
    e.set_option("inplace_relabel", 1)
    for op in ops:
        m = O.gate_matrix(op[1], op[2])
        e.apply_gate(m, op[3], op[1]); o.apply_gate(m, op[3])
    e.flush()
    st = e.stats()
    assert st["fused_relabels"] == 0
    if n > 5:
        assert st["permute_sweeps"] >= 1
    assert rel_l2(e.column(0), o.column(0)) < TOL
    words = O.splitmix64_words(5, shots + 8)
    re_, ro = np.zeros(shots, dtype=np.uint64), np.zeros(shots, dtype=np.uint64)
    # a second relabelled batch that ends in measure_all (no fused leaf totals on this path)
    for op in W.qft_ops(n, measure=False):
        m = O.gate_matrix(op[1], op[2])
        e.apply_gate(m, op[3], op[1]); o.apply_gate(m, op[3])
    e.measure_all_into(list(range(n)), re_, E.Rng(words=words)); o.measure_all_into(list(range(n)), ro, O.Rng(words=words))
    assert np.array_equal(re_, ro)
    e.close()


def test_inplace_relabel_multi_column():
    """in-place passes run over every dense column of a branched state"""
    n, shots = 14, 64
    e, o = pair(n, shots)
    e.set_option("inplace_relabel", 1)
    words = O.splitmix64_words(9, 4 * shots)
    rng_e, rng_o = E.Rng(words=words), O.Rng(words=words)
    re_, ro = np.zeros(shots, dtype=np.uint64), np.zeros(shots, dtype=np.uint64)
    for q in (0, 5):
        m = O.gate_matrix("h"); e.apply_gate(m, [q], "H"); o.apply_gate(m, [q])
        e.measure_into(q, q, re_, rng_e); o.measure_into(q, q, ro, rng_o)
    assert np.array_equal(re_, ro) and e.counts == o.counts and len(e.counts) > 1
    for op in W.u3_layer_ops(n, seed=6) + W.qft_ops(n, measure=False):
        m = O.gate_matrix(op[1], op[2]); e.apply_gate(m, op[3], op[1]); o.apply_gate(m, op[3])
    e.flush()
    for c in range(len(o.counts)):
        assert rel_l2(e.column(c), o.column(c)) < TOL
    e.close()


def test_reference_gate_kats_on_the_engine():
    """the reference's own gate known-answer vectors (tests/golden/gate_kats.json, src/gates/*.rs unit tests)
    through the CUDA engine: the gate on the first k qubits of every column of `state`"""
    from tests import kat_fixtures as K
    for case in K.cases("apply"):
        state, want = K.carray(case["state"]), K.carray(case["result"])
        m = K.matrix_of(case["gate"], O.gate_matrix)
        k, n = int(np.log2(m.shape[0])), int(np.log2(state.shape[0]))
        for col in range(state.shape[1]):
            if not np.any(state[:, col]):
                continue
            e = E.VectorState(n, 1)
            e.set_column(0, state[:, col])
            e.apply_gate(m, list(range(k)), case["gate"]["name"])
            assert np.abs(e.column(0) - want[:, col]).max() <= 1e-12, (case["source"], col)
            e.close()
