"""The fusion planner checked on the CPU: the sweep programs it produces (q1t_plan_dump, the raw structs of
csrc/program.h) are executed by tests/plan_interpreter.py -- a numpy restatement of what sweep_kernel / ladder_kernel
do with them -- and compared with the CPU oracle applying the same gates one by one (1e-12; the GPU parity bar is
1e-10).  Covers every built-in gate (dense 2x2 with controls, all diagonal gates through the phase polynomial,
ladder rounds, linear-phase rounds), Swap relabelling undone the three ways the engine knows (fused into the last
sweep, a separate relabel sweep, the in-place passes), tile sizes 8..12, 64- and 128-byte coalescing, both packings.
Every address table the kernels read is checked against the logical layout it is derived from."""
import math

import numpy as np
import pytest

from oracle import oracle as O
from q1tsim_b200 import workloads as W
from tests import plan_interpreter as PI

TOL = 1e-12
GATE_POOL = [("h", 0), ("x", 0), ("y", 0), ("z", 0), ("s", 0), ("sdg", 0), ("t", 0), ("tdg", 0), ("v", 0), ("vdg", 0), ("i", 0),
             ("rx", 1), ("ry", 1), ("rz", 1), ("u1", 1), ("u2", 2), ("u3", 3), ("cx", 0), ("cy", 0), ("cz", 0), ("ch", 0), ("cs", 0),
             ("csdg", 0), ("ct", 0), ("ctdg", 0), ("cv", 0), ("cvdg", 0), ("swap", 0), ("crx", 1), ("cry", 1), ("crz", 1), ("cu1", 1),
             ("cu2", 2), ("cu3", 3), ("ccx", 0), ("ccz", 0), ("ccrx", 1), ("ccry", 1), ("ccrz", 1)]


def _random_ops(n, reps, seed):
    rs = np.random.default_rng(seed)
    ops = []
    for _ in range(reps):
        for name, npar in GATE_POOL:
            params = tuple(rs.uniform(-3, 3, size=npar))
            k = int(round(math.log2(O.gate_matrix(name, params).shape[0])))
            ops.append(("gate", name, params, [int(b) for b in rs.permutation(n)[:k]]))
    return ops


def _check(n, ops, tile_bits, coalesce=3, balanced=0, relabel_mode=0, seed=1):
    gates = [(O.gate_matrix(o[1], o[2]), o[3]) for o in ops if o[0] == "gate"]
    sweeps, perm = PI.plan(n, gates, tile_bits, coalesce, balanced | (relabel_mode << 4))
    if relabel_mode:
        assert perm == list(range(n))
    r = np.random.default_rng(seed)
    psi = r.normal(size=1 << n) + 1j * r.normal(size=1 << n)
    psi /= np.linalg.norm(psi)
    got = PI.run_plan(sweeps, perm, psi)
    o = O.OracleState(n, 1, mode=1, order=1)
    o.set_column(0, psi)
    for m, b in gates:
        o.apply_gate(m, b)
    assert np.linalg.norm(got - o.column(0)) < TOL
    for P, _ in sweeps:
        PI.check_tables(P)
    return sweeps


def test_struct_layout_matches_the_library():
    PI.check_layout()


@pytest.mark.parametrize("n,tile_bits,coalesce,balanced", [(9, 8, 3, 0), (10, 8, 2, 0), (11, 9, 3, 1), (12, 10, 3, 0), (13, 12, 3, 0),
                                                           (13, 11, 2, 1), (14, 12, 2, 0)])
def test_every_builtin_gate(n, tile_bits, coalesce, balanced):
    _check(n, _random_ops(n, 2, seed=n), tile_bits, coalesce, balanced, seed=n)


@pytest.mark.parametrize("n,tile_bits", [(8, 8), (10, 8), (12, 9), (13, 12)])
def test_qft_ladders(n, tile_bits):
    sweeps = _check(n, W.u3_layer_ops(n) + W.qft_ops(n, measure=False), tile_bits)
    # the QFT part runs as ladder rounds (ROUND_PH), not through the op interpreter
    assert any(P.rounds[r].kind == PI.ROUND_PH for P, _ in sweeps for r in range(P.nrounds))


def test_random_circuit_cfg2_shape():
    # SURVEY 8(d) cfg2 at reduced size: H / U3 layers and CX / CS / CT pair layers
    _check(12, W.random_circuit_ops(12, 20, measure=False), 10)


@pytest.mark.parametrize("n,tile_bits", [(8, 8), (10, 8), (12, 9), (13, 12), (14, 10)])
@pytest.mark.parametrize("relabel_mode", [1, 2])
def test_swap_relabelling_undone(n, tile_bits, relabel_mode):
    """Swap gates move no data (swap.rs:78-88 becomes a relabel); canonical order comes back fused into the last
    sweep / by one relabel sweep (mode 1) or by tile-closed in-place passes (mode 2)"""
    rs = np.random.default_rng(n)
    ops = W.u3_layer_ops(n) + W.qft_ops(n, measure=False)
    for _ in range(n):
        a, b = [int(v) for v in rs.permutation(n)[:2]]
        ops.append(("gate", "swap", (), [a, b]))
    ops += W.u3_layer_ops(n, seed=5)
    sweeps = _check(n, ops, tile_bits, relabel_mode=relabel_mode, seed=3)
    if relabel_mode == 2:
        for P, _ in sweeps:
            if P.nrounds == 0:           # a relabel pass: must be safe with source == destination
                T, no = P.T, P.n_outer
                assert sorted(P.tsrc[i] for i in range(T)) == sorted(P.tdst[i] for i in range(T))
                assert [P.osrc[i] for i in range(no)] == [P.odst[i] for i in range(no)]


@pytest.mark.parametrize("n,relabel_mode", [(13, 0), (16, 0), (16, 1), (22, 1), (30, 1), (33, 1)])
def test_tma_layout_of_dense_ladder_sweeps(n, relabel_mode):
    """dense ladder sweeps load their tiles by TMA: the tensor description (box, strides, requests) and the
    shared-memory tables rewritten by apply_tma_layout are consistent, and the rounds' LDS/STS stay conflict-free"""
    gates = [(O.gate_matrix(o[1], o[2]), o[3]) for o in W.qft_ops(n, measure=False)]
    sweeps, perm = PI.plan(n, gates, 12, 3, (relabel_mode << 4) | 0x100)
    got = 0
    for P, _ in sweeps:
        if P.tma_nreq > 0:
            PI.check_tma_tables(P)
            got += 1
        else:
            PI.check_tables(P)
    assert got == len([1 for P, _ in sweeps if P.nrounds > 0]), "every QFT ladder sweep must qualify for TMA loads"


@pytest.mark.parametrize("n,tile_bits,relabel_mode", [(14, 8, 0), (14, 8, 1), (15, 8, 1), (16, 8, 2), (17, 9, 1)])
def test_relabelling_stores_in_the_middle_of_a_plan(n, tile_bits, relabel_mode):
    """Planner mid_relabel (q1t_plan_dump balance bit 9): a ladder sweep whose targets are high index bits stores its tile
    into the contiguous low block; the later sweeps are planned in the layout that leaves, the qubit map is composed with
    it, and the final layout is restored as usual.  Checked for QFT-like circuits (all ladders) against the oracle."""
    ops = W.qft_ops(n, measure=False)
    if n == 14:
        ops = ops + W.qft_ops(n, measure=False, swaps=False)
    plain = _check(n, ops, tile_bits, relabel_mode=relabel_mode)
    sweeps = _check(n, ops, tile_bits, balanced=0x200, relabel_mode=relabel_mode)
    assert len(sweeps) <= len(plain) + (1 if relabel_mode == 1 else 3)
    mids = [P for P, _ in sweeps[:-1] if P.relabel and P.nrounds > 0]      # (gate sweeps: the final in-place passes of mode 2 also relabel)
    assert mids, "no sweep in the middle of the plan stores relabelled"
    for P in mids:
        # the store of such a sweep covers whole contiguous tiles: its destination tile bits are 0..T-1
        assert sorted(P.tdst[i] for i in range(P.T)) == list(range(P.T))


def test_mid_plan_relabel_with_mixed_gates():
    """gates that do not run as ladders (controls, dense 2x2) between QFT blocks: only ladder sweeps relabel, the
    positions of everything planned afterwards follow"""
    n = 13
    ops = W.qft_ops(n, measure=False) + _random_ops(n, 1, seed=5) + W.qft_ops(n, measure=False)
    _check(n, ops, 8, balanced=0x200, relabel_mode=1)
    _check(n, ops, 9, coalesce=2, balanced=0x201, relabel_mode=0)


@pytest.mark.parametrize("n,tile_bits,balanced,relabel_mode", [(14, 8, 1, 1), (15, 8, 1, 1), (16, 8, 0, 1), (16, 9, 1, 2), (17, 9, 1, 1), (13, 8, 1, 0), (18, 12, 1, 1), (19, 12, 1, 1)])
def test_relabelling_stores_that_rotate_the_next_targets_into_the_low_bits(n, tile_bits, balanced, relabel_mode):
    """Planner mid_relabel 2 (q1t_plan_dump balance bit 10): the store of a sweep also moves the next targets, which ride
    in its tile as passengers, into the low (coalescing) positions, where the next sweep gets them for free"""
    for ops in (W.qft_ops(n, measure=False), W.qft_ops(n, measure=False, swaps=False),
                W.qft_ops(n, measure=False) + _random_ops(n, 1, seed=n) + W.qft_ops(n, measure=False, swaps=False)):
        plain = _check(n, ops, tile_bits, balanced=balanced, relabel_mode=relabel_mode)
        sweeps = _check(n, ops, tile_bits, balanced=balanced | 0x400, relabel_mode=relabel_mode)
        assert len(sweeps) <= len(plain) + 3


def test_rotating_stores_balance_a_qft():
    """QFT-16 with 2^8 tiles: 8 + 5 + 3 steps (5 + 3, 5, 3 -> four rounds and a short last sweep) become sweeps of two
    full rounds where the tile allows it; fewer or equal rounds in total"""
    n, T = 16, 8
    ops = W.qft_ops(n, measure=False)
    gates = [(O.gate_matrix(o[1], o[2]), o[3]) for o in ops if o[0] == "gate"]
    a, _ = PI.plan(n, gates, T, 3, 0x201 | 0x10)
    b, _ = PI.plan(n, gates, T, 3, 0x401 | 0x10)
    rounds = lambda sw: sum(P.nrounds for P, _ in sw)
    assert len(b) <= len(a) and rounds(b) <= rounds(a)


@pytest.mark.parametrize("seed", range(8))
def test_relabelling_stores_on_random_ladder_circuits(seed):
    """H layers in random qubit orders with random controlled phases in between (everything the ladder rounds take), so that
    the sweeps' targets, the passengers that ride along and the low positions they are rotated into differ from case to
    case; both planner modes, both packings, all three ways of restoring the layout"""
    rs = np.random.default_rng(100 + seed)
    n = int(rs.integers(13, 17))
    tile_bits = int(rs.integers(8, 11))
    ops = []
    for _ in range(int(rs.integers(1, 4))):
        order = [int(q) for q in rs.permutation(n)]
        for i, q in enumerate(order):
            ops.append(("gate", "h", (), [q]))
            for p in order[i + 1:]:
                if rs.random() < 0.5:
                    ops.append(("gate", "cu1", (float(rs.uniform(-3, 3)),), [p, q]))
        if rs.random() < 0.5:
            a, b = [int(x) for x in rs.permutation(n)[:2]]
            ops.append(("gate", "swap", (), [a, b]))
    for flag in (0x200, 0x400):
        for balanced in (0, 1):
            _check(n, ops, tile_bits, balanced=balanced | flag, relabel_mode=int(rs.integers(0, 3)), seed=seed)
