"""world_size-2 and -4 `gloo` tests (CPU) of the multi-GPU host logic: the sharded
composition driven over a numpy local backend must reproduce the single-state oracle
bit for bit in everything that is integer or canonical-order (outcomes, counts, w0) and
to 1e-12 in amplitudes."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, case, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from q1tsim_b200 import sharded as S
        from q1tsim_b200 import workloads as W
        from tests import numpy_local
        shots = 300
        words = O.splitmix64_words(11, 8 * shots + 64)
        st = S.ShardedState(n, shots, local_factory=numpy_local.factory)
        ref = O.OracleState(n, shots, mode=1, order=1)
        rng_s, rng_o = O.Rng(words=words), O.Rng(words=words)
        G = O.gate_matrix
        out = {"rank": rank}

        def both(name, params, bits):
            m = G(name, params)
            st.apply_gate(m, bits, name)
            ref.apply_gate(m, bits)

        if case == "gates":
            rs = np.random.default_rng(5)
            names = [("h", 0), ("x", 0), ("u3", 3), ("cx", 0), ("cu1", 1), ("cs", 0), ("swap", 0), ("rz", 1), ("ccx", 0), ("crx", 1), ("t", 0), ("cz", 0)]
            for rep in range(40):
                name, npar = names[rep % len(names)]
                k = int(np.log2(G(name, [0.3] * npar).shape[0]))
                bits = [int(b) for b in rs.permutation(n)[:k]]
                both(name, list(rs.uniform(-2, 2, size=npar)), bits)
            full = st.gather_column(0)
            out["amp_err"] = float(np.linalg.norm(full - ref.column(0)))
            out["exchanges"] = st.exchanges
            # reductions: same amplitudes on both sides -> the rank-ordered chain must be bit-exact
            st.canonicalize()
            psi = ref.column(0)
            nl = 1 << st.n_local
            st.local.write_column(0, psi[rank * nl:(rank + 1) * nl])
            for qb in (0, 1, n - 1):
                out["w0_%d" % qb] = (float(st.marginal0(qb)[0]), float(ref.marginal0(qb, order=1)[0]))
            out["total"] = (float(st.column_totals()[0]), float(ref.column_totals(order=1)[0]))
        elif case == "qft_sample":
            ops = W.u3_layer_ops(n, seed=1) + W.qft_ops(n, measure=False)
            st.run_ops(ops, G)                       # look-ahead remap planning
            for op in ops:
                ref.apply_gate(G(op[1], op[2]), op[3])
            g_ = st.g
            out["canonical_after_run"] = all(st.where[q_] == st._canonical(q_) for q_ in range(g_))
            out["exchanges_run"] = st.exchanges
            out["amp_err"] = float(np.linalg.norm(st.gather_column(0) - ref.column(0)))
            psi = ref.column(0)                       # identical amplitudes -> bit-exact sampling
            nl = 1 << st.n_local
            st.local.write_column(0, psi[rank * nl:(rank + 1) * nl])
            cb = list(range(n))
            rs_, ro = np.zeros(shots, dtype=np.uint64), np.zeros(shots, dtype=np.uint64)
            st.peek_all_into(cb, rs_, rng_s); ref.peek_all_into(cb, ro, rng_o)
            out["peek_equal"] = bool(np.array_equal(rs_, ro))
            st.measure_all_into(cb, rs_, rng_s); ref.measure_all_into(cb, ro, rng_o)
            out["measure_equal"] = bool(np.array_equal(rs_, ro))
            out["counts_equal"] = st.counts == ref.counts
            out["consumed"] = (rng_s.consumed, rng_o.consumed)
            out["exchanges"] = st.exchanges
            # collapsed columns: every column is a basis state held by exactly one rank
            out["col_err"] = float(max(np.linalg.norm(st.gather_column(c) - ref.column(c)) for c in range(min(ref.ncols, 6))))
        elif case == "branching":
            both("h", (), [0]); both("h", (), [n - 1]); both("cx", (), [0, 2]); both("ry", (0.7,), [1])
            cs, co = np.zeros(shots, dtype=np.uint64), np.zeros(shots, dtype=np.uint64)
            for qb, cbit in ((0, 0), (n - 1, 1), (1, 2)):          # global and local qubits
                st.measure_into(qb, cbit, cs, rng_s); ref.measure_into(qb, cbit, co, rng_o)
                both("h", (), [qb])
            ctl = ((cs >> np.uint64(0)) & np.uint64(1)).astype(np.uint8)
            st.apply_conditional_gate(ctl, G("x"), [0], "X"); ref.apply_conditional_gate(ctl, G("x"), [0])
            st.apply_conditional_gate(ctl, G("cz"), [0, 3], "CZ"); ref.apply_conditional_gate(ctl, G("cz"), [0, 3])
            st.reset(0, rng_s); ref.reset(0, rng_o)
            st.peek_into(2, 5, cs, rng_s); ref.peek_into(2, 5, co, rng_o)
            out["cstate_equal"] = bool(np.array_equal(cs, co))
            out["counts_equal"] = st.counts == ref.counts
            out["consumed"] = (rng_s.consumed, rng_o.consumed)
            out["col_err"] = float(max(np.linalg.norm(st.gather_column(c) - ref.column(c)) for c in range(ref.ncols)))
        elif case == "run_ops_basis":
            # cfg4-shaped op list through run_ops: X/Y/Z-basis mid-circuit measurements (circuit.rs:667-703),
            # conditional gates on the classical word, a reset, X-basis measure_all at the end
            ops = W.ghz_branching_ops(n)[:-1] + [("reset", 1), ("gate", "ry", (0.4,), [1]), ("peek", n - 1, 5, "Y"),
                                                  ("measure_all", list(range(n)), "X")]
            cs = np.zeros(shots, dtype=np.uint64)
            st.run_ops(ops, G, cs, rng_s)
            oc = O.OracleCircuit(n, n, mode=1, order=1)
            W.load_ops(oc, ops)
            oc.execute(shots, rng_o)
            out["cstate_equal"] = bool(np.array_equal(cs, oc.c_state))
            out["consumed"] = (rng_s.consumed, rng_o.consumed)
        elif case.startswith("replicated"):
            # replicated start + multi-bit remap (numpy peer-group emulation) + reuse through reset_all: QFT from
            # |0..0> needs no exchange for its gates; the canonical layout comes back in ONE remap
            group = case.endswith("group")
            st = S.ShardedState(n, shots, local_factory=numpy_local.factory if not group else numpy_local.factory_group)
            ops = W.qft_ops(n, measure=False)
            for rep in range(2):
                if rep:
                    st.reset_all()
                ref = O.OracleState(n, shots, mode=1, order=1)
                st.run_ops(ops, G)
                for op in ops:
                    ref.apply_gate(G(op[1], op[2]), op[3])
                out["exchanges_gates_%d" % rep] = st.exchanges
                out["amp_err_%d" % rep] = float(np.linalg.norm(st.gather_column(0) - ref.column(0)))
                out["remaps_%d" % rep] = st.remaps
                out["replayed_%d" % rep] = getattr(st.local, "replayed", 0)
                st.exchanges = st.remaps = 0
            # gates on pinned rank bits: X / Y keep the pin, a controlled gate with a pinned target forces the depin
            st.reset_all()
            ref = O.OracleState(n, shots, mode=1, order=1)
            for name, params, bits in (("x", (), [0]), ("h", (), [n - 1]), ("cx", (), [0, n - 2]), ("y", (), [1]), ("u3", (0.3, 0.2, 0.1), [0]),
                                       ("cz", (), [0, 1]), ("cx", (), [n - 1, 1]), ("h", (), [1]), ("swap", (), [0, n - 3]), ("ry", (0.7,), [0])):
                both(name, params, bits)
            out["amp_err_pins"] = float(np.linalg.norm(st.gather_column(0) - ref.column(0)))
            # dense product-state input (from_qubit_coefs): nothing is pinned, the global H gates need their remap
            coefs = W.product_state_coefs(n, seed=3)
            sp = S.ShardedState.from_qubit_coefs(coefs, shots, local_factory=numpy_local.factory if not group else numpy_local.factory_group)
            rp = O.OracleState.from_qubit_coefs(coefs, shots)
            sp.run_ops(ops, G)
            for op in ops:
                rp.apply_gate(G(op[1], op[2]), op[3])
            out["amp_err_product"] = float(np.linalg.norm(sp.gather_column(0) - rp.column(0)))
            out["remaps_product"] = sp.remaps
            cb = list(range(n))
            rs_, ro = np.zeros(shots, dtype=np.uint64), np.zeros(shots, dtype=np.uint64)
            psi = rp.column(0)
            nl = 1 << sp.n_local
            sp.local.write_column(0, psi[rank * nl:(rank + 1) * nl])
            sp.measure_all_into(cb, rs_, rng_s); rp.measure_all_into(cb, ro, rng_o)
            out["measure_equal"] = bool(np.array_equal(rs_, ro))
        q.put(out)
    finally:
        dist.destroy_process_group()


def _run(world, n, case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return outs


@pytest.mark.parametrize("world,n", [(2, 12), (4, 12)])
def test_sharded_gates_and_marginals(world, n):
    for o in _run(world, n, "gates"):
        assert o["amp_err"] < 1e-12
        assert o["exchanges"] >= 1
        for k, v in o.items():
            if isinstance(v, tuple) and k != "consumed":
                assert v[0] == v[1], (k, v)          # canonical order: bit-exact across ranks


@pytest.mark.parametrize("world,n", [(2, 11), (4, 12)])
def test_sharded_qft_sampling_bit_exact(world, n):
    for o in _run(world, n, "qft_sample"):
        assert o["amp_err"] < 1e-12
        assert o["canonical_after_run"]          # look-ahead: global qubits already home, no extra exchange
        assert o["exchanges_run"] <= 3 * int(np.log2(world))
        assert o["peek_equal"] and o["measure_equal"] and o["counts_equal"]
        assert o["consumed"][0] == o["consumed"][1]
        assert o["col_err"] == 0.0


@pytest.mark.parametrize("world,n", [(2, 11), (4, 12)])
def test_sharded_branching(world, n):
    for o in _run(world, n, "branching"):
        assert o["cstate_equal"] and o["counts_equal"]
        assert o["consumed"][0] == o["consumed"][1]
        assert o["col_err"] < 1e-12


@pytest.mark.parametrize("world,n", [(2, 12), (4, 13)])
def test_sharded_run_ops_basis_changes_and_reset(world, n):
    for o in _run(world, n, "run_ops_basis"):
        assert o["cstate_equal"]
        assert o["consumed"][0] == o["consumed"][1]


@pytest.mark.parametrize("world,n,case", [(2, 12, "replicated"), (4, 13, "replicated"), (2, 12, "replicated_group"), (4, 13, "replicated_group"),
                                          (8, 14, "replicated_group")])
def test_sharded_replicated_start_and_multi_bit_remap(world, n, case):
    g = int(np.log2(world))
    for o in _run(world, n, case):
        for rep in (0, 1):
            assert o["amp_err_%d" % rep] < 1e-12
            # a QFT from |0..0> costs no exchange at all: replicated start for its gates, and the run starts from the
            # layout that its Swap relabels turn into the canonical one (|0..0> is symmetric)
            assert o["remaps_%d" % rep] == 0
        assert o["replayed_0"] == 0 and o["replayed_1"] > 0         # the second run of the op list replays the taped schedule
        assert o["amp_err_pins"] < 1e-12
        assert o["amp_err_product"] < 1e-12
        # a dense input has nothing pinned: the global qubits come on chip in one multi-bit remap (peer group)
        assert 1 <= o["remaps_product"] <= (2 if case.endswith("group") else 3 * g)
        assert o["measure_equal"]
