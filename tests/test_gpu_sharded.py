"""Parity of the sharded engine (one process per shard) against the oracle.

* `test_sharded_engine_vs_oracle`: one process per GPU over NCCL (needs >= 2 CUDA devices:
  `gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`).
* `test_sharded_engine_on_one_gpu`: the SAME code path -- CUDA IPC peer mappings, mailbox barriers, the in-place
  multi-bit remap kernel, replicated start -- with 2, 4 and 8 processes that all use device 0 (gloo carries the few
  host-side messages; the GPU time-slices between the processes), so a 1-GPU box exercises every CUDA kernel of the
  multi-GPU path."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import oracle as O
        from q1tsim_b200 import engine as E
        from q1tsim_b200 import sharded as S
        from q1tsim_b200 import workloads as W
        shots = 1000
        words = O.splitmix64_words(21, 8 * shots + 64)
        out = {"rank": rank}
        G = O.gate_matrix
        # ---- gates on every qubit incl. the global ones, swaps, controlled gates ----
        st = S.ShardedState(n, shots)
        ref = O.OracleState(n, shots, mode=1, order=1)
        O.lib().orc_set_threads(4)
        rs = np.random.default_rng(5)
        names = [("h", 0), ("x", 0), ("u3", 3), ("cx", 0), ("cu1", 1), ("cs", 0), ("swap", 0), ("rz", 1), ("ccx", 0), ("crx", 1), ("t", 0), ("cz", 0)]
        for rep in range(36):
            name, npar = names[rep % len(names)]
            m = G(name, list(rs.uniform(-2, 2, size=npar)))
            k = int(np.log2(m.shape[0]))
            bits = [int(b) for b in rs.permutation(n)[:k]]
            st.apply_gate(m, bits, name); ref.apply_gate(m, bits)
        full = st.gather_column(0)
        out["gates_rel_l2"] = float(np.linalg.norm(full - ref.column(0)) / np.linalg.norm(ref.column(0)))
        out["exchanges"] = st.exchanges
        # ---- QFT + sampling: identical amplitudes on both sides -> bit-exact outcomes ----
        st = S.ShardedState(n, shots)
        ref = O.OracleState(n, shots, mode=1, order=1)
        for op in W.u3_layer_ops(n, seed=1) + W.qft_ops(n, measure=False):
            m = G(op[1], op[2])
            st.apply_gate(m, op[3], op[1]); ref.apply_gate(m, op[3])
        out["qft_rel_l2"] = float(np.linalg.norm(st.gather_column(0) - ref.column(0)))
        out["exchanges"] += st.exchanges
        st.canonicalize()
        psi = ref.column(0)
        nl = 1 << st.n_local
        st.local.write_column(0, psi[rank * nl:(rank + 1) * nl])
        for qb in (0, 1, n - 1):
            out["w0_%d" % qb] = (float(st.marginal0(qb)[0]), float(ref.marginal0(qb, order=1)[0]))
        cb = list(range(n))
        rs_, ro = np.zeros(shots, dtype=np.uint64), np.zeros(shots, dtype=np.uint64)
        rng_s, rng_o = E.Rng(words=words), O.Rng(words=words)
        st.peek_all_into(cb, rs_, rng_s); ref.peek_all_into(cb, ro, rng_o)
        out["peek_equal"] = bool(np.array_equal(rs_, ro))
        st.measure_into(0, 40, rs_, rng_s); ref.measure_into(0, 40, ro, rng_o)           # a global qubit
        st.measure_into(n - 1, 41, rs_, rng_s); ref.measure_into(n - 1, 41, ro, rng_o)   # a local qubit
        out["measure_equal"] = bool(np.array_equal(rs_, ro))
        out["counts_equal"] = st.counts == ref.counts
        st.measure_all_into(cb, rs_, rng_s); ref.measure_all_into(cb, ro, rng_o)
        out["measure_all_equal"] = bool(np.array_equal(rs_, ro))
        out["consumed"] = (rng_s.consumed, rng_o.consumed)
        q.put(out)
    finally:
        dist.destroy_process_group()


def _worker_one_gpu(rank, world, port, n, q, fused=False):
    sys.path.insert(0, ROOT)
    if fused:
        os.environ["Q1T_FUSED_REMAP"] = "1"      # read when an engine state is created (engine.cu)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from q1tsim_b200 import engine as E
        from q1tsim_b200 import sharded as S
        from q1tsim_b200 import workloads as W
        shots = 500
        G = O.gate_matrix
        out = {"rank": rank}
        O.lib().orc_set_threads(2)
        ops = W.qft_ops(n, measure=False)
        cb = list(range(n))
        # (a) QFT from |0..0>, twice through reset_all: replicated start (no exchange for the gates), one multi-bit
        #     remap for the canonical layout, sampling bit-exact on identical amplitudes
        st = S.ShardedState(n, shots, device=0)
        out["group_ok"] = bool(st.group_ok)
        for rep in range(2):
            if rep:
                st.reset_all()
            st.run_ops(ops, G)
            ref = O.OracleState(n, shots, mode=1, order=1)
            for op in ops:
                ref.apply_gate(G(op[1], op[2]), op[3])
            full = st.gather_column(0)
            out["qft_rel_l2_%d" % rep] = float(np.linalg.norm(full - ref.column(0)) / np.linalg.norm(ref.column(0)))
            out["remaps_%d" % rep] = st.remaps
            st.remaps = 0
        # (b) dense input: product state, QFT with the look-ahead remap planner, against the oracle and the closed form
        coefs = W.product_state_coefs(n, seed=7)
        sp = S.ShardedState.from_qubit_coefs(coefs, shots, device=0)
        rp = O.OracleState.from_qubit_coefs(coefs, shots)
        sp.run_ops(ops, G)
        for op in ops:
            rp.apply_gate(G(op[1], op[2]), op[3])
        full = sp.gather_column(0)
        out["product_rel_l2"] = float(np.linalg.norm(full - rp.column(0)) / np.linalg.norm(rp.column(0)))
        out["product_closed_form"] = float(np.linalg.norm(full - W.qft_of_product_state(n, coefs, np.arange(1 << n))))
        out["remaps_product"] = sp.remaps
        out["fused_remaps_product"] = int(sp.local.st.stats().get("fused_remaps", 0))
        # (c) identical amplitudes on both sides -> bit-exact outcomes through the rank-ordered canonical chain
        psi = rp.column(0)
        nl = 1 << sp.n_local
        sp.local.write_column(0, psi[rank * nl:(rank + 1) * nl])
        words = O.splitmix64_words(33, 4 * shots + 64)
        rng_s, rng_o = E.Rng(words=words), O.Rng(words=words)
        rs_, ro = np.zeros(shots, dtype=np.uint64), np.zeros(shots, dtype=np.uint64)
        sp.measure_into(0, 40, rs_, rng_s); rp.measure_into(0, 40, ro, rng_o)
        sp.measure_all_into(cb, rs_, rng_s); rp.measure_all_into(cb, ro, rng_o)
        out["measure_equal"] = bool(np.array_equal(rs_, ro))
        # (d) random gates on every qubit incl. pinned and unpinned global ones
        st.reset_all()
        ref = O.OracleState(n, shots, mode=1, order=1)
        rs = np.random.default_rng(11)
        names = [("h", 0), ("x", 0), ("u3", 3), ("cx", 0), ("cu1", 1), ("cs", 0), ("swap", 0), ("rz", 1), ("ccx", 0), ("y", 0), ("t", 0), ("cz", 0)]
        for rep in range(30):
            name, npar = names[rep % len(names)]
            m = G(name, list(rs.uniform(-2, 2, size=npar)))
            k = int(np.log2(m.shape[0]))
            bits = [int(b) for b in rs.permutation(n)[:k]]
            st.apply_gate(m, bits, name); ref.apply_gate(m, bits)
        out["gates_rel_l2"] = float(np.linalg.norm(st.gather_column(0) - ref.column(0)) / np.linalg.norm(ref.column(0)))
        dist.barrier()
        st.local.group_close(); sp.local.group_close()
        dist.barrier()
        q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 16), (4, 17), (8, 18), (2, 23)])
def test_sharded_engine_on_one_gpu(world, n):
    import torch.multiprocessing as mp
    if _ngpu() < 1:
        pytest.skip("needs a GPU")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_one_gpu, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for o in outs:
        assert o["group_ok"]
        assert o["qft_rel_l2_0"] < 1e-10 and o["qft_rel_l2_1"] < 1e-10
        assert o["remaps_0"] == 0 and o["remaps_1"] == 0          # QFT from |0..0>: no exchange at all (replicated start + initial layout)
        assert o["product_rel_l2"] < 1e-10 and o["product_closed_form"] < 1e-10
        assert 1 <= o["remaps_product"] <= 2                      # dense input: one multi-bit remap (kernels.cu group_swap_kernel)
        assert o["measure_equal"]
        assert o["gates_rel_l2"] < 1e-10


@pytest.mark.parametrize("world,n", [(2, 16), (4, 17), (8, 18), (2, 22)])
def test_sharded_engine_on_one_gpu_fused_remap(world, n):
    """Option fused_remap: the remap of the dense QFT is not a swap pass; the sweep that follows reads its tiles from the
    peers' shards (kernels.cu ladder_kernel<.., REMOTE>, engine.cu issue_sweeps).  Same checks as above, plus: the
    fused path did run for the dense input, and remaps that no ladder sweep follows fall back to the swap pass."""
    import torch.multiprocessing as mp
    if _ngpu() < 1:
        pytest.skip("needs a GPU")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_one_gpu, args=(r, world, port, n, q, True)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for o in outs:
        assert o["group_ok"]
        assert o["qft_rel_l2_0"] < 1e-10 and o["qft_rel_l2_1"] < 1e-10
        assert o["product_rel_l2"] < 1e-10 and o["product_closed_form"] < 1e-10
        assert 1 <= o["remaps_product"] <= 2
        assert o["fused_remaps_product"] >= 1
        assert o["measure_equal"]
        assert o["gates_rel_l2"] < 1e-10


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world,n", [(2, 16), (2, 21), (4, 20), (8, 21)])
def test_sharded_engine_vs_oracle(world, n):
    import torch.multiprocessing as mp
    if _ngpu() < world:
        pytest.skip("not enough GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for o in outs:
        assert o["gates_rel_l2"] < 1e-10 and o["qft_rel_l2"] < 1e-10
        assert o["exchanges"] >= 1
        for k, v in o.items():
            if k.startswith("w0_"):
                assert v[0] == v[1], (k, v)
        assert o["peek_equal"] and o["measure_equal"] and o["counts_equal"] and o["measure_all_equal"]
        assert o["consumed"][0] == o["consumed"][1]
