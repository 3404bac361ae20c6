"""debug: fused remap read on one GPU with 2 processes"""
import os, sys, socket
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import numpy as np

def worker(rank, world, port, n, q, fused):
    sys.path.insert(0, ROOT)
    if fused:
        os.environ["Q1T_FUSED_REMAP"] = "1"
    import torch, torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from q1tsim_b200 import engine as E, sharded as S, workloads as W
    G = O.gate_matrix
    out = {"rank": rank}
    coefs = W.product_state_coefs(n, seed=7)
    cases = {
        "h0": [("g", "h", [], [0])],
        "h0_h5": [("g", "h", [], [0]), ("g", "h", [], [5])],
        "h0_cu1": [("g", "h", [], [0]), ("g", "cu1", [0.3], [1, 0]), ("g", "h", [], [1])],
        "qft": W.qft_ops(n, measure=False),
    }
    only = os.environ.get("DBG_CASES")
    for name, ops in cases.items():
        if only and name not in only.split(","):
            continue
        sp = S.ShardedState.from_qubit_coefs(coefs, 16, device=0)
        rp = O.OracleState.from_qubit_coefs(coefs, 16)
        for op in ops:
            m = G(op[1], op[2])
            sp.apply_gate(m, op[3], op[1]); rp.apply_gate(m, op[3])
        full = sp.gather_column(0)
        ref = rp.column(0)
        st = sp.local.st.stats()
        out[name] = (float(np.linalg.norm(full - ref) / np.linalg.norm(ref)), sp.remaps, st["fused_remaps"], st["sweeps"], st["permute_sweeps"])
        if name == "h0" and rank == 0:
            bad = np.nonzero(np.abs(full - ref) > 1e-9)[0]
            out["h0_bad"] = (len(bad), [int(b) for b in bad[:8]], [int(b) for b in bad[-4:]])
        dist.barrier(); sp.local.group_close(); dist.barrier()
    q.put(out)
    dist.destroy_process_group()

if __name__ == "__main__":
    import torch.multiprocessing as mp
    world, n = int(sys.argv[1]), int(sys.argv[2])
    fused = len(sys.argv) < 4 or sys.argv[3] != "0"
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    ps = [ctx.Process(target=worker, args=(r, world, port, n, q, fused)) for r in range(world)]
    [p.start() for p in ps]
    for _ in ps:
        print(q.get(timeout=600), flush=True)
    [p.join(60) for p in ps]
