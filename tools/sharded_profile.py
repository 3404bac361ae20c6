#!/usr/bin/env python
"""Host-side profile of one sharded QFT step (run under torchrun, or alone for a 1-rank "sharded" state):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_profile.py [qubits per shard]
Prints ms per step and, on rank 0, the cProfile of one step."""
import cProfile, io, math, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from q1tsim_b200 import engine as E, sharded as S, workloads as W

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("gloo", rank=0, world_size=1)
nl = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = nl + int(round(math.log2(world)))
shots = 8192
ops = W.qft_ops(n, measure=True)
st = S.ShardedState(n, shots, device=local)
res = np.zeros(shots, dtype=np.uint64)
rng = E.Rng(seed=2)

def step():
    st.reset_all()
    st.run_ops(ops, E.gate_matrix, res, rng)

for _ in range(3):
    step()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print("rank %d: %.2f ms per step, remaps per step %.1f" % (rank, 1e3 * (time.perf_counter() - t0) / 5, st.remaps / 8.0), flush=True)
if rank == 0:
    pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000], flush=True)
else:
    step()
dist.barrier()
st.local.group_close()
dist.barrier()
dist.destroy_process_group()
