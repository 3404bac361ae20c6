#!/usr/bin/env python
"""A/B of the fused remap read on N GPUs (torchrun): the dense-input QFT leg of bench.py with the remap as a swap pass
(Q1T_FUSED_REMAP=0) and read through by the sweep that follows it (=1).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/fused_remap_ab.py [local qubits]"""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def main():
    import torch
    import torch.distributed as dist
    from q1tsim_b200 import engine as E
    rank, world, local = bench.dist_env()
    dev = local % max(E.lib().q1t_device_count(), 1)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    g = int(round(math.log2(world)))
    nl = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    args = None                                             # (sharded_leg does not read it)
    for fused in (0, 1, 0, 1):
        os.environ["Q1T_FUSED_REMAP"] = str(fused)          # read when the shard's engine state is created
        leg = bench.sharded_leg(args, rank, world, dev, nl + g, 8192, 4, 2, dist, torch, dense=True)
        if rank == 0:
            print(json.dumps({"fused_remap": fused, "qubits": nl + g, "gpus": world, "ms_per_circuit": leg["ms_per_step"],
                              "breakdown_ms": leg["breakdown_ms"], "exchange": {k: leg["exchange"][k] for k in ("remaps_per_step", "swap_kernel_ms_per_step", "gb_per_s_per_direction")},
                              "sweep_launches": leg["sweep_launches_per_step"], "verified": leg["verified"]["ok"],
                              "rel_l2": leg["verified"]["max_over_ranks_rel_l2"]}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
