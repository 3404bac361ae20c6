#!/usr/bin/env python
"""Where a launch-bound circuit (cfg2: random-20, depth 100, 1500 gates, 102 sweeps) spends its time: host lowering of the
gates, the sweep batch issued launch by launch, the batch replayed as a CUDA graph."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from q1tsim_b200 import engine as E, workloads as W

n, depth = 20, 100
ops = W.random_circuit_ops(n, depth, measure=False)
gates = [(E.gate_matrix(o[1], o[2]), o[3], o[1]) for o in ops]
for graphs in (0, 1):
    st = E.VectorState(n, 1024)
    st.set_option("graphs", graphs)
    t_queue = t_flush = 0.0
    reps = 12
    for rep in range(reps + 3):
        st.reset_all()
        t0 = time.perf_counter()
        for m, b, name in gates:
            st.apply_gate(m, b, name)
        t1 = time.perf_counter()
        st.flush()
        t2 = time.perf_counter()
        if rep >= 3:
            t_queue += t1 - t0
            t_flush += t2 - t1
    s = st.stats()
    print(json.dumps({"graphs": graphs, "queue_1500_gates_ms": 1e3 * t_queue / reps, "flush_ms": 1e3 * t_flush / reps,
                      "sweeps_per_run": s["sweeps"] / (reps + 3), "graph_captures": s["graph_captures"], "graph_replays": s["graph_replays"],
                      "h2d_bytes_per_run": s["h2d_bytes"] / (reps + 3)}), flush=True)
    st.close()
