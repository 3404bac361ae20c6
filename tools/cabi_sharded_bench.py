#!/usr/bin/env python
"""QFT-(q + log2 P) + measure_all through the ffi.rs-shaped C ABI on a state sharded over P devices of ONE process
(circuit_set_devices): ms per execute(), remaps.   usage: cabi_sharded_bench.py <qubits per shard> <dev,dev,...> [reps]"""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from q1tsim_b200 import circuit as QC, engine as E, workloads as W

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 30
devs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
n = nl + int(round(math.log2(len(devs))))
shots = 8192
c = QC.Circuit(n, n)
c.set_devices(devs)
W.load_ops(c, W.qft_ops(n, measure=True))
rng = E.Rng(seed=2)
ts = []
for r in range(reps + 2):
    t0 = time.perf_counter()
    c.execute(shots, rng)
    cs = c.cstate()
    ts.append(time.perf_counter() - t0)
ts = sorted(ts[2:])
cs = np.asarray(cs, dtype=np.uint64)
# QFT|0..0> is the uniform superposition: amplitudes 2^(-n/2) everywhere after a peek (here: collapsed, so check outcomes)
print(json.dumps({"qubits": n, "devices": devs, "ms_per_execute_median": 1e3 * ts[len(ts) // 2], "ms_per_execute_min": 1e3 * ts[0],
                  "gate_amp_updates_per_s": W.gate_count(W.qft_ops(n)) * float(1 << n) / ts[len(ts) // 2],
                  "counters": c.sharded_counters(), "outcomes_in_range": bool((cs >> np.uint64(n)).max() == 0),
                  "outcomes_distinct": int(np.unique(cs).size)}), flush=True)
