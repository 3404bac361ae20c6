#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (cfg1, cfg2, cfg4) through the ffi.rs-compatible
Circuit ABI on one GPU: wall time per execute(), engine sweep statistics, and the CPU
oracle's time for the same circuit where it finishes in seconds.  One JSON line each."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from q1tsim_b200 import circuit as QC  # noqa: E402
from q1tsim_b200 import engine as E  # noqa: E402
from q1tsim_b200 import workloads as W  # noqa: E402


def run(name, nq, ops, shots, reps=5, oracle_too=True):
    c = QC.Circuit(nq, nq)
    W.load_ops(c, ops)
    c.execute(shots, E.Rng(seed=1))            # warm-up
    c.execute(shots, E.Rng(seed=1))
    ts = []
    for r in range(reps):
        t0 = time.perf_counter()
        c.execute(shots, E.Rng(seed=2 + r))
        cs = c.cstate()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    dt = ts[len(ts) // 2]                      # median; the minimum is reported too (host latency of a fresh box is noisy)
    st = c.engine_stats()
    line = {"config": name, "qubits": nq, "gates": W.gate_count(ops), "shots": shots, "ms_per_execute": 1e3 * dt,
            "ms_per_execute_min": 1e3 * ts[0], "reps": reps,
            "gate_amp_updates_per_s": W.gate_count(ops) * float(1 << nq) / dt,
            "sweeps": st["sweeps"], "fallback_sweeps": st["fallback_sweeps"], "kernel_launches": st["kernel_launches"],
            "distinct_outcomes": int(len(set(cs.tolist())))}
    if oracle_too:
        from oracle import oracle as O
        o = O.OracleCircuit(nq, nq, mode=0, order=0)
        W.load_ops(o, ops)
        t0 = time.perf_counter()
        o.execute(shots, O.Rng(seed=2))
        line["cpu_port_ms"] = 1e3 * (time.perf_counter() - t0)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    run("cfg1 README QFT-3 + measure_all, 8192 runs", 3, W.qft_ops(3), 8192, reps=15)
    run("cfg2 random-20 depth 100 + measure_all, 1024 runs", 20, W.random_circuit_ops(20, 100), 1024, reps=15, oracle_too="--no-oracle" not in sys.argv)
    run("cfg4 GHZ-24 + X/Y/Z mid-circuit measurements + conditional gates", 24, W.ghz_branching_ops(24), 1024, reps=9, oracle_too=False)
    run("cfg4 (small) GHZ-16 branching", 16, W.ghz_branching_ops(16), 1024, reps=15, oracle_too="--no-oracle" not in sys.argv)
