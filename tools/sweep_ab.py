#!/usr/bin/env python
"""A/B timing of the QFT sweeps under planner/kernel options (CUDA events around every sweep launch).
usage: [Q1T_LIB=...] python tools/sweep_ab.py [n] [coalesce,balance,tile ...]   e.g.  30 3,-1,12 2,-1,12
Checks the result against the closed form (uniform amplitudes for input |0..0>) on a few probes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from q1tsim_b200 import engine as E, workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
combos = [tuple(int(x) for x in a.split(",")) for a in sys.argv[2:]] or [(3, -1, 12)]
ops = W.qft_ops(n, measure=False)
gates = [(E.gate_matrix(o[1], o[2]), o[3]) for o in ops]
for (coal, bal, tile) in combos:
    st = E.VectorState(n, 1)
    st.set_option("tile_bits", tile)
    st.set_option("coalesce_bits", coal)
    st.set_option("balance", bal)
    for rep in range(4):
        st.reset_all()
        if rep == 1:
            st.set_timing(True); st.reset_stats()
        for m, b in gates:
            st.apply_gate(m, b)
        st.flush()
    s = st.stats()
    err = 0.0
    for off in (0, (1 << n) - 4096, (1 << (n - 1)) + 12345 * 64, 3 << (n - 3)):
        err = max(err, float(np.max(np.abs(st.column(0, off, 4096) - 2.0 ** (-n / 2)))))
    err = max(err, abs(st.column_totals()[0] - 1.0))
    print(json.dumps({"lib": os.path.basename(E.LIB_PATH), "n": n, "coalesce": coal, "balance": bal, "tile": tile,
                      "sweeps_per_circuit": s["sweeps"] / 3, "avg_sweep_ms": s["sweep_ms"] / s["sweeps"],
                      "circuit_sweep_ms": s["sweep_ms"] / 3, "max_abs_err_probe": err}), flush=True)
    st.close()
