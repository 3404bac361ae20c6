#!/usr/bin/env python
"""Timing probe (results are WRONG in the skip modes): how long do the QFT-30 sweeps take with
their global loads and/or stores removed?  Separates the memory phase from the register rounds."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from q1tsim_b200 import engine as E, workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ops = W.qft_ops(n, measure=False)
gates = [(E.gate_matrix(o[1], o[2]), o[3]) for o in ops]
for skip in (0, 1, 2, 3):
    st = E.VectorState(n, 1)
    st.set_option("dbg_skip", skip)
    for rep in range(3):
        st.reset_all()
        if rep == 1:
            st.set_timing(True); st.reset_stats()
        # a non-trivial input so that sweep 1 also loads: write one amplitude (materialises the column)
        st.set_column(0, np.array([1.0 + 0j]), 0)
        for m, b in gates:
            st.apply_gate(m, b)
        st.flush()
    s = st.stats()
    print(json.dumps({"skip_loads": bool(skip & 1), "skip_stores": bool(skip & 2), "sweeps": s["sweeps"], "avg_sweep_ms": s["sweep_ms"] / s["sweeps"]}), flush=True)
    st.close()
