#!/usr/bin/env python
"""Dense user blocks (the unfused fallback): GB/s of one block on k targets of a 2^n state, 32 B per amplitude."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from q1tsim_b200 import engine as E, workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
rs = np.random.default_rng(1)
for k, bits in [(2, [3, 17]), (3, [20, 9, 14]), (3, [n - 1, n - 2, n - 3]), (5, [4, 8, 12, 16, 20]), (6, [3, 6, 9, 12, 15, 18]), (8, list(range(2, 10)))]:
    a = rs.normal(size=(1 << k, 1 << k)) + 1j * rs.normal(size=(1 << k, 1 << k))
    u, _ = np.linalg.qr(a)
    st = E.VectorState.from_qubit_coefs(W.product_state_coefs(n, seed=1), 1)
    st.apply_gate(u, bits, "user"); st.flush()
    st.set_timing(True); st.reset_stats()
    reps = 3
    for _ in range(reps):
        st.apply_gate(u, bits, "user"); st.flush()
    s = st.stats()
    ms = s["sweep_ms"] / reps
    print(json.dumps({"n": n, "k": k, "bits": bits, "ms": ms, "GBps": (32 << n) / (ms * 1e-3) / 1e9, "fallback_sweeps": s["fallback_sweeps"] / reps,
                      "norm": st.column_totals()[0]}), flush=True)
    st.close()
