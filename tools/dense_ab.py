#!/usr/bin/env python
"""Dense-sweep A/B: the QFT-n gate list on a dense product state (every sweep reads and writes 32 B per amplitude),
timed per sweep launch with CUDA events, under engine options given as key=value sets, e.g.
    python tools/dense_ab.py 30 tma=0 tma=1 tma=1,prefetch_ahead=1
The result is checked against the closed form (workloads.qft_of_product_state) on windows of the output."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from q1tsim_b200 import engine as E, workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
configs = [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in a.split(",")) for a in sys.argv[2:]] or [{}]
coefs = W.product_state_coefs(n, seed=1)
gates = [(E.gate_matrix(o[1], o[2]), o[3]) for o in W.qft_ops(n, measure=False)]
peak = 6550.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
for cfg in configs:
    st = E.VectorState.from_qubit_coefs(coefs, 1)
    for k, v in cfg.items():
        st.set_option(k, v)
    for m, b in gates:            # once untimed (plan cache, first-launch costs); the result of this pass is checked
        st.apply_gate(m, b)
    st.flush()
    err = 0.0
    N = 1 << n
    rs = np.random.default_rng(3)
    for off in [0, N - 4096] + [int(v) & ~4095 for v in rs.integers(0, N - 4096, size=6)]:
        idx = np.arange(off, off + 4096, dtype=np.int64)
        want = W.qft_of_product_state(n, coefs, idx)
        got = st.column(0, off, 4096)
        err = max(err, float(np.linalg.norm(got - want) / np.linalg.norm(want)))
    st.set_timing(True); st.reset_stats()
    reps = 3
    for _ in range(reps):
        for m, b in gates:
            st.apply_gate(m, b)
        st.flush()
    s = st.stats()
    ms = s["sweep_ms"] / max(s["sweeps"], 1)
    gbs = s["sweep_bytes"] / max(s["sweeps"], 1) / (ms * 1e-3) / 1e9
    print(json.dumps({"n": n, "cfg": cfg, "sweeps_per_circuit": s["sweeps"] / reps, "tma_sweeps": s.get("tma_sweeps"),
                      "avg_sweep_ms": ms, "circuit_sweep_ms": s["sweep_ms"] / reps, "GBps": gbs, "frac_of_measured_peak": gbs / peak,
                      "window_rel_l2_vs_closed_form": err, "norm": st.column_totals()[0]}), flush=True)
    st.close()
