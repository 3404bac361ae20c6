mkdir -p gpurun_out
Q1T_SWEEP_LOG=1 timeout 300 python - <<'PY' 2>&1 | grep -v "launch 0.0[0-4]" | tail -40
import sys; sys.path.insert(0, '.')
import numpy as np
from q1tsim_b200 import engine as E, workloads as W
for n, swaps, topdown in ((30, False, False), (30, True, False), (31, False, True), (33, False, True)):
    st = E.VectorState(n, 64)
    ops = W.qft_ops(n, measure=False, swaps=swaps)
    gates = [(E.gate_matrix(o[1], o[2]), [(n - 1 - b) if topdown else b for b in o[3]]) for o in ops]
    res = np.zeros(64, dtype=np.uint64)
    for rep in range(2):
        st.reset_all()
        if rep == 1:
            st.set_timing(True); st.reset_stats()
            print("---- n", n, "swaps", swaps, "topdown", topdown, flush=True)
        for m, b in gates:
            st.apply_gate(m, b)
        st.measure_all_into(list(range(n)), res, E.Rng(seed=1))
    s = st.stats()
    print({k: s[k] for k in ("sweeps", "sweep_ms", "read_ms", "sweep_bytes")}, flush=True)
    st.close()
PY
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_n1_e.json 2>/dev/null; cut -c1-330 gpurun_out/bench_r2_n1_e.json
