mkdir -p gpurun_out
Q1T_DEBUG_BCAST=1 Q1T_SWEEP_LOG=1 timeout 300 python - <<'PY' 2>&1 | grep -v "launch 0.0[0-4]" | tail -40
import sys; sys.path.insert(0, '.')
import numpy as np
from q1tsim_b200 import engine as E, workloads as W
n = 30
for inplace in (0, 1):
    for swaps in (False,):
        st = E.VectorState(n, 64)
        st.set_option("inplace_relabel", 1 if inplace else -1)
        gates = [(E.gate_matrix(o[1], o[2]), o[3]) for o in W.qft_ops(n, measure=False, swaps=swaps)]
        res = np.zeros(64, dtype=np.uint64)
        for rep in range(2):
            st.reset_all()
            if rep == 1:
                st.set_timing(True); st.reset_stats()
                print("---- inplace", inplace, "swaps", swaps, flush=True)
            for m, b in gates:
                st.apply_gate(m, b)
            st.measure_all_into(list(range(n)), res, E.Rng(seed=1))
        print(st.stats(), flush=True)
        st.close()
PY
