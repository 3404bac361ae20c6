mkdir -p gpurun_out
Q1T_SWEEP_LOG=1 timeout 300 python - <<'PY' 2>&1 | grep -v "launch 0.0[0-4]" | tail -40
import sys; sys.path.insert(0, '.')
import numpy as np
from oracle import oracle as O
from q1tsim_b200 import engine as E, workloads as W
# bit-exactness of the in-place fused leaf totals against the separate read pass, small size
for n in (14, 20):
    outs = []
    for inplace in (1, -1):
        st = E.VectorState(n, 512)
        st.set_option("inplace_relabel", inplace)
        gates = [(E.gate_matrix(o[1], o[2]), [n - 1 - b for b in o[3]]) for o in W.qft_ops(n, measure=False, swaps=False)]
        for m, b in gates:
            st.apply_gate(m, b)
        res = np.zeros(512, dtype=np.uint64)
        st.measure_all_into(list(range(n)), res, E.Rng(words=O.splitmix64_words(3, 600)))
        outs.append((res.copy(), st.stats()["read_passes"]))
        st.close()
    print("n", n, "outcomes equal", bool(np.array_equal(outs[0][0], outs[1][0])), "read passes (in place, out of place)", outs[0][1], outs[1][1], flush=True)
n = 33
st = E.VectorState(n, 64)
gates = [(E.gate_matrix(o[1], o[2]), [n - 1 - b for b in o[3]]) for o in W.qft_ops(n, measure=False, swaps=False)]
res = np.zeros(64, dtype=np.uint64)
for rep in range(2):
    st.reset_all()
    if rep == 1:
        st.set_timing(True); st.reset_stats(); print("---- n 33 topdown", flush=True)
    for m, b in gates:
        st.apply_gate(m, b)
    st.measure_all_into(list(range(n)), res, E.Rng(seed=1))
s = st.stats(); print({k: s[k] for k in ("sweeps", "sweep_ms", "read_ms", "read_passes")}, flush=True)
PY
