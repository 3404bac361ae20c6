import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from q1tsim_b200 import engine as E, workloads as W, circuit as QC
n, shots = 30, 8192
ops = W.qft_ops(n, measure=True)
rng = E.Rng(seed=2)
def t(label, f, reps=4):
    f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); ts.append(1e3 * (time.perf_counter() - t0))
    print(label, ["%.2f" % x for x in ts], flush=True)
    return r
def full():
    c = QC.Circuit(n, n, 0); W.load_ops(c, ops); c.execute(shots, rng); out = c.cstate(); st = c.engine_stats(); c.close(); return st
print(json.dumps(t("e2e full", full)))
c = QC.Circuit(n, n, 0); W.load_ops(c, ops)
t("construct+load", lambda: (lambda cc: (W.load_ops(cc, ops), cc.close()))(QC.Circuit(n, n, 0)))
t("execute only", lambda: c.execute(shots, rng))
t("cstate", lambda: c.cstate())
print(json.dumps(c.engine_stats()))
gates = [(E.gate_matrix(o[1], o[2]), o[3], o[1]) for o in ops if o[0] == "gate"]
st = E.VectorState(n, shots, 0); res = np.zeros(shots, dtype=np.uint64)
def direct():
    st.reset_all()
    for m, b, name in gates: st.apply_gate(m, b, name)
    st.measure_all_into(list(range(n)), res, rng)
t("direct step", direct)
def direct_nomeas():
    st.reset_all()
    for m, b, name in gates: st.apply_gate(m, b, name)
    st.flush()
t("direct gates+flush", direct_nomeas)
def queue_only():
    st.reset_all()
    for m, b, name in gates: st.apply_gate(m, b, name)
t("direct queue only (python+lowering)", queue_only)
st.flush()
t("measure_all only", lambda: st.measure_all_into(list(range(n)), res, rng))
