import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from q1tsim_b200 import engine as E, workloads as W
n, shots = 10, 16
for track in (1, 0):
    e, o = E.VectorState(n, shots), O.OracleState(n, shots, mode=1, order=1)
    e.set_option("track_support", track)
    for q in range(n):
        m = O.gate_matrix("h"); e.apply_gate(m, [q], "H"); o.apply_gate(m, [q])
    words = O.splitmix64_words(11, shots)
    re_, ro = np.zeros(shots, dtype=np.uint64), np.zeros(shots, dtype=np.uint64)
    e.measure_all_into(list(range(n)), re_, E.Rng(words=words)); o.measure_all_into(list(range(n)), ro, O.Rng(words=words))
    print("outcomes equal", np.array_equal(re_, ro), sorted(set(int(x) for x in re_)))
    for op in W.qft_ops(n, measure=False):
        m = O.gate_matrix(op[1], op[2]); e.apply_gate(m, op[3], op[1]); o.apply_gate(m, op[3])
    ec = [e.column(c) for c in range(e.ncols)]; oc = [o.column(c) for c in range(o.ncols)]
    print("track", track, "counts", e.counts, o.counts, e.stats()["sweeps"])
    for c in range(len(oc)):
        errs = [float(np.linalg.norm(ec[c] - oc[k])) for k in range(len(oc))]
        print(c, "err vs own %.2e" % errs[c], "best match col", int(np.argmin(errs)), "%.2e" % min(errs), "norm %.4f" % np.linalg.norm(ec[c]))
